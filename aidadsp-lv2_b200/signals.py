"""Synthetic test signals: a 20 Hz -> 20 kHz linear sweep over 10 s plus uniform
noise, per-stream phase offset and seed, generated in float32 on the host so that the CPU oracle
and the GPU engine consume bit-identical input."""
import numpy as np

FS = 48000


def sweep_noise(stream, n_samples, start=0, seed0=12345, fs=FS, duration_s=10.0):
    """x[n] = 0.5*sin(2*pi*(20 t + (20000-20)/(2*T) t^2) + 0.37*s) + 0.05*u[n],  u ~ U(-1,1) seeded by seed0+s."""
    n = np.arange(start, start + n_samples, dtype=np.float64)
    t = n / fs
    phase = 2.0 * np.pi * (20.0 * t + (20000.0 - 20.0) / (2.0 * duration_s) * t * t) + 0.37 * stream
    rng = np.random.Generator(np.random.Philox(key=seed0 + stream))
    if start:
        rng.bit_generator.advance(start)  # Philox: one draw per sample below (float64 uniform)
    u = rng.uniform(-1.0, 1.0, size=n_samples)
    return (0.5 * np.sin(phase) + 0.05 * u).astype(np.float32)


def sweep_noise_block(n_streams, n_samples, first_stream=0, seed0=12345):
    """[n_streams][n_samples] planar block."""
    out = np.empty((n_streams, n_samples), dtype=np.float32)
    for s in range(n_streams):
        out[s] = sweep_noise(first_stream + s, n_samples, seed0=seed0)
    return out


def triangle(period_s, n_buffers, buffer=128, phase=0.0, fs=FS):
    """Slow triangle automation in [0,1] sampled once per buffer (C3's PARAM1/PARAM2)."""
    t = (np.arange(n_buffers) * buffer / fs) / period_s + phase
    return (2.0 * np.abs(t - np.floor(t + 0.5))).astype(np.float32)
