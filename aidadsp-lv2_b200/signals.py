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


def _hash32(x):
    x = x.astype(np.uint32)
    x ^= x >> np.uint32(16); x *= np.uint32(0x7feb352d); x ^= x >> np.uint32(15); x *= np.uint32(0x846ca68b); x ^= x >> np.uint32(16)
    return x


def hash_noise_block(n_streams, n_samples, start=0, first_stream=0, seed=0):
    """u(s, n) of the device-side generator (csrc/probes.cu: gen_input_kernel), bit for bit: [n_streams][n_samples] in [-1, 1)."""
    with np.errstate(over="ignore"):
        s = (np.arange(n_streams, dtype=np.uint64) + np.uint64(first_stream)).astype(np.uint32)[:, None]
        n = (np.arange(n_samples, dtype=np.uint64) + np.uint64(start)).astype(np.uint32)[None, :]
        h = _hash32(s * np.uint32(0x9E3779B1) + n * np.uint32(0x85EBCA77) + np.uint32(seed))
    return (h >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -23) - np.float32(1.0)


def sweep_hash_block(n_streams, n_samples, start=0, first_stream=0, seed=0, fs=FS):
    """numpy mirror of aidax_generate_input: the sweep of SURVEY 8(d) with per-stream phase + 0.05 * hash noise, planar."""
    n = np.arange(start, start + n_samples, dtype=np.float64)[None, :]
    t = n / fs
    gs = (np.arange(n_streams, dtype=np.float64) + first_stream)[:, None]
    phase = 6.283185307179586 * (20.0 * t + 999.0 * t * t) + 0.37 * gs
    u = hash_noise_block(n_streams, n_samples, start, first_stream, seed).astype(np.float64)
    return (0.5 * np.sin(phase) + 0.05 * u).astype(np.float32)
