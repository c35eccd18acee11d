#!/usr/bin/env python3
"""CPU emulation of the gate-nonlinearity approximations (decision aid, NOT the product, not a parity oracle): the LSTM
cell in numpy fp32 with different sigmoid/tanh evaluations, 10 s of sweep+noise through a bundled (trained) LSTM-12
capture, compared with the C oracle.  Shows how much of the 10-second error budget each approximation spends.
  python tools/gate_probe.py [model_index] [seconds]"""
import glob, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aidadsp_lv2_b200  # noqa
from aidadsp_lv2_b200 import models, signals
from oracle import pyoracle

f32 = np.float32
L2E = f32(1.4426950408889634)
rng = np.random.default_rng(1)

def ex2(x, ulp):
    y = np.exp2(x.astype(np.float64)).astype(np.float32)
    if ulp:
        y = (y * (1 + rng.uniform(-ulp, ulp, size=y.shape) * 2.0 ** -23)).astype(np.float32)
    return y

def rcp(x, ulp):
    y = (1.0 / x.astype(np.float64)).astype(np.float32)
    if ulp:
        y = (y * (1 + rng.uniform(-ulp, ulp, size=y.shape) * 2.0 ** -23)).astype(np.float32)
    return y

def gates(mode):
    if mode == "libm":
        return (lambda x: (1.0 / (1.0 + np.exp(-x))).astype(np.float32)), (lambda x: np.tanh(x).astype(np.float32))
    if mode.startswith("e"):  # e<ex2 ulp>r<rcp ulp>, e.g. e2r0
        eu, ru = [float(v) for v in mode[1:].split("r")]
    else:
        eu, ru = {"fast_exact": (0, 0), "fast_mufu": (1.0, 0.5), "fast_mufu2": (2.0, 1.0), "tanh_fix": (1.0, 0.5)}[mode]
    def sig(x):
        return rcp((f32(1) + ex2((x * -L2E).astype(np.float32), eu)).astype(np.float32), ru)
    def tanh(x):
        t = (f32(1) - f32(2) * rcp((f32(1) + ex2((x * (2 * L2E)).astype(np.float32), eu)).astype(np.float32), ru)).astype(np.float32)
        if mode == "tanh_fix":
            x2 = (x * x).astype(np.float32)
            p = (x * (f32(1) + x2 * (f32(-1 / 3) + x2 * (f32(2 / 15) + x2 * f32(-17 / 315))))).astype(np.float32)
            t = np.where(np.abs(x) < f32(0.25), p, t)
        return t
    return sig, tanh

def run(m, x, mode):
    H = m.hidden
    B = np.concatenate([m.U, m.W, m.b.reshape(1, -1)], axis=0).astype(np.float32)
    sig, tanh = gates(mode)
    h = np.zeros(H, f32); c = np.zeros(H, f32); y = np.empty(len(x), f32)
    a = np.zeros(H + 2, f32); a[H + 1] = 1.0
    dw = m.dense_w.astype(f32)
    for t in range(len(x)):
        a[:H] = h; a[H] = x[t]
        z = (a @ B).astype(f32)
        i = sig(z[:H]); fg = sig(z[H:2 * H]); g = tanh(z[2 * H:3 * H]); o = sig(z[3 * H:])
        c = (i * g + fg * c).astype(f32)
        h = (o * tanh(c)).astype(f32)
        y[t] = h @ dw + f32(m.dense_b)
    return y

if __name__ == "__main__":
    idx = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    secs = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
    p = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "models", "*.json")))[idx]
    m = models.load_json(p)
    x = signals.sweep_noise_block(1, 480000)[0][:int(48000 * secs)]
    ref = pyoracle.OracleModel(m).forward(x)
    print(os.path.basename(p))
    for mode in sys.argv[3:] or ("libm", "fast_exact", "fast_mufu", "fast_mufu2", "tanh_fix"):
        y = run(m, x, mode)
        err = np.abs(y - ref)
        print("%-12s max abs %.3e   per 2 s: %s" % (mode, err.max(), " ".join("%.1e" % e.max() for e in np.array_split(err, 5))), flush=True)
