#!/usr/bin/env python3
"""CPU emulation of the tensor-core gate contraction's operand precision (decision aid for the tcgen05 path,
SURVEY.md 7 "accept/reject precision by measurement").  NOT the product and not a parity oracle: it restates
the LSTM cell in numpy with the [h, x, 1] . [U; W; b] contraction's operands rounded the way each tcgen05 operand
format would round them, accumulates in fp32, and compares the model output with the C oracle (fp32 reference
restatement) on the same sweep+noise input.

  python tools/precision_probe.py [seconds]

modes: fp32 (control), tf32 (1 product, 10-bit mantissa, truncated), bf16 (1 product), bf16x2 (A split, B hi),
       bf16x3 (Ah.Bh + Ah.Bl + Al.Bh), tf32x3
"""
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aidadsp_lv2_b200  # noqa: E402,F401  (import shim registers the package)
from aidadsp_lv2_b200 import models, signals, workloads  # noqa: E402
from oracle import pyoracle  # noqa: E402


def bf16(x):
    """round-to-nearest-even to bfloat16, returned as float32"""
    u = np.asarray(x, dtype=np.float32).view(np.uint32)
    r = ((u >> 16) & 1) + 0x7FFF
    return ((u + r) & 0xFFFF0000).astype(np.uint32).view(np.float32)


def tf32(x):
    """truncate to tf32 (the tensor core ignores the low 13 mantissa bits)"""
    u = np.asarray(x, dtype=np.float32).view(np.uint32)
    return (u & 0xFFFFE000).astype(np.uint32).view(np.float32)


def split(x, rnd):
    hi = rnd(x)
    lo = rnd((x - hi).astype(np.float32))
    return hi, lo


def contraction(mode, B):
    """returns f(a) -> a @ B for a row vector a, with operand rounding per mode; fp32 accumulate"""
    B = B.astype(np.float32)
    if mode == "fp32":
        return lambda a: a @ B
    rnd = bf16 if mode.startswith("bf16") else tf32
    Bh, Bl = split(B, rnd)
    if mode in ("bf16", "tf32"):
        return lambda a: rnd(a) @ Bh
    if mode == "bf16x2":
        def f(a):
            ah, al = split(a, rnd)
            return ah @ Bh + al @ Bh
        return f

    def f3(a):
        ah, al = split(a, rnd)
        return (ah @ Bh + ah @ Bl) + al @ Bh
    return f3


def run_lstm(m, x, mode):
    H = m.hidden
    # K rows: U (H), W (in = 1), bias; columns Keras order i,f,c,o
    B = np.concatenate([m.U, m.W, m.b.reshape(1, -1)], axis=0).astype(np.float32)
    f = contraction(mode, B)
    h = np.zeros(H, np.float32)
    c = np.zeros(H, np.float32)
    y = np.empty(len(x), np.float32)
    a = np.zeros(H + 2, np.float32)
    a[H + 1] = 1.0
    dw = m.dense_w.astype(np.float32)
    for t in range(len(x)):
        a[:H] = h
        a[H] = x[t]
        z = f(a).astype(np.float32)
        i = 1.0 / (1.0 + np.exp(-z[:H]))
        fg = 1.0 / (1.0 + np.exp(-z[H:2 * H]))
        g = np.tanh(z[2 * H:3 * H])
        o = 1.0 / (1.0 + np.exp(-z[3 * H:]))
        c = (i * g + fg * c).astype(np.float32)
        h = (o * np.tanh(c)).astype(np.float32)
        y[t] = h @ dw + np.float32(m.dense_b)
    return y


def main():
    secs = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
    n = int(48000 * secs)
    x = signals.sweep_noise_block(1, 480000)[0][:n]
    cases = [("bundled LSTM-12 " + os.path.basename(p)[:24], models.load_json(p))
             for p in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "models", "*.json")))[:2]]
    cases.append(("synthetic LSTM-16 (C2)", workloads.get("c2").models[0]))
    cases.append(("synthetic LSTM-40 (C5)", workloads.get("c5").models[0]))
    print("%-44s %-8s %12s %10s" % ("model", "mode", "max abs err", "SNR dB"))
    for name, m in cases:
        ref = pyoracle.OracleModel(m).forward(x)
        for mode in ("fp32", "tf32", "bf16", "bf16x2", "bf16x3", "tf32x3"):
            y = run_lstm(m, x, mode)
            err = float(np.abs(y - ref).max())
            snr = 10 * np.log10(float(np.sum(ref.astype(np.float64) ** 2)) / max(float(np.sum((y - ref).astype(np.float64) ** 2)), 1e-300))
            print("%-44s %-8s %12.3g %10.1f" % (name, mode, err, snr), flush=True)


if __name__ == "__main__":
    main()
