#!/usr/bin/env python3
"""Chooses the seeds of the synthetic benchmark models.

Weights are U(-1/sqrt(H), 1/sqrt(H)) from seed = 1000 + model_id (+ 10000*k on re-seed).  A model is
accepted only if it is contractive enough for parity to be meaningful: oracle float vs oracle double,
10 s of the sweep+noise signal, max abs <= 5e-5.  The accepted table is committed as
tests/golden/synthetic_models.json; models are re-created from (cell, hidden, input_size, seed) at run time.

Config mapping (BASELINE.json configs): C2 = LSTM-16/1, C3 = GRU-16/3, C5 = LSTM-40/1,
C4 = 64 models cycling {LSTM,GRU} x H in {8,12,16,20,24,32,40} x in {1,2,3}."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from aidadsp_lv2_b200 import models, signals  # noqa: E402
from oracle import pyoracle  # noqa: E402

N = 480000


def inputs(input_size):
    x = np.zeros((N, input_size), dtype=np.float32)
    x[:, 0] = signals.sweep_noise(0, N)
    t = np.arange(N) / 48000.0
    if input_size >= 2:
        x[:, 1] = signals.triangle(1.0, N, buffer=1)
    if input_size >= 3:
        x[:, 2] = signals.triangle(1.7, N, buffer=1)
    return x


def accept(cell, hidden, input_size, seed):
    m = models.synthetic(cell, hidden, input_size, seed)
    x = inputs(input_size)
    yf = pyoracle.OracleModel(m).forward(x)
    yd = pyoracle.OracleModel(m, use_double=True).forward(x)
    err = float(np.abs(yf - yd).max())
    return err <= 5e-5 and float(np.abs(yd).max()) > 1e-3, err


def pick(cell, hidden, input_size, model_id):
    for k in range(20):
        seed = 1000 + model_id + 10000 * k
        ok, err = accept(cell, hidden, input_size, seed)
        if ok:
            return dict(model_id=model_id, cell=cell, hidden=hidden, input_size=input_size, seed=seed, float_vs_double=err)
    raise RuntimeError("no contractive seed found for %s" % ((cell, hidden, input_size),))


def main():
    table = {"c2": pick(0, 16, 1, 0), "c3": pick(1, 16, 3, 1), "c5": pick(0, 40, 1, 2), "c4": []}
    hs = (8, 12, 16, 20, 24, 32, 40)
    for i in range(64):
        cell = i % 2
        hidden = hs[(i // 2) % len(hs)]
        input_size = 1 + (i // 14) % 3
        table["c4"].append(pick(cell, hidden, input_size, 100 + i))
        print(table["c4"][-1], flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "synthetic_models.json"), "w") as f:
        json.dump(table, f, indent=1)
    print("written")


if __name__ == "__main__":
    main()
