"""Scratch throughput probe (not the contract bench): N streams x LSTM-H, device-resident, wall clock."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
import aidadsp_lv2_b200 as ax
from aidadsp_lv2_b200 import models

def run(n_streams, hidden, cell=0, tile_mode=0, iters=20, input_size=1, eq=True):
    eng = ax.Engine(max_streams=n_streams, max_block=128)
    eng.configure(tile_mode=tile_mode, math_mode=0)
    mid = eng.register_model(models.synthetic(cell, hidden, input_size, seed=1))
    c = ax.default_controls(EQPOS=1.0 if eq else 0.0, EQBYPASS=0.0 if eq else 1.0)
    for i in range(n_streams):
        s = eng.open_stream(); eng.set_model(s, mid); eng.set_controls(s, c, loading=False)
    x = torch.rand(128, n_streams, device="cuda") - 0.5
    y = torch.empty_like(x)
    for _ in range(3):
        eng.render_device(x.data_ptr(), y.data_ptr(), n_streams, 128)
    eng.synchronize()
    ext = torch.cuda.ExternalStream(eng.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(ext)
    for _ in range(iters):
        eng.render_device(x.data_ptr(), y.data_ptr(), n_streams, 128)
    e1.record(ext)
    eng.synchronize()
    wall = (time.perf_counter() - t0) / iters
    dt = e0.elapsed_time(e1) * 1e-3 / iters
    print("  (wall %.1f us/call)" % (wall * 1e6), end="")
    sps = n_streams * 128 / dt
    print("N=%7d H=%2d cell=%d in=%d tile=%d: %8.1f us/buffer  %.3e samples/s  = %8.0f RT streams  (%.2f TFLOP/s model)" % (
        n_streams, hidden, cell, input_size, tile_mode, dt * 1e6, sps, sps / 48000, sps * (8 * hidden * (hidden + input_size) + 10 * hidden + 1) / 1e12), flush=True)
    eng.close()

if __name__ == "__main__":
    for n, tm in [(1024, 1), (1024, 0), (1024, 2), (8192, 0), (65536, 2), (131072, 2), (262144, 2)]:
        run(n, 16, tile_mode=tm)
    run(65536, 40, tile_mode=2); run(1024, 40, tile_mode=1)
    run(8192, 16, cell=1, input_size=3, tile_mode=0)
    run(131072, 12, tile_mode=2); run(131072, 8, tile_mode=2)
