"""Scratch: the e2e arm alone (N LV2 instances through the batch host, host buffers).
   python tools/e2e_one.py [N] [periods] [threads]     (env: AIDAX_BATCH_CHUNK, AIDAX_KERNEL_STREAMS)"""
import os, sys, time, tempfile
sys.path.insert(0, ".")
import numpy as np
import aidadsp_lv2_b200 as ax
from aidadsp_lv2_b200 import workloads, signals, models

n = int(sys.argv[1]) if len(sys.argv) > 1 else 378880
periods = int(sys.argv[2]) if len(sys.argv) > 2 else 30
threads = int(sys.argv[3]) if len(sys.argv) > 3 else len(os.sched_getaffinity(0))
wl = workloads.get("c2", n)
tmp = tempfile.mkdtemp()
path = os.path.join(tmp, "m.json")
models.save_json(wl.models[0], path)
host = ax.BatchHost(ax.plugin_path(), n_instances=n, max_block=128, n_threads=threads, device=0)
for port in range(4, 25):
    if port != 23:
        host.set_control(port, float(wl.controls[port - 4]))
host.load_model(path)
blk = signals.sweep_noise_block(min(n, 256), 128)
host.input[:, :128] = np.tile(blk, ((n + blk.shape[0] - 1) // blk.shape[0], 1))[:n]
for _ in range(5):
    host.run(128)
tm0 = host.timing()
l0 = host.launch_count
t0 = time.perf_counter()
for _ in range(periods):
    host.run(128)
dt = (time.perf_counter() - t0) / periods
tm1 = host.timing()
print("N=%d threads=%d chunk=%s kstreams=%s: %.2f ms/period = %.2f G samples/s (run() phase %.2f ms, end() %.2f ms, %d launches/period) |out|max=%.3f" % (
    n, threads, os.environ.get("AIDAX_BATCH_CHUNK", "dflt"), os.environ.get("AIDAX_KERNEL_STREAMS", "dflt"), dt * 1e3, n * 128 / dt / 1e9,
    (tm1[0] - tm0[0]) / periods * 1e3, (tm1[1] - tm0[1]) / periods * 1e3, (host.launch_count - l0) // periods, float(np.abs(host.output[:, :128]).max())), flush=True)
host.close()
