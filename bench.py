#!/usr/bin/env python3
"""bench.py -- the hot-path benchmark (contract: one JSON line on stdout from rank 0).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4|c5|capacity]

A step = ONE audio period: every stream of the workload renders one 128-sample buffer through the full
rt-neural-generic chain (in_lpf -> pre-gain -> EQ -> LSTM/GRU+Dense -> dc blocker -> EQ -> master gain).
Default workload = BASELINE.json configs[1] (C2): 1024 streams x one LSTM-16 model, fused pre-EQ.

  value   whole-job stream-samples per second with the period's input already resident in HBM
          (stream-interleaved [128][N] device buffers), CUDA-event timed on the engine's stream.
  e2e     the same metric through the reference-facing surface: N LV2 plugin instances driven by the
          batch-render host with HOST (pinned) port buffers -- control decode, H2D copy, kernels, D2H
          copy and the final synchronise are all inside the timed region (wall clock).
  roofline    FP32 FFMA pipe of the recurrence kernel: algorithmic flop per launch /
              mean CUDA-event launch duration vs the FFMA throughput measured on this GPU in this run.
  cpu_baseline  the UNMODIFIED reference (oracle/_ref/rt-neural-generic.so, xsimd SSE2) on the host cores.
  capacity    same chain at the stream count that fills one GPU: the real-time streams-per-GPU figure.

--impl reference times the reference's own CPU implementation on all host cores (rank 0 only).
Multi-GPU (torchrun): streams shard over ranks with no data-path collective (weak scaling: 1024 per GPU).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T = 128
FS = 48000
L2_BYTES = 126 * 1024 * 1024
METRIC = "stream-samples/s through the full rt-neural-generic chain (real-time 48 kHz streams = value/48000; us per 128-sample buffer in us_per_buffer)"


class Silence:
    """The reference prints layer info to stdout while parsing models (rt-neural-generic.cpp:1034); keep it off our JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.null)
        os.close(self.saved)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def write_models(wl, tmpdir):
    from aidadsp_lv2_b200 import models
    paths = []
    for i, m in enumerate(wl.models):
        p = os.path.join(tmpdir, "model_%02d.json" % i)
        models.save_json(m, p)
        paths.append(p)
    return paths


def setup_engine(eng, wl, n_local, first_global, n_global):
    """Open n_local streams (global ids first_global..), bind models in contiguous runs, latch controls."""
    mids = [eng.register_model(m) for m in wl.models]
    for s in range(n_local):
        sid = eng.open_stream(FS)
        eng.set_model(sid, mids[wl.model_of_stream(first_global + s, n_global)])
        eng.set_controls(sid, wl.controls, loading=False)
    return mids


def device_bench(wl, n_local, first_global, n_global, steps, warmup, device, tile_mode=0):
    """Returns dict with total_ms (timed region), kernel_ms (mean per-step launch span), launches, clocks."""
    import numpy as np
    import torch
    import aidadsp_lv2_b200 as ax
    from aidadsp_lv2_b200 import signals
    eng = ax.Engine(device=device, max_streams=n_local, max_block=T)
    eng.configure(tile_mode=tile_mode, math_mode=0)
    setup_engine(eng, wl, n_local, first_global, n_global)
    dev = torch.device("cuda", device)
    # input ring larger than L2 so every period's audio comes from HBM (timing hygiene)
    bytes_per_step = 2 * T * n_local * 4
    ring = max(2, min(steps + warmup, (2 * L2_BYTES + bytes_per_step - 1) // bytes_per_step))
    base = torch.from_numpy(np.ascontiguousarray(signals.sweep_noise_block(min(n_local, 256), T).T)).to(dev)  # [T][<=256]
    reps = (n_local + base.shape[1] - 1) // base.shape[1]
    xin = [(base.repeat(1, reps)[:, :n_local] * (1.0 - 0.001 * i)).contiguous() for i in range(ring)]
    yout = [torch.empty_like(xin[0]) for _ in range(ring)]
    p1 = p2 = None
    if wl.automate_params:
        p1 = signals.triangle(1.0, steps + warmup + 1, T)
        p2 = signals.triangle(1.7, steps + warmup + 1, T)
    ext = torch.cuda.ExternalStream(eng.cuda_stream, device=dev)

    def step(i):
        if p1 is not None:
            ph = np.arange(n_local, dtype=np.float32) * 0.01
            eng.set_params(0, np.mod(p1[i] + ph, 1.0).astype(np.float32), np.mod(p2[i] + ph, 1.0).astype(np.float32))
        eng.render_device(xin[i % ring].data_ptr(), yout[i % ring].data_ptr(), n_local, T, layout=ax.LAYOUT_INTERLEAVED)

    for i in range(warmup):
        step(i)
    eng.synchronize()
    torch.cuda.synchronize(dev)
    return eng, ext, step, (xin, yout)


def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    import aidadsp_lv2_b200 as ax
    from aidadsp_lv2_b200 import workloads, sharding, engine as engmod

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = workloads.get(args.workload)
    n_local = wl.n_streams            # weak scaling: the named configuration per GPU
    n_global = n_local * world
    first_global = n_local * rank
    steps, warmup = args.steps, max(args.warmup, 3)

    sampler = ClockSampler(local_rank)
    sampler.start()
    # ------------------------------------------------------------ device-resident arm (value)
    eng, ext, step, keep = device_bench(wl, n_local, first_global, n_global, steps, warmup, local_rank, args.tile_mode)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    launches0 = eng.launch_count
    e0.record(ext)
    for i in range(steps):
        ev[i][0].record(ext)
        step(warmup + i)
        ev[i][1].record(ext)
    e1.record(ext)
    e1.synchronize()
    torch.cuda.synchronize(dev)
    total_ms = e0.elapsed_time(e1)
    launches = eng.launch_count - launches0
    span_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    if world > 1:
        dist.barrier()
        t = torch.tensor([total_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / steps
    value = n_global * T / (ms_per_step * 1e-3)

    # ------------------------------------------------------------ end-to-end arm (plugin + batch host, HOST buffers)
    tmp = tempfile.mkdtemp(prefix="aidax_bench_")
    paths = write_models(wl, tmp)
    host = ax.BatchHost(ax.plugin_path(), n_instances=n_local, max_block=T, n_threads=args.host_threads, device=local_rank)
    for port in range(4, 25):
        if port != 23:
            host.set_control(port, float(wl.controls[port - 4]))
    for s in range(n_local):
        host.load_model(paths[wl.model_of_stream(first_global + s, n_global)], instance=s)
    host.input[:, :T] = np.tile(keep[0][0].T.cpu().numpy()[: min(n_local, 256)], ((n_local + 255) // 256, 1))[:n_local]
    for i in range(warmup + 2):
        host.run(T)
    assert host.get_control("MODELINSIZE") == wl.models[wl.model_of_stream(first_global, n_global)].input_size
    e2e_launch0 = host.launch_count
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        host.run(T)
    e2e_s = time.perf_counter() - t0
    e2e_launches = host.launch_count - e2e_launch0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    clocks = sampler.stop()
    e2e_value = n_global * T * steps / e2e_s
    assert host.batched

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ------------------------------------------------------------ roofline of the dominant kernel (FP32 FMA pipe)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    fp32_peak = engmod.measure_fp32_peak(local_rank, 200)
    flop_per_launch = sum(wl.flops_per_sample(first_global + s) for s in range(0, n_local, max(1, n_local // 64))) / len(range(0, n_local, max(1, n_local // 64))) * n_local * T
    launches_per_step = max(1, launches // steps)
    kernel_ms = span_ms  # CUDA-event span of one period's launches on the engine stream
    achieved_tflops = flop_per_launch / (kernel_ms * 1e-3) / 1e12
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    io_gbs = (2 * 4 * T * n_local) / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "fp32", "kernel": "chain_kernel (fused chain + register-tiled LSTM/GRU recurrence)",
                "achieved": round(achieved_tflops, 3), "peak": round(fp32_peak, 2), "unit": "TFLOP/s", "frac": round(achieved_tflops / fp32_peak, 4),
                "peak_source": "FFMA loop measured in this run (aidax_measure_fp32_peak); FP32 CUDA-core peak is not in MEASURED_PEAKS.json",
                "flop_per_stream_sample": round(flop_per_launch / (n_local * T), 1), "launch_ms": round(kernel_ms, 5),
                "launches_per_step": launches_per_step, "traffic": None,
                "hbm": {"achieved": round(io_gbs, 2), "peak": hbm_peak, "unit": "GB/s", "frac": round(io_gbs / hbm_peak, 5),
                        "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback 6650 GB/s", "algorithmic_bytes_per_stream_sample": 8}}

    out = {"metric": METRIC, "value": round(value, 1), "unit": "samples/s", "n_gpus": world, "steps": steps, "warmup": warmup,
           "ms_per_step": round(ms_per_step, 6), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32 (f64 biquad state)", "data": "synthetic",
           "config": {"workload": "%s: %s" % (wl.name.upper(), wl.description), "streams_per_gpu": n_local, "streams_total": n_global,
                      "buffer": T, "sample_rate": FS, "layout": "stream-interleaved [128][N] in HBM",
                      "l2": "input/output ring of %d buffer pairs (%.0f MiB) > L2, fresh buffer every step" % (len(keep[0]), len(keep[0]) * 2 * T * n_local * 4 / 2**20),
                      "tile_mode": args.tile_mode, "parallelism": "streams sharded over %d GPU(s), no collective" % world},
           "realtime_streams": int(value // FS), "us_per_buffer": round(ms_per_step * 1e3, 2),
           "e2e": {"value": round(e2e_value, 1), "unit": "samples/s", "h2d_bytes_per_step": 4 * T * n_local, "d2h_bytes_per_step": 4 * T * n_local,
                   "us_per_buffer": round(e2e_s / steps * 1e6, 2), "path": "aidax-batch-host -> rt-neural-generic.so run() x N -> #batch end() -> aidax_render_host (pinned arena)",
                   "host_threads": args.host_threads, "timer": "wall clock (call returns after D2H + stream synchronise)"},
           "gpu_launches": int(launches), "e2e_gpu_launches": int(e2e_launches), "clocks": clocks, "roofline": roofline}

    # ------------------------------------------------------------ capacity (same chain, GPU-filling stream count)
    if world == 1 and not args.no_capacity and wl.name == "c2":
        host.close()
        eng.close()
        del keep
        torch.cuda.empty_cache()
        cw = workloads.get("capacity", args.capacity_streams)
        ceng, cext, cstep, ckeep = device_bench(cw, cw.n_streams, 0, cw.n_streams, 20, 3, local_rank, 0)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(cext)
        for i in range(20):
            cstep(3 + i)
        c1.record(cext)
        c1.synchronize()
        cms = c0.elapsed_time(c1) / 20
        csps = cw.n_streams * T / (cms * 1e-3)
        ctf = csps * cw.flops_per_sample() / 1e12
        out["capacity"] = {"streams": cw.n_streams, "us_per_buffer": round(cms * 1e3, 1), "samples_per_s": round(csps, 1),
                           "realtime_streams_per_gpu": int(csps // FS), "meets_deadline": bool(cms * 1e-3 <= T / FS),
                           "fp32_tflops": round(ctf, 2), "fp32_frac": round(ctf / fp32_peak, 4)}
        ceng.close()

    # ------------------------------------------------------------ CPU baseline (reference binary, host cores)
    if world == 1 and not args.no_cpu_baseline:
        try:
            out["cpu_baseline"] = run_reference_sample(wl, min(steps, 100), 3, paths)
        except Exception as ex:  # the reference build is test infrastructure; its absence must not fail the GPU arm
            out["cpu_baseline"] = {"unavailable": str(ex)}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_reference_sample(wl, steps, warmup, paths, n_streams=None):
    """Times oracle/_ref/rt-neural-generic.so (the unmodified reference, xsimd SSE2 default build) through the same
    batch host, one instance per stream, all host cores.  Returns the cpu_baseline object."""
    import numpy as np
    import aidadsp_lv2_b200 as ax
    from aidadsp_lv2_b200 import signals
    ref = os.path.join(ROOT, "oracle", "_ref", "rt-neural-generic.so")
    if not os.path.exists(ref):
        raise RuntimeError("oracle/_ref/rt-neural-generic.so not built (oracle/build_ref.sh needs /root/reference)")
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    n = n_streams or wl.n_streams
    with Silence():
        host = ax.BatchHost(ref, n_instances=n, max_block=T, n_threads=cores)
        for port in range(4, 25):
            if port != 23:
                host.set_control(port, float(wl.controls[port - 4]))
        for s in range(n):
            host.load_model(paths[wl.model_of_stream(s, n)], instance=s)
        host.input[:, :T] = np.tile(signals.sweep_noise_block(min(n, 64), T), ((n + 63) // 64, 1))[:n]
        for _ in range(warmup + 2):
            host.run(T)
        t0 = time.perf_counter()
        for _ in range(steps):
            host.run(T)
        dt = time.perf_counter() - t0
        insz = host.get_control("MODELINSIZE")
        host.close()
    assert insz >= 1.0
    v = n * T * steps / dt
    return {"value": round(v, 1), "unit": "samples/s", "cores": cores, "kind": "reference",
            "sample": "%d periods of %d streams x %d samples (%s), unmodified rt-neural-generic.so (xsimd SSE2) via aidax-batch-host, %d threads"
                      % (steps, n, T, wl.name.upper(), cores),
            "us_per_buffer": round(dt / steps * 1e6, 1), "realtime_streams": int(v // FS), "seconds": round(dt, 2)}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores."""
    if rank != 0:
        return
    from aidadsp_lv2_b200 import workloads
    wl = workloads.get(args.workload)
    tmp = tempfile.mkdtemp(prefix="aidax_bench_ref_")
    paths = write_models(wl, tmp)
    steps, warmup = args.steps, max(args.warmup, 3)
    # bounded sample: cap the run at roughly 2 minutes of wall clock on the host cores
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    per_step_s = wl.n_streams * 45e-6 / max(cores, 1)
    if per_step_s * (steps + warmup) > 120.0:
        steps = max(5, int(120.0 / per_step_s) - warmup)
    cb = run_reference_sample(wl, steps, warmup, paths)
    ms = cb["seconds"] / steps * 1e3
    out = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "samples/s", "n_gpus": world, "steps": steps, "warmup": warmup,
           "ms_per_step": round(ms, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (f64 biquad state)",
           "data": "synthetic",
           "config": {"workload": "%s: %s" % (wl.name.upper(), wl.description), "streams_per_gpu": wl.n_streams, "streams_total": wl.n_streams,
                      "buffer": T, "sample_rate": FS, "note": "CPU arm: host cores only, no GPU; runs once on rank 0 regardless of --gpus"},
           "realtime_streams": cb["realtime_streams"], "us_per_buffer": cb["us_per_buffer"], "cpu_baseline": cb,
           "e2e": {"value": cb["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5", "capacity"])
    ap.add_argument("--tile-mode", type=int, default=0)
    ap.add_argument("--host-threads", type=int, default=1)
    ap.add_argument("--capacity-streams", type=int, default=131072)
    ap.add_argument("--no-capacity", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
