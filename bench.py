#!/usr/bin/env python3
"""bench.py -- the hot-path benchmark (contract: one JSON line on stdout from rank 0).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--workload c2|c3|c4|c5] [--streams S] [--tile-mode M]

A step = ONE audio period: every stream of the workload renders one 128-sample buffer through the full
rt-neural-generic chain (in_lpf -> pre-gain -> EQ -> LSTM/GRU+Dense -> dc blocker -> EQ -> master gain).

Workload.  BASELINE.json's metric is "real-time 48 kHz streams per box (LSTM-16); samples/s/GPU; us per 128-sample
buffer".  It is quoted on configs[1]'s model and chain (C2: one LSTM-16 AIDA-X model, fused pre-EQ, stream-interleaved
buffers) -- but "streams per box" is a capacity, so the default run renders that exact model and chain at a stream
count that FILLS a GPU (378 880 = 5 full waves of the tensor-core kernel, period still inside the 2.667 ms deadline);
configs[1]'s literal 1024 streams are measured beside it and reported under "c2_1024" (latency case).
--streams 1024 makes the literal configuration the main line.

  value        whole-job stream-samples per second with the period's input already resident in HBM
               (stream-interleaved [128][N] device buffers), CUDA-event timed on the engine's stream.
  realtime_streams  floor(value / 48000): streams the job sustains in real time (period_ms must stay <= 2.667).
  e2e          the same metric through the reference-facing surface: N LV2 plugin instances driven by the batch-render
               host with HOST (pinned) port buffers -- control decode in every run(), H2D copy, kernels, D2H copy and
               the final synchronise are inside the timed region (wall clock).
  roofline     the dominant kernel against the FP32 FFMA pipe (algorithmic flop per stream-sample),
               peak measured live with an FFMA loop; "pipes" adds the XU (MUFU) and tensor-pipe views that explain it.
  cpu_baseline the UNMODIFIED reference (oracle/_ref/rt-neural-generic.so, xsimd SSE2) on the host cores, bounded sample.

--impl reference times the reference's own CPU implementation on all host cores (rank 0 only).
Multi-GPU (torchrun): streams shard over ranks with no data-path collective (weak scaling: same streams per GPU).
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
CAPACITY_STREAMS = 378880   # 5 x 592 CTAs x 128 streams: whole waves of chain_kernel_tc<LSTM,16> (4 CTAs/SM x 148 SMs)
METRIC = "stream-samples/s through the full rt-neural-generic chain, LSTM-16 (real-time 48 kHz streams = value/48000; us per 128-sample buffer in us_per_buffer)"


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
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw"

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

    def mark(self):
        return len(self.rows)

    def stop(self, lo=0, hi=None):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        rows = self.rows[lo:hi] if (hi or lo) and len(self.rows[lo:hi]) >= 2 else self.rows
        sm = sorted(int(r[0]) for r in rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        pw = [float(r[6]) for r in rows if len(r) >= 7 and r[6].replace(".", "", 1).isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "power_w_max": max(pw) if pw else None}


_JSON_FD = None


def emit_json(out):
    line = (json.dumps(out) + "\n").encode()
    sys.stdout.flush()
    os.write(_JSON_FD if _JSON_FD is not None else 1, line)


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
    """Engine + device-resident interleaved buffers; returns (engine, external stream, step function, buffers)."""
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


def timed_device_steps(eng, ext, step, steps, warmup, dev, world, dist):
    """K steps bracketed by barrier + synchronize, CUDA events on the engine stream, MAX over ranks."""
    import numpy as np
    import torch
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
    return total_ms, span_ms, launches


def kernel_name(wl, n_local, tile_mode):
    m = wl.models[0]
    tc = m.hidden <= 40 and (tile_mode == 3 or (tile_mode == 0 and n_local >= 4096))
    return ("chain_kernel_tc (fused chain + tcgen05 3xTF32 gate GEMM in TMEM, thread-per-stream gate epilogue)" if tc
            else "chain_kernel_w / chain_kernel (fused chain + FP32 recurrence: warp-per-stream latency shape for small batches, register-tiled shapes above)"), tc


def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    import aidadsp_lv2_b200 as ax
    from aidadsp_lv2_b200 import workloads, engine as engmod

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_local = args.streams or (CAPACITY_STREAMS if args.workload == "c2" else None)
    wl = workloads.get(args.workload, n_local)
    n_local = wl.n_streams            # weak scaling: the same stream count per GPU
    n_global = n_local * world
    first_global = n_local * rank
    steps, warmup = args.steps, max(args.warmup, 3)

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    # ------------------------------------------------------------ device-resident arm (value)
    eng, ext, step, keep = device_bench(wl, n_local, first_global, n_global, steps, warmup, local_rank, args.tile_mode)
    mark0 = sampler.mark()
    total_ms, span_ms, launches = timed_device_steps(eng, ext, step, steps, warmup, dev, world, dist)
    mark1 = sampler.mark()
    ms_per_step = total_ms / steps
    value = n_global * T / (ms_per_step * 1e-3)
    ring_pairs = len(keep[0])
    eng.close()
    del keep, step, ext
    torch.cuda.empty_cache()

    # ------------------------------------------------------------ end-to-end arm (plugin + batch host, HOST buffers)
    tmp = tempfile.mkdtemp(prefix="aidax_bench_")
    paths = write_models(wl, tmp)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    host_threads = args.host_threads or max(1, cores // world)
    e2e_steps = max(5, min(steps, args.e2e_steps))
    host = ax.BatchHost(ax.plugin_path(), n_instances=n_local, max_block=T, n_threads=host_threads, device=local_rank)
    for port in range(4, 25):
        if port != 23:
            host.set_control(port, float(wl.controls[port - 4]))
    n_models = len(wl.models)
    if n_models == 1:
        host.load_model(paths[0])
    else:
        for s in range(n_local):
            host.load_model(paths[wl.model_of_stream(first_global + s, n_global)], instance=s)
    from aidadsp_lv2_b200 import signals
    blk = signals.sweep_noise_block(min(n_local, 256), T)
    host.input[:, :T] = np.tile(blk, ((n_local + blk.shape[0] - 1) // blk.shape[0], 1))[:n_local]
    for i in range(5):
        host.run(T)   # model load + 2048-sample pre-warm + fade-in happen here
    assert host.get_control("MODELINSIZE") == wl.models[wl.model_of_stream(first_global, n_global)].input_size
    e2e_launch0 = host.launch_count
    tm0 = host.timing()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        host.run(T)
    e2e_s = time.perf_counter() - t0
    e2e_launches = host.launch_count - e2e_launch0
    tm1 = host.timing()
    e2e_phase_us = [(tm1[i] - tm0[i]) / max(tm1[2] - tm0[2], 1.0) * 1e6 for i in (0, 1)]
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    clocks = sampler.stop(mark0, mark1)
    e2e_value = n_global * T * e2e_steps / e2e_s
    assert host.batched
    assert float(np.abs(host.output[:, :T]).max()) > 1e-4, "e2e output is silent"
    host.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ------------------------------------------------------------ roofline of the dominant kernel
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    fp32_peak = engmod.measure_fp32_peak(local_rank, 200)
    try:
        pcie = engmod.measure_pcie(local_rank, 256 << 20)
    except Exception:
        pcie = None
    sample_ids = range(0, n_local, max(1, n_local // 64))
    flop_per_sample = sum(wl.flops_per_sample(first_global + s) for s in sample_ids) / len(sample_ids)
    flop_per_launch = flop_per_sample * n_local * T
    launches_per_step = max(1, launches // steps)
    kernel_ms = span_ms  # CUDA-event span of one period's launches on the engine stream
    achieved_tflops = flop_per_launch / (kernel_ms * 1e-3) / 1e12
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    io_gbs = (2 * 4 * T * n_local) / (kernel_ms * 1e-3) / 1e9
    kname, is_tc = kernel_name(wl, n_local, args.tile_mode)
    m0 = wl.models[0]
    sm_clock = (clocks.get("sm_mhz") or 1965) * 1e6
    xu_evals = (5 if m0.cell == 0 else 3) * m0.hidden          # transcendental evaluations per stream-sample
    mufu_ops = xu_evals + (xu_evals + 3) // 4 if is_tc else 2 * xu_evals  # ex2 each + one batched rcp per 4 (tc) / ex2+rcp each
    xu_rate = n_local * T * mufu_ops / (kernel_ms * 1e-3)
    pipes = {"xu_mufu": {"achieved_gops": round(xu_rate / 1e9, 1), "peak_gops": round(148 * 16 * sm_clock / 1e9, 1),
                         "frac": round(xu_rate / (148 * 16 * sm_clock), 4), "note": "MUFU ex2/rcp per stream-sample x rate vs 16 lanes/clk/SM at the sampled SM clock (F2F conversions of the f64 biquads share this pipe and are not counted)"}}
    if is_tc:
        kp = (m0.hidden + 4 + 7) // 8 * 8
        tc_flops = 3 * 2 * 4 * m0.hidden * kp * n_local * T / (kernel_ms * 1e-3)
        tf32_peak = peaks.get("bf16_tflops", 1590.0) / 2.0
        pipes["tensor_tf32"] = {"achieved_tflops": round(tc_flops / 1e12, 1), "peak_tflops": round(tf32_peak, 1), "frac": round(tc_flops / 1e12 / tf32_peak, 4),
                                "note": "executed tcgen05 flop (3 TF32 products, K padded to %d) vs half the measured bf16 burst peak (tf32 runs at half the bf16 rate)" % kp}
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tr = json.load(f)
        key = "%s_%d" % (wl.name, n_local)
        if key in tr:
            traffic = tr[key]
    except Exception:
        pass
    roofline = {"bound": "fp32", "kernel": kname,
                "achieved": round(achieved_tflops, 3), "peak": round(fp32_peak, 2), "unit": "TFLOP/s", "frac": round(achieved_tflops / fp32_peak, 4),
                "peak_source": "FFMA loop measured in this run (aidax_measure_fp32_peak); FP32 CUDA-core peak is not in MEASURED_PEAKS.json",
                "flop_per_stream_sample": round(flop_per_sample, 1), "launch_ms": round(kernel_ms, 5),
                "launches_per_step": launches_per_step,
                "traffic": None if traffic is None else traffic["bytes_per_launch"], "traffic_detail": traffic, "pipes": pipes,
                "hbm": {"achieved": round(io_gbs, 2), "peak": hbm_peak, "unit": "GB/s", "frac": round(io_gbs / hbm_peak, 5),
                        "peak_source": "MEASURED_PEAKS.json (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)", "algorithmic_bytes_per_stream_sample": 8}}

    period_ms = T / FS * 1e3
    literal = args.workload == "c2" and n_local == 1024
    desc = wl.description if literal or args.workload != "c2" else \
        "configs[1]'s model and chain (one LSTM-16 model, fused pre-EQ, stream-interleaved 128-sample buffers) at the GPU-filling stream count %d per GPU; configs[1]'s literal 1024 streams under c2_1024" % n_local
    out = {"metric": METRIC, "value": round(value, 1), "unit": "samples/s", "n_gpus": world, "steps": steps, "warmup": warmup,
           "ms_per_step": round(ms_per_step, 6), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32 (tf32x3 tensor-core gate GEMM, f64 biquad state)" if is_tc else "f32 (f64 biquad state)", "data": "synthetic",
           "config": {"workload": "%s: %s" % (wl.name.upper(), desc), "streams_per_gpu": n_local, "streams_total": n_global,
                      "buffer": T, "sample_rate": FS, "layout": "stream-interleaved [128][N] in HBM",
                      "l2": "input/output ring of %d buffer pairs (%.0f MiB) > L2, fresh buffer every step" % (ring_pairs, ring_pairs * 2 * T * n_local * 4 / 2**20),
                      "tile_mode": args.tile_mode, "parallelism": "streams sharded over %d GPU(s), no collective" % world},
           "realtime_streams": int(value // FS), "realtime_streams_per_gpu": int(value // FS // world), "us_per_buffer": round(ms_per_step * 1e3, 2),
           "meets_deadline": bool(ms_per_step <= period_ms),
           "e2e": {"value": round(e2e_value, 1), "unit": "samples/s", "h2d_bytes_per_step": 4 * T * n_local, "d2h_bytes_per_step": 4 * T * n_local,
                   "us_per_buffer": round(e2e_s / e2e_steps * 1e6, 2), "steps": e2e_steps, "realtime_streams": int(e2e_value // FS),
                   "meets_deadline": bool(e2e_s / e2e_steps * 1e3 <= period_ms),
                   "phase_us": {"run_calls": round(e2e_phase_us[0], 1), "batch_end_copies_kernels_sync": round(e2e_phase_us[1], 1)},
                   "path": "aidax-batch-host -> rt-neural-generic.so run() x N (each completed chunk of instances is shipped at once: aidax_render_host_submit = pinned H2D -> kernels -> D2H on separate CUDA streams) -> #batch end() waits",
                   "host_threads": host_threads, "timer": "wall clock (call returns after D2H + stream synchronise)",
                   "pcie": None if pcie is None else {
                       "h2d_gbs": round(pcie[0], 1), "d2h_gbs": round(pcie[1], 1), "duplex_gbs_per_direction": round(pcie[2], 1),
                       "copy_floor_us": round(4 * T * n_local / (pcie[2] * 1e9) * 1e6, 1),
                       "frac_of_copy_floor": round((4 * T * n_local / (pcie[2] * 1e9)) / (e2e_s / e2e_steps), 3),
                       "note": "pinned-memory cudaMemcpyAsync rates measured in this run over 256 MiB; copy_floor = one period's bytes per direction at the duplex rate: the e2e period cannot be shorter"}},
           "gpu_launches": int(launches), "e2e_gpu_launches": int(e2e_launches), "clocks": clocks, "roofline": roofline}

    # ------------------------------------------------------------ configs[1] literal (1024 streams): the latency case
    if world == 1 and args.workload == "c2" and not literal and not args.no_literal:
        lw = workloads.get("c2", 1024)
        leng, lext, lstep, lkeep = device_bench(lw, 1024, 0, 1024, 200, 10, local_rank, 0)
        lms, lspan, _ = timed_device_steps(leng, lext, lstep, 200, 10, dev, 1, None)
        lus = lms / 200 * 1e3
        out["c2_1024"] = {"streams": 1024, "us_per_buffer": round(lus, 2), "samples_per_s": round(1024 * T / (lus * 1e-6), 1),
                          "realtime_margin": round(period_ms * 1e3 / lus, 1), "kernel": kernel_name(lw, 1024, 0)[0]}
        leng.close()
        del lkeep

    # ------------------------------------------------------------ CPU baseline (reference binary, host cores)
    if world == 1 and not args.no_cpu_baseline:
        try:
            out["cpu_baseline"] = run_reference_sample(wl, 300, 3, paths, n_streams=min(n_local, 4096))
        except Exception as ex:  # the reference build is test infrastructure; its absence must not fail the GPU arm
            out["cpu_baseline"] = {"unavailable": str(ex)}
    emit_json(out)
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
            "sample": "%d periods of %d streams x %d samples of the same model and chain (%s), unmodified rt-neural-generic.so (xsimd SSE2) via aidax-batch-host, %d threads"
                      % (steps, n, T, wl.name.upper(), cores),
            "us_per_buffer": round(dt / steps * 1e6, 1), "realtime_streams": int(v // FS), "seconds": round(dt, 2)}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores."""
    if rank != 0:
        return
    from aidadsp_lv2_b200 import workloads
    n_cfg = args.streams or (CAPACITY_STREAMS if args.workload == "c2" else None)
    wl = workloads.get(args.workload, n_cfg)
    tmp = tempfile.mkdtemp(prefix="aidax_bench_ref_")
    paths = write_models(wl, tmp)
    steps, warmup = args.steps, max(args.warmup, 3)
    # bounded sample: the CPU renders min(N, 4096) streams of the workload per step, capped at ~2 minutes of wall clock
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    n = min(wl.n_streams, 4096)
    per_step_s = n * 45e-6 / max(cores, 1)
    if per_step_s * (steps + warmup) > 120.0:
        steps = max(5, int(120.0 / per_step_s) - warmup)
    cb = run_reference_sample(wl, steps, warmup, paths, n_streams=n)
    ms = cb["seconds"] / steps * 1e3
    out = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "samples/s", "n_gpus": world, "steps": steps, "warmup": warmup,
           "ms_per_step": round(ms, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (f64 biquad state)",
           "data": "synthetic",
           "config": {"workload": "%s: %s" % (wl.name.upper(), wl.description), "streams_per_gpu": wl.n_streams, "streams_total": wl.n_streams,
                      "buffer": T, "sample_rate": FS, "sample_streams": n,
                      "note": "CPU arm: host cores only, no GPU; each step renders a bounded sample (%d streams) of the workload; throughput is independent of the stream count on a CPU; runs once on rank 0 regardless of --gpus" % n},
           "realtime_streams": cb["realtime_streams"], "us_per_buffer": cb["us_per_buffer"], "cpu_baseline": cb,
           "e2e": {"value": cb["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit_json(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--streams", type=int, default=0, help="streams per GPU (default: c2 -> %d, others -> the BASELINE.json count)" % CAPACITY_STREAMS)
    ap.add_argument("--tile-mode", type=int, default=0)
    ap.add_argument("--host-threads", type=int, default=0, help="batch-host threads per rank (default: host cores / ranks)")
    ap.add_argument("--e2e-steps", type=int, default=30)
    ap.add_argument("--no-literal", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE line: rank 0's JSON.  Native libraries write there too (NCCL prints its version banner
    # on communicator creation, the reference prints layer info while parsing models), so file descriptor 1 is pointed
    # at stderr for the whole run and the JSON line goes to the saved descriptor.
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
