#!/usr/bin/env python3
"""bench.py -- the hot-path benchmark (contract: one JSON line on stdout from rank 0).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--workload c2|c3|c4|c5] [--streams S] [--scaling weak|strong] [--tile-mode M]

A step = ONE audio period: every stream of the workload renders one 128-sample buffer through the full
rt-neural-generic chain (in_lpf -> pre-gain -> EQ -> LSTM/GRU+Dense -> dc blocker -> EQ -> master gain).

Workload.  BASELINE.json's metric is "real-time 48 kHz streams per box (LSTM-16); samples/s/GPU; us per 128-sample
buffer".  It is quoted on configs[1]'s model and chain (C2: one LSTM-16 AIDA-X model, fused pre-EQ, stream-interleaved
buffers) -- but "streams per box" is a capacity, so the default run renders that exact model and chain at a stream
count that FILLS a GPU (whole waves of the tensor-core kernel, period still inside the 2.667 ms deadline);
configs[1]'s literal 1024 streams are measured beside it and reported under "c2_1024" (latency case).
--streams 1024 makes the literal configuration the main line.

  value        whole-job stream-samples per second with the period's input already resident in HBM
               (stream-interleaved [128][N] device buffers), CUDA-event timed on the engine's stream.
  realtime_streams  floor(value / 48000): streams the job sustains in real time (period_ms must stay <= 2.667).
  e2e          the same metric through the reference-facing surface: LV2 plugin instances driven by the batch-render
               host with HOST (pinned) port buffers -- the control scan of every instance, H2D copy, kernels, D2H copy and
               the final synchronise are inside the timed region (wall clock).  Host-fed audio costs 512 B per stream and
               period in each direction, so this arm is bound by PCIe, not by the GPU: it is measured at the LARGEST
               instance count whose period stays inside the 2.667 ms deadline (found by a short search against the
               concurrently measured copy rate of all ranks), and additionally at the north-star share of 125 000
               streams per GPU and at the full instance count.
  roofline     the dominant kernel against MEASURED pipe peaks (csrc/probes.cu): XU (MUFU + F2F), tcgen05 TF32 at the
               kernel's own MMA shape, FP32, FP64, HBM; `bound` names the pipe with the largest fraction.
  cpu_baseline the UNMODIFIED reference (oracle/_ref/rt-neural-generic.so, xsimd SSE2; AVX2 flavour beside it) on the
               host cores, one pinned thread per core, >= 2 s per flavour.

--impl reference times the reference's own CPU implementation on all host cores (rank 0 only).
Multi-GPU (torchrun): streams shard over ranks with no data-path collective.  --scaling weak (default): the same
stream count per GPU; --scaling strong: BASELINE.json's fixed total (configs[3]: 65536, configs[4]: 262144) cut into
contiguous shards (aidadsp_lv2_b200.sharding.shard_range).
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
CAPACITY_STREAMS = 378880   # 5 x 296 CTAs x 256 streams: whole waves of chain_kernel_tc3<LSTM,16> (2 CTAs/SM x 148 SMs)
CHUNK = 16384               # instances per hand-over range of the plugin path (AIDAX_BATCH_CHUNK default)
GRAIN = 4096                # granularity of the end-to-end instance-count search
NORTH_STAR_SHARE = 125000   # 1 M streams per 8-GPU box
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
    # every stream gets its own input (sweep + noise with a per-stream phase and noise sequence), generated on the device
    # (kernel K6, aidax_generate_input): ring buffer i holds samples [i*T, (i+1)*T) of all n_local streams, [T][n_local]
    xin = [torch.empty((T, n_local), dtype=torch.float32, device=dev) for _ in range(ring)]
    for i, x in enumerate(xin):
        eng.generate_input(x.data_ptr(), n_local, T, start=i * T, first=first_global, layout=ax.LAYOUT_INTERLEAVED, seed=2024, device=device)
    eng.synchronize()
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
    return ("chain_kernel_tc3 (fused chain, warp-specialised: recurrence warps with the tcgen05 3xTF32 gate GEMM in TMEM and a thread-per-stream gate epilogue, filter warps one sub-block ahead/behind over a shared-memory ring)" if tc
            else "chain_kernel_w / chain_kernel (fused chain + FP32 recurrence: warp-per-stream latency shape for small batches, register-tiled shapes above)"), tc


def _host_cores():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def e2e_measure(host, n_active, periods, dist, world, dev):
    """`periods` back-to-back host.run(T) calls with n_active instances: (mean seconds, max seconds, launches, phase us).
    Wall clock around each call (it returns after the D2H copy and the stream synchronise); MAX over ranks."""
    import numpy as np
    import torch
    host.set_active(n_active)
    for _ in range(2):
        host.run(T)
    tm0, l0 = host.timing(), host.launch_count
    if world > 1:
        dist.barrier()
    per = np.empty(periods)
    for i in range(periods):
        t0 = time.perf_counter()
        host.run(T)
        per[i] = time.perf_counter() - t0
    tm1 = host.timing()
    launches = host.launch_count - l0
    mean_s, max_s, p95_s = float(per.mean()), float(per.max()), float(np.sort(per)[int(0.95 * (periods - 1))])
    if world > 1:
        t = torch.tensor([mean_s, max_s, p95_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        mean_s, max_s, p95_s = (float(v) for v in t.tolist())
    phase = [(tm1[i] - tm0[i]) / max(tm1[2] - tm0[2], 1.0) * 1e6 for i in (0, 1)]
    return mean_s, max_s, p95_s, launches, phase


def e2e_arm(args, wl, paths, n_inst, first_global, n_global, rank, world, local_rank, dist, dev):
    """The plugin path: n_inst LV2 instances of the workload in one batch host.  Returns the e2e object of the JSON line."""
    import numpy as np
    import torch
    import aidadsp_lv2_b200 as ax
    from aidadsp_lv2_b200 import signals, engine as engmod
    period_s = T / FS
    host_threads = args.host_threads or max(1, _host_cores() // world)
    # the copy ceiling with every rank copying at once: what the box's PCIe / host memory gives this GPU under load
    pcie = None
    try:
        if world > 1:
            dist.barrier()
        pcie = engmod.measure_pcie(local_rank, 128 << 20)
        if world > 1:
            t = torch.tensor([pcie[0], pcie[1], pcie[2]], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            pcie = tuple(float(v) for v in t.tolist())
    except Exception:
        pcie = None
    if world > 1:   # every rank must take the same decisions below (the measurements inside them are collectives)
        ok = torch.tensor([1 if pcie else 0], device=dev, dtype=torch.int64)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            pcie = None

    # The hand-over granularity follows the copy rate: a chunk should cost ~0.17 ms of copy time in each direction (16384
    # instances at the 49 GB/s of a lone GPU; on the 8-GPU box of this pool every GPU gets ~12 GB/s when all ranks copy at once,
    # and a 16384-instance chunk would be a quarter of the period).  Set before the plugin's engine is created; an explicit
    # AIDAX_BATCH_CHUNK wins.
    chunk = CHUNK
    if pcie and "AIDAX_BATCH_CHUNK" not in os.environ:
        chunk = max(2048, min(CHUNK, int(round(pcie[2] * 340 / 2048)) * 2048))
        os.environ["AIDAX_BATCH_CHUNK"] = str(chunk)
    elif "AIDAX_BATCH_CHUNK" in os.environ:
        chunk = int(os.environ["AIDAX_BATCH_CHUNK"])
    host = ax.BatchHost(ax.plugin_path(), n_instances=n_inst, max_block=T, n_threads=host_threads, device=local_rank)
    for port in range(4, 25):
        if port != 23:
            host.set_control(port, float(wl.controls[port - 4]))
    if len(wl.models) == 1:
        host.load_model(paths[0])
    else:
        for s in range(n_inst):
            host.load_model(paths[wl.model_of_stream(first_global + s, n_global)], instance=s)
    blk = signals.sweep_noise_block(min(n_inst, 256), T)
    host.input[:, :T] = np.tile(blk, ((n_inst + blk.shape[0] - 1) // blk.shape[0], 1))[:n_inst]
    for i in range(5):
        host.run(T)   # model load + 2048-sample pre-warm + fade-in happen here
    assert host.batched
    assert host.get_control("MODELINSIZE") == wl.models[wl.model_of_stream(first_global, n_global)].input_size

    points = {}

    def point(n, periods):
        mean_s, max_s, p95_s, launches, phase = e2e_measure(host, n, periods, dist, world, dev)
        points[n] = {"instances_per_gpu": n, "us_per_buffer": round(mean_s * 1e6, 1), "us_p95": round(p95_s * 1e6, 1), "us_max": round(max_s * 1e6, 1),
                     "samples_per_s": round(n * world * T / mean_s, 1), "meets_deadline": bool(mean_s <= period_s), "periods": periods, "launches": launches,
                     "phase_us": {"control_scan": round(phase[0], 1), "batch_end_copies_kernels_sync": round(phase[1], 1)}}
        return mean_s

    # ---- the largest instance count whose period meets the deadline
    if args.e2e_streams:
        n_rt = min(n_inst, args.e2e_streams)
        point(n_rt, args.e2e_steps)
    elif args.scaling == "strong" or n_inst <= GRAIN or args.workload != "c2" or args.streams:
        n_rt = n_inst   # a configuration with a given stream count is measured at that count
        point(n_rt, args.e2e_steps)
    else:
        target = 0.93 * period_s   # mean period the search aims at: headroom for the slowest period
        guess = n_inst
        if pcie:
            guess = int(0.8 * pcie[2] * 1e9 * period_s / (4 * T))
        n = max(GRAIN, min(n_inst, guess) // GRAIN * GRAIN)
        if world > 1:
            t = torch.tensor([n], device=dev, dtype=torch.int64)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            n = int(t.item())
        best = None
        for _ in range(5):
            if n in points:
                break
            mean_s = point(n, 10)
            if mean_s <= period_s and (best is None or n > best):
                best = n
            scale = target / mean_s
            nxt = max(GRAIN, min(n_inst, int(n * scale)) // GRAIN * GRAIN)
            if world > 1:   # every rank renders the same count
                t = torch.tensor([nxt], device=dev, dtype=torch.int64)
                dist.all_reduce(t, op=dist.ReduceOp.MIN)
                nxt = int(t.item())
            if nxt == n or (mean_s <= period_s and nxt < n):
                break
            n = nxt
        # refine between the largest count that met the deadline and the smallest that did not (the proportional steps above
        # stop at the 4096 grain): bisect, at most three more points
        for _ in range(3):
            passing = [k for k in points if points[k]["meets_deadline"]]
            failing = [k for k in points if not points[k]["meets_deadline"]]
            if not passing or not failing:
                break
            lo, hi = max(passing), min(k for k in failing if k > max(passing)) if any(k > max(passing) for k in failing) else None
            if hi is None or hi - lo <= GRAIN:
                break
            mid = (lo + hi) // 2 // GRAIN * GRAIN
            if world > 1:
                t = torch.tensor([mid], device=dev, dtype=torch.int64)
                dist.all_reduce(t, op=dist.ReduceOp.MIN)
                mid = int(t.item())
            if mid in points or mid <= lo:
                break
            if point(mid, 10) <= period_s:
                best = mid
        n_rt = best if best is not None else min(points, key=lambda k: points[k]["us_per_buffer"])
        point(n_rt, args.e2e_steps)   # the reported measurement
        # real time means (nearly) every period: the reported count must also keep the 95th percentile inside the deadline
        while (not points[n_rt]["meets_deadline"] or points[n_rt]["us_p95"] > period_s * 1e6) and n_rt > GRAIN:
            n_rt -= 2 * GRAIN if n_rt > 8 * GRAIN else GRAIN
            point(n_rt, args.e2e_steps)
    main = points[n_rt]
    # ---- the north-star share (1 M streams per 8-GPU box) and the full instance count, for the record
    extra = {}
    if not args.e2e_streams and args.scaling != "strong" and args.workload == "c2" and not args.streams:
        if n_inst >= NORTH_STAR_SHARE and NORTH_STAR_SHARE not in points:
            point(NORTH_STAR_SHARE, 10)
        if NORTH_STAR_SHARE in points:
            extra["north_star_share_125000"] = points[NORTH_STAR_SHARE]
        if n_inst not in points:
            point(n_inst, 6)
        extra["all_instances"] = points[n_inst]
    host.set_active(n_inst)
    assert float(np.abs(host.output[:n_rt, :T]).max()) > 1e-4, "e2e output is silent"
    bulk, cpus = host.bulk, host.cpus
    host.close()

    mean_s = main["us_per_buffer"] * 1e-6
    e2e = {"value": main["samples_per_s"], "unit": "samples/s", "h2d_bytes_per_step": 4 * T * n_rt, "d2h_bytes_per_step": 4 * T * n_rt,
           "streams_per_gpu": n_rt, "streams_total": n_rt * world, "us_per_buffer": main["us_per_buffer"], "us_p95": main["us_p95"], "us_max": main["us_max"],
           "steps": main["periods"], "realtime_streams": int(n_rt * world) if main["meets_deadline"] else int(main["samples_per_s"] // FS),
           "meets_deadline": main["meets_deadline"], "deadline_us": round(period_s * 1e6, 1), "phase_us": main["phase_us"],
           "selection": "largest multiple of %d instances per GPU whose mean AND 95th-percentile period stay inside the deadline (search: %s)" %
                        (GRAIN, ", ".join("%d -> %.0f us" % (k, points[k]["us_per_buffer"]) for k in sorted(points))),
           "path": "aidax-batch-host -> #batch begin() -> run_block() over the instances (control-row scan; run() only for instances with events or changed "
                   "controls) -> each completed chunk of instances is shipped at once (aidax_render_host_submit: pinned H2D -> kernels -> D2H on separate "
                   "CUDA streams) -> end() waits" if bulk else "aidax-batch-host -> run() x N -> chunks shipped as they complete -> end() waits",
           "bulk_run_block": bool(bulk), "chunk_instances": chunk, "host_threads": host_threads, "host_cpus": cpus, "timer": "wall clock per period (the call returns after D2H + stream synchronise); mean over steps, max over ranks",
           "gpu_launches": main["launches"]}
    e2e.update(extra)
    if pcie:
        floor_s = 4 * T * n_rt / (pcie[2] * 1e9)
        e2e["pcie"] = {"h2d_gbs": round(pcie[0], 1), "d2h_gbs": round(pcie[1], 1), "duplex_gbs_per_direction": round(pcie[2], 1),
                       "ranks_copying_at_once": world, "copy_floor_us": round(floor_s * 1e6, 1), "frac_of_copy_floor": round(floor_s / mean_s, 3),
                       "deadline_bound_streams_per_gpu": int(pcie[2] * 1e9 * period_s / (4 * T)),
                       "note": "pinned-memory cudaMemcpyAsync rates over 128 MiB measured by ALL ranks at the same time (MIN over ranks); copy_floor = one period's "
                               "bytes per direction at the duplex rate; deadline_bound = instances whose audio alone fills a 2.667 ms period at that rate"}
    return e2e


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
    n_cfg = args.streams or (CAPACITY_STREAMS if args.workload == "c2" else None)
    wl = workloads.get(args.workload, n_cfg)
    strong = args.scaling == "strong"
    if strong:
        # BASELINE.json configs[3]/[4]: a FIXED total sharded over the GPUs in contiguous blocks
        n_global = wl.n_streams
        first_global, hi = sharding.shard_range(n_global, rank, world)
        n_local = hi - first_global
    else:
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
    n_inst = n_local if (strong or args.workload != "c2" or args.streams) else min(n_local, args.e2e_instances)
    e2e = e2e_arm(args, wl, paths, n_inst, first_global, n_global, rank, world, local_rank, dist, dev)
    clocks = sampler.stop(mark0, mark1)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ------------------------------------------------------------ roofline of the dominant kernel, against measured pipe peaks
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    pk = engmod.measure_pipe_peaks(local_rank, 700)
    sample_ids = range(0, n_local, max(1, n_local // 64))
    flop_per_sample = sum(wl.flops_per_sample(first_global + s) for s in sample_ids) / len(sample_ids)
    launches_per_step = max(1, launches // steps)
    kernel_ms = span_ms  # CUDA-event span of one period's launches on the engine stream
    rate = n_local * T / (kernel_ms * 1e-3)      # stream-samples per second through the kernel(s)
    kname, is_tc = kernel_name(wl, n_local, args.tile_mode)
    m0 = wl.models[0]
    lstm = m0.cell == 0
    n_bq = int(round((workloads.chain_flops(wl.controls) - 11) / 9.0))   # active biquads of the chain
    xu_evals = (5 if lstm else 3) * m0.hidden                      # transcendental evaluations per stream-sample
    mufu_ops = xu_evals + (xu_evals + 3) // 4 if is_tc else 2 * xu_evals   # tc: one ex2 each + one batched rcp per 4; fp32 shapes: ex2 + rcp each
    f2f_ops = 2 * n_bq                                             # float->double and double->float per biquad and sample (f64 state, float between stages)
    fp64_flop = 9 * n_bq
    pipes = {}
    xu_frac = rate * mufu_ops / (pk["mufu_gops"] * 1e9) + rate * f2f_ops / (pk["f2f_gops"] * 1e9)
    pipes["xu"] = {"frac": round(xu_frac, 4), "mufu_ops_per_stream_sample": mufu_ops, "f2f_ops_per_stream_sample": f2f_ops,
                   "achieved_gops": round(rate * (mufu_ops + f2f_ops) / 1e9, 1), "peak_mufu_gops": round(pk["mufu_gops"], 1), "peak_f2f_gops": round(pk["f2f_gops"], 1),
                   "note": "XU pipe time = MUFU ex2/rcp lane-ops / measured MUFU rate + f32<->f64 conversion lane-ops / measured F2F rate (both run on the XU pipe)"}
    fp32_frac = rate * flop_per_sample / (pk["fp32_tflops"] * 1e12)
    pipes["fp32_equivalent"] = {"frac": round(fp32_frac, 4), "achieved_tflops": round(rate * flop_per_sample / 1e12, 2), "peak_tflops": round(pk["fp32_tflops"], 2),
                                "flop_per_stream_sample": round(flop_per_sample, 1),
                                "note": "SURVEY 8(d) algorithmic flop (the work the reference does on its FP32 SIMD units) against the measured FFMA rate; in the tensor-core shape "
                                        "the gate contraction does not run on this pipe, so this is an equivalent-work figure, not a pipe utilisation"}
    pipes["fp64"] = {"frac": round(rate * fp64_flop / (pk["fp64_tflops"] * 1e12 / 2.0), 4), "achieved_tops": round(rate * fp64_flop / 1e12, 3), "peak_tops": round(pk["fp64_tflops"] / 2.0, 2),
                     "note": "%d unfused DMUL/DADD per stream-sample (%d biquads) against the measured DFMA instruction rate" % (fp64_flop, n_bq)}
    if is_tc:
        kp = (m0.hidden + 4 + 7) // 8 * 8
        ncol = 4 * m0.hidden
        tc_flop = 3 * 2 * ncol * kp                                 # executed per stream-sample: 3 TF32 products, K padded
        tf32_peak = pk["tf32_tflops_n64"] if ncol <= 64 else (pk["tf32_tflops_n160"] if ncol <= 160 else pk["tf32_tflops_n256"])
        pipes["tensor_tf32"] = {"frac": round(rate * tc_flop / (tf32_peak * 1e12), 4), "achieved_tflops": round(rate * tc_flop / 1e12, 1), "peak_tflops": round(tf32_peak, 1),
                                "algorithmic_frac": round(rate * 2 * ncol * (m0.hidden + m0.input_size) / (tf32_peak * 1e12), 4),
                                "note": "executed tcgen05 flop (3 TF32 products, K padded to %d, N = %d) against the MEASURED kind::tf32 rate of M=128 MMAs of the nearest probed N "
                                        "(256/160/64; dense tf32 peak at N=256: %.0f TFLOP/s)" % (kp, ncol, pk["tf32_tflops_n256"])}
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    io_gbs = 8 * rate / 1e9
    pipes["hbm"] = {"frac": round(io_gbs / hbm_peak, 5), "achieved": round(io_gbs, 2), "peak": hbm_peak, "unit": "GB/s",
                    "peak_source": "MEASURED_PEAKS.json (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)", "algorithmic_bytes_per_stream_sample": 8}
    cands = {k: v["frac"] for k, v in pipes.items() if k != "fp32_equivalent" or not is_tc}
    bound = max(cands, key=cands.get)
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tr = json.load(f)
        key = "%s_%d" % (wl.name, n_local)
        if key in tr:
            traffic = tr[key]
    except Exception:
        pass
    bp = pipes[bound]
    if bound == "xu":
        ach, peak_v, unit = bp["achieved_gops"], round(bp["achieved_gops"] / max(bp["frac"], 1e-9), 1), "G lane-ops/s"
    elif bound == "tensor_tf32":
        ach, peak_v, unit = bp["achieved_tflops"], bp["peak_tflops"], "TFLOP/s"
    elif bound == "hbm":
        ach, peak_v, unit = bp["achieved"], bp["peak"], "GB/s"
    elif bound == "fp64":
        ach, peak_v, unit = bp["achieved_tops"], bp["peak_tops"], "T lane-ops/s"
    else:
        ach, peak_v, unit = bp["achieved_tflops"], bp["peak_tflops"], "TFLOP/s"
    roofline = {"bound": {"xu": "xu (MUFU + F2F pipe)", "tensor_tf32": "tensor", "hbm": "hbm", "fp64": "fp64", "fp32_equivalent": "fp32"}[bound], "kernel": kname,
                "achieved": ach, "peak": peak_v, "unit": unit, "frac": round(ach / peak_v, 4) if peak_v else None,
                "peak_source": "measured in this run with saturating register/shared-memory resident loops (aidax_measure_pipe_peaks, csrc/probes.cu) at %.0f MHz; HBM from MEASURED_PEAKS.json" % pk["sm_mhz"],
                "launch_ms": round(kernel_ms, 5), "launches_per_step": launches_per_step,
                "traffic": None if traffic is None else traffic["bytes_per_launch"], "traffic_detail": traffic, "pipes": pipes, "measured_peaks": {k: round(v, 2) for k, v in pk.items()},
                "note": "the kernel is latency-bound below every pipe's peak; `bound` names the pipe with the largest busy fraction (DESIGN.md 3.2)"}

    period_ms = T / FS * 1e3
    literal = args.workload == "c2" and n_local == 1024
    desc = wl.description if literal or args.workload != "c2" else \
        "configs[1]'s model and chain (one LSTM-16 model, fused pre-EQ, stream-interleaved 128-sample buffers) at the GPU-filling stream count %d per GPU; configs[1]'s literal 1024 streams under c2_1024" % n_local
    out = {"metric": METRIC, "value": round(value, 1), "unit": "samples/s", "n_gpus": world, "steps": steps, "warmup": warmup,
           "ms_per_step": round(ms_per_step, 6), "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
           "dtype": "f32 (tf32x3 tensor-core gate GEMM, f64 biquad state)" if is_tc else "f32 (f64 biquad state)", "data": "synthetic",
           "config": {"workload": "%s: %s" % (wl.name.upper(), desc), "streams_per_gpu": n_local, "streams_total": n_global,
                      "buffer": T, "sample_rate": FS, "layout": "stream-interleaved [128][N] in HBM", "input": "per-stream sweep + noise generated on the device (aidax_generate_input)",
                      "l2": "input/output ring of %d buffer pairs (%.0f MiB) > L2, fresh buffer every step" % (ring_pairs, ring_pairs * 2 * T * n_local * 4 / 2**20),
                      "tile_mode": args.tile_mode,
                      "parallelism": ("fixed total of %d streams cut into contiguous shards over %d GPU(s), no collective" if strong else "%d streams per GPU, streams sharded over %d GPU(s), no collective") % (n_global if strong else n_local, world)},
           "realtime_streams": int(value // FS), "realtime_streams_per_gpu": int(value // FS // world), "us_per_buffer": round(ms_per_step * 1e3, 2),
           "meets_deadline": bool(ms_per_step <= period_ms), "e2e": e2e,
           "gpu_launches": int(launches), "e2e_gpu_launches": int(e2e["gpu_launches"]), "clocks": clocks, "roofline": roofline}

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
            out["cpu_baseline"] = run_reference_sample(wl, None, 3, paths, n_streams=min(n_local, 4096), seconds=2.5, both_flavours=True)
        except Exception as ex:  # the reference build is test infrastructure; its absence must not fail the GPU arm
            out["cpu_baseline"] = {"unavailable": str(ex)}
    emit_json(out)
    if world > 1:
        dist.destroy_process_group()


def _time_reference_binary(so_path, wl, paths, n, cores, warmup, steps, seconds, periods_per_step=1):
    """One flavour of the reference through the batch host: one instance per stream, one pinned thread per core.
    steps None -> as many periods as fill `seconds`."""
    import numpy as np
    import aidadsp_lv2_b200 as ax
    from aidadsp_lv2_b200 import signals
    with Silence():
        host = ax.BatchHost(so_path, n_instances=n, max_block=T, n_threads=cores)
        for port in range(4, 25):
            if port != 23:
                host.set_control(port, float(wl.controls[port - 4]))
        for s in range(n):
            host.load_model(paths[wl.model_of_stream(s, n)], instance=s)
        host.input[:, :T] = np.tile(signals.sweep_noise_block(min(n, 64), T), ((n + 63) // 64, 1))[:n]
        for _ in range(2):
            host.run(T)   # model load, pre-warm, fade-in
        t0 = time.perf_counter()
        for _ in range(max(warmup, 1)):
            host.run(T)
        est = (time.perf_counter() - t0) / max(warmup, 1)
        if steps is None:
            periods = max(10, int(seconds / max(est, 1e-6)))
        else:
            if seconds and periods_per_step == 0:
                periods_per_step = max(1, int(seconds / max(est * steps, 1e-9) + 0.999))
            periods = steps * max(1, periods_per_step)
        t0 = time.perf_counter()
        for _ in range(periods):
            host.run(T)
        dt = time.perf_counter() - t0
        insz = host.get_control("MODELINSIZE")
        cpus = host.cpus
        host.close()
    assert insz >= 1.0
    return dt, periods, cpus


def run_reference_sample(wl, steps, warmup, paths, n_streams=None, seconds=2.5, both_flavours=False, periods_per_step=1):
    """Times oracle/_ref/rt-neural-generic.so (the unmodified reference, xsimd SSE2 default build; optionally the AVX2+FMA
    flavour beside it) through the same batch host, one instance per stream, one pinned thread per host core
    (BASELINE.md section 3).  Returns the cpu_baseline object."""
    ref = os.path.join(ROOT, "oracle", "_ref", "rt-neural-generic.so")
    if not os.path.exists(ref):
        raise RuntimeError("oracle/_ref/rt-neural-generic.so not built (oracle/build_ref.sh needs /root/reference)")
    cores = _host_cores()
    n = n_streams or wl.n_streams
    dt, periods, cpus = _time_reference_binary(ref, wl, paths, n, cores, warmup, steps, seconds, periods_per_step)
    v = n * T * periods / dt
    us_stream = dt / periods / n * cores * 1e6      # one core's time per 128-sample buffer of one stream
    out = {"value": round(v, 1), "unit": "samples/s", "cores": cores, "kind": "reference",
           "sample": "%d periods of %d streams x %d samples of the same model and chain (%s), unmodified rt-neural-generic.so (xsimd SSE2) via aidax-batch-host, %d pinned threads"
                     % (periods, n, T, wl.name.upper(), cores),
           "us_per_buffer": round(dt / periods * 1e6, 1), "realtime_streams": int(v // FS), "seconds": round(dt, 2), "periods": periods,
           "us_per_buffer_per_stream_one_core": round(us_stream, 2), "x_realtime_per_core": round(T / FS * 1e6 / us_stream, 1), "host_cpus": cpus}
    avx = os.path.join(ROOT, "oracle", "_ref", "rt-neural-generic-avx2.so")
    if both_flavours and os.path.exists(avx):
        # In its own process: the two flavours are the same C++ sources built with different vector widths, and their
        # inline template instantiations (STB_GNU_UNIQUE symbols) would be merged by the dynamic loader if both lived
        # in one address space -- the second one loaded crashes.
        try:
            code = ("import json, sys; sys.path.insert(0, %r); import bench; from aidadsp_lv2_b200 import workloads; "
                    "wl = workloads.get(%r, %d); "
                    "dt, periods, _ = bench._time_reference_binary(%r, wl, %r, %d, %d, %d, None, %r); print('AVX2', json.dumps([dt, periods]))"
                    % (ROOT, wl.name, wl.n_streams, avx, list(paths), n, cores, warmup, seconds))
            res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
            line = [l for l in res.stdout.splitlines() if l.startswith("AVX2 ")]
            if res.returncode != 0 or not line:
                raise RuntimeError("exit %d: %s" % (res.returncode, (res.stderr or "").strip()[-160:]))
            dt2, periods2 = json.loads(line[-1][5:])
            v2 = n * T * periods2 / dt2
            us2 = dt2 / periods2 / n * cores * 1e6
            out["avx2_fma_flavour"] = {"value": round(v2, 1), "unit": "samples/s", "seconds": round(dt2, 2), "periods": periods2,
                                       "us_per_buffer_per_stream_one_core": round(us2, 2), "x_realtime_per_core": round(T / FS * 1e6 / us2, 1),
                                       "note": "same sources built with -mavx2 -mfma (RTNEURAL_USE_AVX), not the reference's default; timed in a process of its own"}
        except Exception as ex:
            out["avx2_fma_flavour"] = {"unavailable": str(ex)[:200]}
    return out


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
    # Bounded sample: the CPU renders min(N, 4096) streams of the workload; a STEP is as many periods of that sample as
    # make the K timed steps last about 2.5 s (a 0.15 s region wandered by +-15 % between runs), capped at ~2 minutes.
    cores = _host_cores()
    n = min(wl.n_streams, 4096)
    per_period_s = n * 45e-6 / max(cores, 1)
    if per_period_s * (steps + warmup) > 120.0:
        steps = max(5, int(120.0 / per_period_s) - warmup)
    cb = run_reference_sample(wl, steps, warmup, paths, n_streams=n, seconds=2.5, both_flavours=False, periods_per_step=0)
    pps = max(1, cb["periods"] // steps)
    ms = cb["seconds"] / steps * 1e3
    out = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "samples/s", "n_gpus": world, "steps": steps, "warmup": warmup,
           "ms_per_step": round(ms, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (f64 biquad state)",
           "data": "synthetic",
           "config": {"workload": "%s: %s" % (wl.name.upper(), wl.description), "streams_per_gpu": wl.n_streams, "streams_total": wl.n_streams,
                      "buffer": T, "sample_rate": FS, "sample_streams": n, "periods_per_step": pps,
                      "note": "CPU arm: host cores only, no GPU; each step renders %d periods of a bounded sample (%d streams) of the workload; throughput is independent of the stream count on a CPU; runs once on rank 0 regardless of --gpus" % (pps, n)},
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
    ap.add_argument("--e2e-instances", type=int, default=278528, help="plugin instances per GPU the end-to-end arm creates (c2 default run): 17 hand-over chunks, above what PCIe can feed in 2.667 ms")
    ap.add_argument("--e2e-streams", type=int, default=0, help="measure the end-to-end arm at exactly this many instances per GPU instead of searching for the largest count that meets the deadline")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="strong: the workload's fixed total (BASELINE.json configs[3], [4]) sharded over the GPUs")
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
