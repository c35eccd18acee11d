"""GPU: the BASELINE.json configurations at reduced stream counts against the CPU oracle, plus size-independent
properties at full size (stream independence: every stream of a shared-model batch fed the same audio produces
the same output, whatever warp/tile it lands in)."""
import numpy as np
import pytest

import aidadsp_lv2_b200 as ax
from aidadsp_lv2_b200 import signals, workloads
from oracle import pyoracle

pytestmark = pytest.mark.gpu
T = 128


def _setup(eng, wl, n):
    mids = [eng.register_model(m) for m in wl.models]
    for s in range(n):
        sid = eng.open_stream()
        eng.set_model(sid, mids[wl.model_of_stream(s, n)])
        eng.set_controls(sid, wl.controls, loading=False)


def _oracle_stream(wl, s, n):
    o = pyoracle.OracleStream()
    o.set_model(wl.models[wl.model_of_stream(s, n)])
    return o


@pytest.mark.parametrize("name,n,tile_mode", [("c2", 96, 0), ("c2", 1024, 0), ("c3", 256, 0), ("c4", 640, 0), ("c4", 2048, 2), ("c5", 200, 2),
                                              ("c2", 1024, 3), ("c3", 300, 3), ("c4", 2048, 3), ("c5", 200, 3)])
def test_config_vs_oracle_subset(native_libs, name, n, tile_mode):
    wl = workloads.get(name, n)
    eng = ax.Engine(max_streams=n, max_block=T)
    eng.configure(tile_mode=tile_mode)
    _setup(eng, wl, n)
    NB = 16
    x = signals.sweep_noise_block(n, T * NB)
    check = sorted(set(list(range(0, n, max(1, n // 24))) + [n - 1]))
    orcs = {s: _oracle_stream(wl, s, n) for s in check}
    p1 = signals.triangle(1.0, NB, T)
    p2 = signals.triangle(1.7, NB, T)
    worst = 0.0
    for b in range(NB):
        c = wl.controls.copy()
        if wl.automate_params:
            ph = np.arange(n, dtype=np.float32) * 0.01
            a1, a2 = np.mod(p1[b] + ph, 1.0).astype(np.float32), np.mod(p2[b] + ph, 1.0).astype(np.float32)
            eng.set_params(0, a1, a2)
        y = eng.render_host(np.ascontiguousarray(x[:, b * T:(b + 1) * T]))
        for s, o in orcs.items():
            if wl.automate_params:
                c[ax.CTRL["PARAM1"]], c[ax.CTRL["PARAM2"]] = a1[s], a2[s]
            yo = o.run(c, x[s, b * T:(b + 1) * T])
            worst = max(worst, float(np.abs(y[s] - yo).max()))
    assert worst <= 1e-4, worst
    eng.close()


@pytest.mark.parametrize("name,n", [("c2", 1024), ("c5", 16384)])
def test_shared_model_streams_are_independent_and_identical(native_libs, name, n):
    """Property at scale: identical input on every stream -> identical output on every stream (bit for bit),
    in both layouts; and interleaved == planar."""
    import torch
    wl = workloads.get(name, n)
    eng = ax.Engine(max_streams=n, max_block=T)
    _setup(eng, wl, n)
    x1 = signals.sweep_noise(0, T * 4)
    for b in range(4):
        xb = np.tile(x1[b * T:(b + 1) * T], (n, 1))
        y = eng.render_host(xb)
        assert np.array_equal(y, np.tile(y[0], (n, 1)))
    # same engine state on all streams; now one more buffer through the device-resident interleaved path
    xi = torch.from_numpy(np.ascontiguousarray(np.tile(x1[:T], (n, 1)).T)).cuda()
    yi = torch.empty_like(xi)
    eng.render_device(xi.data_ptr(), yi.data_ptr(), n, T, layout=ax.LAYOUT_INTERLEAVED)
    eng.synchronize()
    yi = yi.cpu().numpy()
    assert np.array_equal(yi, np.tile(yi[:, :1], (1, n)))
    assert np.isfinite(yi).all() and np.abs(yi).max() > 1e-4
    eng.close()


def test_partial_render_ranges_and_reopen(native_libs):
    """Rendering sub-ranges of streams and closing/re-opening streams keeps every other stream's state intact."""
    wl = workloads.get("c2", 64)
    eng = ax.Engine(max_streams=64, max_block=T)
    _setup(eng, wl, 64)
    ref = ax.Engine(max_streams=64, max_block=T)
    _setup(ref, wl, 64)
    x = signals.sweep_noise_block(64, T * 6)
    for b in range(6):
        xb = np.ascontiguousarray(x[:, b * T:(b + 1) * T])
        full = ref.render_host(xb)
        lo = eng.render_host(np.ascontiguousarray(xb[:40]), first=0)
        hi = eng.render_host(np.ascontiguousarray(xb[40:]), first=40)
        assert np.abs(np.concatenate([lo, hi]) - full).max() <= 2e-6
    eng.close()
    ref.close()


@pytest.mark.parametrize("name,n,NB", [("c3", 8192, 32), ("c4", 65536, 375), ("c5", 262144, 375), ("c2", 378880, 16)])
def test_full_size_configs(native_libs, name, n, NB):
    """BASELINE.json's literal stream counts (and the bench's GPU-filling C2 count) on one GPU, device-resident
    stream-interleaved buffers.  The oracle cannot render that many streams, so parity at full size is:
      (i) a deterministic subset of streams spanning every model and both ends of the batch against the oracle
          (<= 1e-4, north_star tolerance), and
     (ii) the size-independent property: the batch is built from 256 distinct input signals repeated, so two streams
          with the same model, controls and input must agree bit for bit wherever they sit in the launch.
    C4 and C5 (the configurations BASELINE.json shards over 1/2/4/8 GPUs) run for one second of audio (375 buffers), with
    the recurrent drift of the subset tracked: the error of the last 0.2 s must not exceed 4x that of the first 0.2 s
    (or the 2.5e-5 build-to-build floor of the reference itself)."""
    import torch
    wl = workloads.get(name, n)
    eng = ax.Engine(max_streams=n, max_block=T)
    _setup(eng, wl, n)
    P = 256
    base = signals.sweep_noise_block(P, T * NB)                      # [P][T*NB]; stream s plays base[s % P]
    n_models = len(wl.models)
    per_model = max(1, 24 // n_models)
    check = {0, n - 1, n // 2}
    for m in range(n_models):
        lo, hi = (m * n + n_models - 1) // n_models, ((m + 1) * n + n_models - 1) // n_models
        check.update(int(v) for v in np.linspace(lo, hi - 1, per_model + 1))
    check = sorted(check)
    orcs = {s: _oracle_stream(wl, s, n) for s in check}
    p1 = signals.triangle(1.0, NB, T)
    p2 = signals.triangle(1.7, NB, T)
    reps = (n + P - 1) // P
    worst = 0.0
    per_buf = np.zeros(NB)
    base_dev = torch.from_numpy(base).cuda()                         # [P][T*NB]
    for b in range(NB):
        c = wl.controls.copy()
        if wl.automate_params:
            # per-stream phase repeats with the input signal so that property (ii) still holds
            ph = (np.arange(n) % P).astype(np.float32) * 0.01
            a1, a2 = np.mod(p1[b] + ph, 1.0).astype(np.float32), np.mod(p2[b] + ph, 1.0).astype(np.float32)
            eng.set_params(0, a1, a2)
        xb = base_dev[:, b * T:(b + 1) * T].t().repeat(1, reps)[:, :n].contiguous()  # [T][n]
        yb = torch.empty_like(xb)
        eng.render_device(xb.data_ptr(), yb.data_ptr(), n, T, layout=ax.LAYOUT_INTERLEAVED)
        eng.synchronize()
        assert bool(torch.isfinite(yb).all())
        # (ii) streams P apart inside one model's run are replicas of each other
        for m in range(n_models if b % 25 == 0 or b == NB - 1 else 0):
            lo, hi = (m * n + n_models - 1) // n_models, ((m + 1) * n + n_models - 1) // n_models
            k = (hi - lo) // P
            if k >= 2:
                blk = yb[:, lo:lo + k * P].reshape(T, k, P)
                assert bool((blk == blk[:, :1, :]).all()), "model %d: replicated streams differ" % m
        y = yb[:, check].cpu().numpy()                               # [T][len(check)]
        for j, s in enumerate(check):
            if wl.automate_params:
                c[ax.CTRL["PARAM1"]], c[ax.CTRL["PARAM2"]] = a1[s], a2[s]
            yo = orcs[s].run(c, base[s % P, b * T:(b + 1) * T])
            e = float(np.abs(y[:, j] - yo).max())
            worst = max(worst, e)
            per_buf[b] = max(per_buf[b], e)
    assert worst <= 1e-4, worst
    if NB >= 300:
        k = NB // 5
        assert per_buf[-k:].max() <= max(4 * per_buf[:k].max(), 2.5e-5), (per_buf[:k].max(), per_buf[-k:].max())
    assert eng.launch_count >= NB
    eng.close()


@pytest.mark.parametrize("name,n,tile_mode", [("c4", 8192, 3), ("c5", 4096, 3), ("c4", 8192, 0)])
def test_sharding_does_not_change_a_stream(native_libs, name, n, tile_mode):
    """BASELINE.json configs[3]/[4] are a FIXED total sharded over 1/2/4/8 GPUs (bench.py --scaling strong).  A stream
    must sound the same whichever shard renders it: the total rendered by one engine against the same streams cut into
    2 and 4 contiguous shards (sharding.shard_range), each shard its own engine with stream ids starting at 0.  With the
    kernel shape pinned (tile_mode 3) the result is bit-identical; in auto mode a smaller shard may pick another shape,
    so the bound is the shape-to-shape tolerance."""
    from aidadsp_lv2_b200 import sharding
    wl = workloads.get(name, n)
    NB = 6
    x = signals.sweep_noise_block(n, T * NB)

    def render(lo, hi):
        eng = ax.Engine(max_streams=hi - lo, max_block=T)
        eng.configure(tile_mode=tile_mode)
        mids = [eng.register_model(m) for m in wl.models]
        for s in range(lo, hi):
            sid = eng.open_stream()
            eng.set_model(sid, mids[wl.model_of_stream(s, n)])
            eng.set_controls(sid, wl.controls, loading=False)
        y = np.concatenate([eng.render_host(np.ascontiguousarray(x[lo:hi, b * T:(b + 1) * T])) for b in range(NB)], axis=1)
        eng.close()
        return y

    whole = render(0, n)
    for world in (2, 4):
        parts = [render(*sharding.shard_range(n, r, world)) for r in range(world)]
        y = np.concatenate(parts, axis=0)
        assert y.shape == whole.shape
        if tile_mode == 3:
            assert np.array_equal(y, whole), "world %d: max diff %g" % (world, np.abs(y - whole).max())
        else:
            assert np.abs(y - whole).max() <= 2e-5
