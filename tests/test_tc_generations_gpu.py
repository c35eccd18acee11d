"""GPU: the warp-specialised tensor-core kernel (kernels_tc3.cuh, the engine's default for families of >= 4096 streams)
against the first-generation tensor-core kernel (kernels_tc.cuh, AIDAX_TC_GEN=1).

Same arithmetic, other schedule: recurrence warps and filter warps are decoupled through a shared-memory ring, the A
operand lives in tensor memory, the five EQ bands run as five passes.  Where a stream's hidden units stay in one thread
(H <= 24) the output must be BIT-IDENTICAL; where they are split over two warps (H = 32, 40) only the order of the Dense
partial sums differs (<= 2e-6).  The scenario covers both cells, conditioned inputs, EQ before and after the network,
bypassed and disabled streams, a model bound in the middle of the run (device-side pre-warm) and ragged period lengths;
the oracle comparison of the same shapes is tests/test_engine_gpu.py (tile_mode 3)."""
import os

import numpy as np
import pytest

import aidadsp_lv2_b200 as ax
from aidadsp_lv2_b200 import models

pytestmark = pytest.mark.gpu

PERIODS = [128, 128, 128, 100, 28, 128]


def _render(n, hidden, cell, insz, gen, ts_max=None, tiles=None):
    saved = {k: os.environ.get(k) for k in ("AIDAX_TC_GEN", "AIDAX_TS_MAX", "AIDAX_TC3_TILES")}
    os.environ["AIDAX_TC_GEN"] = str(gen)
    if ts_max:
        os.environ["AIDAX_TS_MAX"] = str(ts_max)
    if tiles:
        os.environ["AIDAX_TC3_TILES"] = str(tiles)
    try:
        eng = ax.Engine(max_streams=n, max_block=128)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    eng.configure(tile_mode=3, math_mode=0)
    mid = eng.register_model(models.synthetic(cell, hidden, insz, seed=7, in_skip=1, in_gain_db=-1.0, out_gain_db=0.5))
    rng = np.random.default_rng(11)
    for i in range(n):
        s = eng.open_stream()
        if i % 3 != 2:
            eng.set_model(s, mid, 0.3, 0.6)  # the others bind after the second period
        eng.set_controls(s, ax.default_controls(EQPOS=float(i % 2), BASS=2.0, TREBLE=-3.0, PARAM1=0.1 + 0.01 * (i % 50), PARAM2=0.9 - 0.01 * (i % 37),
                                                NETBYPASS=1.0 if i % 97 == 5 else 0.0, ENABLED=0.0 if i % 101 == 7 else 1.0), loading=False)
    outs = []
    for b, T in enumerate(PERIODS):
        if b == 2:
            for i in range(2, n, 3):
                eng.set_model(i, mid, 0.2, 0.4)
        outs.append(eng.render_host(rng.random((n, T), dtype=np.float32) - 0.5))
    eng.close()
    return np.concatenate(outs, axis=1)


@pytest.mark.parametrize("tiles", [1, 2])  # one tile of 128 streams per CTA (what a small family gets) / two or three tiles
@pytest.mark.parametrize("cell,hidden,insz", [(0, 16, 1), (1, 16, 3), (0, 8, 1), (1, 12, 2), (0, 20, 2), (0, 24, 1), (1, 24, 3), (0, 32, 1), (1, 40, 3), (0, 40, 2)])
def test_warp_specialised_kernel_matches_generation_1(native_libs, cell, hidden, insz, tiles):
    n = 2 * 384 + 130  # ragged: the last CTA has a partly empty tile (CTAs hold 128, 256 or -- H = 24 -- 384 streams)
    ref = _render(n, hidden, cell, insz, 1)
    y = _render(n, hidden, cell, insz, 3, tiles=tiles)
    assert np.isfinite(y).all()
    d = float(np.abs(y - ref).max())
    if hidden <= 24:
        assert d == 0.0, d
    else:
        assert d <= 2e-6, d


@pytest.mark.parametrize("ts_max", [8, 12, 32])
def test_warp_specialised_kernel_does_not_depend_on_the_sub_block_length(native_libs, ts_max):
    """The ring hands sub-blocks of `ts` samples from the filter warps to the recurrence warps; the output must not depend on ts."""
    n = 256 + 40
    a = _render(n, 16, 0, 1, 3, tiles=2)
    b = _render(n, 16, 0, 1, 3, ts_max=ts_max, tiles=1 + ts_max % 3 % 2)
    assert np.array_equal(a, b)
