"""GPU: the launch planner's choices, read back through aidax_engine_describe_plan.  The thresholds are
measured crossovers (profiles/r02aa_small_families.txt, r02w_tc3_one_tile.txt, r02ab_planner_mixed_periods.txt); this
pins them so that a change shows up as a failing test rather than as a slower bench line."""
import pytest

import aidadsp_lv2_b200 as ax
from aidadsp_lv2_b200 import models

pytestmark = pytest.mark.gpu


def _plan(families):
    """families: list of (cell, hidden, n_streams) -> {(cell name, H): plan line}"""
    total = sum(n for _, _, n in families)
    eng = ax.Engine(max_streams=total, max_block=128)
    ctl = ax.default_controls()
    for cell, hidden, n in families:
        mid = eng.register_model(models.synthetic(cell, hidden, 1, seed=hidden + cell))
        for _ in range(n):
            s = eng.open_stream()
            eng.set_model(s, mid)
            eng.set_controls(s, ctl, loading=False)
    plan = eng.describe_plan()
    eng.close()
    assert sum(p["streams"] for p in plan) == total
    return {(p["cell"], p["H"]): p for p in plan}


def test_single_family_shapes(native_libs):
    assert _plan([(0, 16, 1024)])[("lstm", 16)]["shape"] == "warp"            # configs[1]'s literal case: warp-per-stream latency shape
    p = _plan([(0, 16, 8192)])[("lstm", 16)]
    assert (p["shape"], p["streams_per_cta"], p["ctas"]) == ("tc3", 128, 64)    # underfilled GPU: one-tile CTAs
    p = _plan([(0, 16, 40000)])[("lstm", 16)]
    assert (p["shape"], p["streams_per_cta"], p["threads"]) == ("tc3", 256, 512)  # two tiles, 8 recurrence + 8 filter warps
    assert _plan([(0, 24, 1024)])[("lstm", 24)]["shape"] == "fp32"
    assert _plan([(0, 24, 4096)])[("lstm", 24)]["shape"] == "tc3"
    p = _plan([(0, 24, 60000)])[("lstm", 24)]
    assert (p["shape"], p["streams_per_cta"], p["threads"]) == ("tc3", 384, 768)  # three tiles per CTA
    assert _plan([(1, 20, 512)])[("gru", 20)]["shape"] == "tc3"
    assert _plan([(0, 40, 1024)])[("lstm", 40)]["shape"] == "fp32"
    assert _plan([(0, 40, 2048)])[("lstm", 40)]["shape"] == "tc3"
    assert _plan([(0, 64, 8192)])[("lstm", 64)]["shape"] == "fp32"             # H 64/80: no tensor-core shape


def test_a_period_that_fills_the_gpu_takes_the_tensor_core_shape_for_every_family(native_libs):
    fams = [(c, h, 1200) for c in (0, 1) for h in (8, 12, 16, 20, 24, 32, 40)]
    plan = _plan(fams)                                                        # 16 800 streams: more than 64 per SM
    for (cell, h), p in plan.items():
        assert (p["shape"], p["streams_per_cta"]) == ("tc3", 128), (cell, h, p)
    small = _plan([(c, h, 300) for c in (0, 1) for h in (8, 16, 24, 40)])     # 2 400 streams: latency shapes
    for (cell, h), p in small.items():
        assert p["shape"] == ("warp" if h <= 16 else "fp32"), (cell, h, p)
