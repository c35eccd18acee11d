"""Scratch (round 2): the ping-pong tensor-core kernel (kernels_tc2.cuh) against generation 1 (kernels_tc.cuh) and the
FP32 precise shape -- correctness first, then timing.
   python tools/tc2_check.py check            every compiled (H, SPLIT, ATMEM) shape, both cells, conditioned inputs, model
                                              binding (pre-warm) inside the run, ragged period lengths; max abs difference
                                              to generation 1 on the same input (SPLIT 1 must be bit-identical: same
                                              arithmetic, only the schedule differs) and to the FP32 precise shape
   python tools/tc2_check.py time [H ...]     us per 128-sample period at GPU-filling stream counts"""
import os, sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import numpy as np
import aidadsp_lv2_b200 as ax
from aidadsp_lv2_b200 import models

# (SPLIT, ATM): ATM bit 0 = A operand in tensor memory, bit 1 = one tile per thread group (no ping-pong)
SHAPES = {8: [(1, 1), (1, 0)], 12: [(1, 1), (1, 0)], 16: [(1, 1), (1, 0), (2, 1), (2, 3)], 20: [(1, 1), (1, 0)],
          24: [(2, 1), (1, 0)], 32: [(2, 1), (4, 1), (2, 3)], 40: [(2, 1), (5, 1), (2, 3), (2, 2)]}


def render(n, hidden, cell, insz, env, tile_mode, math_mode, periods, seed=7):
    for k in ("AIDAX_TC_GEN", "AIDAX_TC2_SPLIT", "AIDAX_TC2_ATMEM"):
        os.environ.pop(k, None)
    os.environ.update(env)
    eng = ax.Engine(max_streams=n, max_block=128)
    eng.configure(tile_mode=tile_mode, math_mode=math_mode)
    m = models.synthetic(cell, hidden, insz, seed=seed, in_skip=1, in_gain_db=-1.0, out_gain_db=0.5)
    mid = eng.register_model(m)
    rng = np.random.default_rng(11)
    ctl = [ax.default_controls(EQPOS=float(i % 2), BASS=2.0, TREBLE=-3.0, PARAM1=0.1 + 0.01 * (i % 50), PARAM2=0.9 - 0.01 * (i % 37),
                               NETBYPASS=1.0 if i % 97 == 5 else 0.0, ENABLED=0.0 if i % 101 == 7 else 1.0) for i in range(n)]
    for i in range(n):
        s = eng.open_stream()
        if i % 3 != 2:
            eng.set_model(s, mid, 0.3, 0.6)       # the others bind after the second period (pre-warm inside a running period)
        eng.set_controls(s, ctl[i], loading=False)
    outs = []
    for b, T in enumerate(periods):
        if b == 2:
            for i in range(2, n, 3):
                eng.set_model(i, mid, 0.2, 0.4)
        x = (rng.random((n, T), dtype=np.float32) - 0.5)
        outs.append(eng.render_host(x))
    launches = eng.launch_count
    eng.close()
    return np.concatenate(outs, axis=1), launches


def check(only=None):
    periods = [128, 128, 128, 100, 28, 128]
    bad = 0
    for hidden, shapes in SHAPES.items():
        for cell in (0, 1):
            if only and (hidden, cell) != only:
                continue
            insz = 3 if cell == 1 else (2 if hidden in (20, 40) else 1)
            n = 4096 + 130     # ragged: the last CTA has a partly empty second tile
            ref1, _ = render(n, hidden, cell, insz, {"AIDAX_TC_GEN": "1"}, 3, 0, periods)
            refp, _ = render(n, hidden, cell, insz, {"AIDAX_TC_GEN": "1"}, 1, 1, periods)
            for (split, atm) in shapes:
                y, launches = render(n, hidden, cell, insz, {"AIDAX_TC_GEN": "2", "AIDAX_TC2_SPLIT": str(split), "AIDAX_TC2_ATMEM": str(atm)}, 3, 0, periods)
                d1 = float(np.abs(y - ref1).max())
                dp = float(np.abs(y - refp).max())
                ok = np.isfinite(y).all() and dp < 5e-5 and (d1 == 0.0 if split == 1 else d1 < 2e-6)
                bad += 0 if ok else 1
                print("H=%2d cell=%d in=%d split=%d atmem=%d: vs gen1 %.3g  vs fp32-precise %.3g  (gen1 vs precise %.3g)  %s" %
                      (hidden, cell, insz, split, atm, d1, dp, float(np.abs(ref1 - refp).max()), "ok" if ok else "MISMATCH"), flush=True)
    print("TC2 CHECK", "FAILED" if bad else "OK", flush=True)
    return bad


def timing(hs):
    import torch
    from quick_bench import run
    for hidden in hs:
        for n in ((227328, 454656) if hidden <= 16 else (75776, 262144)):
            for k in ("AIDAX_TC_GEN", "AIDAX_TC2_SPLIT", "AIDAX_TC2_ATMEM"):
                os.environ.pop(k, None)
            os.environ["AIDAX_TC_GEN"] = "1"
            print("gen1           ", end="")
            run(n, hidden, tile_mode=3, iters=10)
            for (split, atm) in SHAPES[hidden]:
                os.environ.update({"AIDAX_TC_GEN": "2", "AIDAX_TC2_SPLIT": str(split), "AIDAX_TC2_ATMEM": str(atm)})
                print("gen2 split=%d atm=%d" % (split, atm), end="")
                run(n, hidden, tile_mode=3, iters=10)


if __name__ == "__main__":
    if sys.argv[1] == "check":
        if len(sys.argv) > 3:
            sys.exit(check((int(sys.argv[2]), int(sys.argv[3]))))
        # one process per (hidden size, cell) with its own time limit: a hang or a fault in one shape must not hide the others
        import subprocess
        bad = 0
        for hidden in SHAPES:
            for cell in (0, 1):
                try:
                    r = subprocess.run([sys.executable, "tools/tc2_check.py", "check", str(hidden), str(cell)], capture_output=True, text=True, timeout=150)
                    lines = [l for l in r.stdout.splitlines() if l.startswith("H=")]
                    print("\n".join(lines) if lines else "H=%2d cell=%d: no result, exit %d: %s" % (hidden, cell, r.returncode, (r.stderr or "").strip()[-200:]), flush=True)
                    bad += 0 if r.returncode == 0 and lines else 1
                except subprocess.TimeoutExpired as ex:
                    out = (ex.stdout or b"").decode() if isinstance(ex.stdout, bytes) else (ex.stdout or "")
                    print("\n".join(l for l in out.splitlines() if l.startswith("H=")), flush=True)
                    print("H=%2d cell=%d: TIMEOUT (hang) after the lines above" % (hidden, cell), flush=True)
                    bad += 1
        print("TC2 CHECK", "FAILED" if bad else "OK", flush=True)
        sys.exit(bad)
    timing([int(v) for v in sys.argv[2:]] or [16, 40, 32, 24, 20, 12, 8])
