"""Scratch (round 2): the warp-specialised tensor-core kernel (kernels_tc3.cuh, AIDAX_TC_GEN=3) against generation 1.
   python tools/tc3_check.py check [H cell]   both cells, conditioned inputs, model binding (pre-warm) inside the run, ragged
                                              period lengths, EQ before/after the network; SPLIT 1 must be BIT-IDENTICAL
                                              to generation 1 (same arithmetic, other schedule)
   python tools/tc3_check.py time [H ...]     us per 128-sample period at GPU-filling stream counts, generation 1 / 2 / 3"""
import os, subprocess, sys
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import numpy as np
from tc2_check import render

HS = [int(v) for v in os.environ.get("TC3_HS", "8 12 16 20 24 32 40").split()]


def check(only=None):
    periods = [128, 128, 128, 100, 28, 128]
    bad = 0
    for hidden in HS:
        for cell in (0, 1):
            if only and (hidden, cell) != only:
                continue
            insz = 3 if cell == 1 else (2 if hidden in (20, 40) else 1)
            n = 4096 + 130
            ref1, _ = render(n, hidden, cell, insz, {"AIDAX_TC_GEN": "1"}, 3, 0, periods)
            y, launches = render(n, hidden, cell, insz, {"AIDAX_TC_GEN": "3"}, 3, 0, periods)
            d1 = float(np.abs(y - ref1).max())
            split1 = hidden <= 24
            ok = np.isfinite(y).all() and (d1 == 0.0 if split1 else d1 < 2e-6)
            bad += 0 if ok else 1
            print("H=%2d cell=%d in=%d gen3 vs gen1: %.3g  %s" % (hidden, cell, insz, d1, "ok" if ok else "MISMATCH"), flush=True)
    return bad


def timing(hs):
    from quick_bench import run
    for hidden in hs:
        for n in ((378880, 454656) if hidden <= 16 else (75776, 262144)):
            for gen in ("1", "2", "3"):
                for k in ("AIDAX_TC_GEN", "AIDAX_TC2_SPLIT", "AIDAX_TC2_ATMEM"):
                    os.environ.pop(k, None)
                os.environ["AIDAX_TC_GEN"] = gen
                print("gen%s " % gen, end="")
                run(n, hidden, tile_mode=3, iters=10)


def sanitizer_run():
    """A short generation-3 scenario for compute-sanitizer (memcheck / racecheck / synccheck): model binding inside the run,
    ragged periods, conditioned inputs, a partly empty last CTA."""
    for hidden, cell, insz in ((16, 0, 1), (40, 1, 3)):
        y, launches = render(2 * 256 + 70, hidden, cell, insz, {"AIDAX_TC_GEN": "3"}, 3, 0, [128, 100, 28, 128])
        print("san H=%d cell=%d: %d launches, finite %s, peak %.4f" % (hidden, cell, launches, bool(np.isfinite(y).all()), float(np.abs(y).max())), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "san":
        sanitizer_run()
        sys.exit(0)
    if sys.argv[1] == "check":
        if len(sys.argv) > 3:
            sys.exit(check((int(sys.argv[2]), int(sys.argv[3]))))
        bad = 0
        for hidden in HS:
            for cell in (0, 1):
                try:
                    r = subprocess.run([sys.executable, "tools/tc3_check.py", "check", str(hidden), str(cell)], capture_output=True, text=True, timeout=120)
                    lines = [l for l in r.stdout.splitlines() if l.startswith("H=")]
                    print("\n".join(lines) if lines else "H=%2d cell=%d: no result, exit %d: %s" % (hidden, cell, r.returncode, (r.stderr or "").strip()[-300:]), flush=True)
                    bad += 0 if r.returncode == 0 and lines else 1
                except subprocess.TimeoutExpired:
                    print("H=%2d cell=%d: TIMEOUT (hang)" % (hidden, cell), flush=True)
                    bad += 1
        print("TC3 CHECK", "FAILED" if bad else "OK", flush=True)
        sys.exit(bad)
    timing([int(v) for v in sys.argv[2:]] or [16])
