#!/bin/bash
# Round 2, GPU call AA: small families -- FP32 shapes (auto) against the one-tile warp-specialised kernel, per hidden size.
TAG=${1:-r02aa}
OUT=$PWD/gpurun_out
mkdir -p $OUT
timeout 900 python - <<'PY' 2>&1 | grep -v "^$" | tee $OUT/r02aa_small_families.txt
import sys, os
sys.path.insert(0, 'tools'); sys.path.insert(0, '.')
from quick_bench import run
for H in (12, 16, 20, 24, 32, 40):
    for N in (512, 1024, 2048, 4096):
        if H <= 16 and N < 2048:
            continue
        os.environ["AIDAX_TC_MIN_STREAMS"] = "1000000"
        print("H=%d N=%d auto-fp32:" % (H, N), end=""); run(N, H, tile_mode=0, iters=20)
        os.environ["AIDAX_TC_MIN_STREAMS"] = "4096"
        print("H=%d N=%d tc3 1-tile:" % (H, N), end=""); run(N, H, tile_mode=3, iters=20)
PY
