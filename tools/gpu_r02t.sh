#!/bin/bash
# Round 2, GPU call T: the tile-count-generic warp-specialised kernel (default build) and the three-tile shape for H = 24.
TAG=${1:-r02t}
OUT=$PWD/gpurun_out
mkdir -p $OUT
echo "== generation tests (default build)"
timeout 600 python -m pytest tests/test_tc_generations_gpu.py tests/test_engine_gpu.py -m gpu -x -q 2>&1 | tail -3
one() { # lib H N gen
  env ${1:+AIDAX_LIB_DIR=$PWD/tools/bin/$1} AIDAX_TC_GEN=$4 timeout 200 python -c "
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'.')
from quick_bench import run
run($3,$2,tile_mode=3,iters=10)" 2>&1 | tail -1
}
echo "== H=24 three tiles: check"
AIDAX_LIB_DIR=$PWD/tools/bin/tc3_h24_3t TC3_HS="24" timeout 300 python tools/tc3_check.py check 2>&1 | tail -3
{
echo "H=24 gen1 262144: $(one tc3_h24_3t 24 262144 1)"
echo "H=24 3-tile 262144: $(one tc3_h24_3t 24 262144 3)"
echo "H=24 gen1 227328: $(one tc3_h24_3t 24 227328 1)"
echo "H=24 3-tile 227328: $(one tc3_h24_3t 24 227328 3)"
echo "H=16 default 378880: $(one '' 16 378880 3)"
} | tee $OUT/${TAG}_h24.txt
