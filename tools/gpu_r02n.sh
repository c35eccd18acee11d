#!/bin/bash
# Round 2, GPU call N: XU relief in the warp-specialised kernel (integer float->double widening in the filter warps; part of the
# tanh(c') exponentials on the FMA pipe).
TAG=${1:-r02n}
OUT=$PWD/gpurun_out
mkdir -p $OUT
one() { # lib-dir-or-"" H N
  local lib=$1 H=$2 N=$3
  env ${lib:+AIDAX_LIB_DIR=$PWD/tools/bin/$lib} timeout 200 python -c "
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'.')
from quick_bench import run
run($N,$H,tile_mode=3,iters=10)" 2>&1 | tail -1
}
echo "== check H=40 GRU (default build)"
timeout 300 python tools/tc3_check.py check 40 1 2>&1 | tail -3
{
echo "default:   $(one '' 16 378880)"
for v in tc3_wi tc3_pp1 tc3_pp2 tc3_pp4 tc3_wi_pp2; do echo "$v: $(one $v 16 378880)"; done
} | tee $OUT/${TAG}_xu.txt
for v in tc3_wi tc3_pp2; do echo "check $v: $(AIDAX_LIB_DIR=$PWD/tools/bin/$v timeout 300 python tools/tc3_check.py check 16 0 2>&1 | tail -1)"; done | tee -a $OUT/${TAG}_xu.txt
