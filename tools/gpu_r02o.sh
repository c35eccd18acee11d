#!/bin/bash
# Round 2, GPU call O: short first sub-block (fill of the filter pipeline).
TAG=${1:-r02o}
OUT=$PWD/gpurun_out
mkdir -p $OUT
one() { # lib-dir-or-"" H N
  local lib=$1 H=$2 N=$3
  env ${lib:+AIDAX_LIB_DIR=$PWD/tools/bin/$lib} timeout 200 python -c "
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'.')
from quick_bench import run
run($N,$H,tile_mode=3,iters=10)" 2>&1 | tail -1
}
{
echo "default:   $(one '' 16 378880)"
for v in tc3_ts0_2 tc3_ts0_4 tc3_ts0_8 tc3_ts0_4_pf; do echo "$v: $(one $v 16 378880)"; done
} | tee $OUT/${TAG}_ts0.txt
for v in tc3_ts0_4; do echo "check $v: $(AIDAX_LIB_DIR=$PWD/tools/bin/$v timeout 300 python tools/tc3_check.py check 16 0 2>&1 | tail -1)"; done | tee -a $OUT/${TAG}_ts0.txt
