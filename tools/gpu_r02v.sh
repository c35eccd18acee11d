#!/bin/bash
# Round 2, GPU call V: H 32/40 with 4 filter warps (two streams per lane) instead of 8.
TAG=${1:-r02v}
OUT=$PWD/gpurun_out
mkdir -p $OUT
one() { # lib H N extra-env
  local lib=$1 H=$2 N=$3; shift 3
  env ${lib:+AIDAX_LIB_DIR=$PWD/tools/bin/$lib} "$@" timeout 200 python -c "
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'.')
from quick_bench import run
run($N,$H,tile_mode=3,iters=10)" 2>&1 | tail -1
}
{
for H in 40 32; do
echo "H=$H default (8 F warps): $(one '' $H 262144)"
echo "H=$H 4 F warps:           $(one tc3_nfw4 $H 262144)"
echo "H=$H default ts 8:        $(one '' $H 262144 AIDAX_TS_MAX=8)"
echo "H=$H 4 F warps ts 8:      $(one tc3_nfw4 $H 262144 AIDAX_TS_MAX=8)"
done
} | tee $OUT/${TAG}_nfw.txt
AIDAX_LIB_DIR=$PWD/tools/bin/tc3_nfw4 TC3_HS="40" timeout 300 python tools/tc3_check.py check 2>&1 | tail -3
