#!/bin/bash
# Round 2, GPU call M: wait loops of the warp-specialised kernel (filter warps poll + sleep; MMA wait variants).
TAG=${1:-r02m}
OUT=$PWD/gpurun_out
mkdir -p $OUT
one() { # lib-dir-or-"" extra-env...
  local lib=$1; shift 1
  env ${lib:+AIDAX_LIB_DIR=$PWD/tools/bin/$lib} "$@" timeout 200 python -c "
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'.')
from quick_bench import run
run(378880,${H:-16},tile_mode=3,iters=10)" 2>&1 | tail -1
}
echo "== check"
timeout 300 python tools/tc3_check.py check 16 0 2>&1 | tail -2
timeout 300 python tools/tc3_check.py check 40 1 2>&1 | tail -2
{
echo "default (F poll 1000 ns, MMA try_wait): $(one '')"
echo "F poll 250:    $(one tc3_f250)"
echo "F poll 4000:   $(one tc3_f4000)"
echo "MMA poll 32:   $(one tc3_r32)"
echo "MMA poll 100:  $(one tc3_r100)"
echo "MMA poll 250:  $(one tc3_r250)"
} | tee $OUT/${TAG}_waits.txt
for v in tc3_r32 tc3_r100; do AIDAX_LIB_DIR=$PWD/tools/bin/$v timeout 300 python tools/tc3_check.py check 16 0 2>&1 | tail -1; done
