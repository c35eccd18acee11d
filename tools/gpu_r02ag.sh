#!/bin/bash
# Round 2, GPU call AG: last small knobs of the warp-specialised kernel (register split 104/24, first sub-block 6 / 12).
OUT=$PWD/gpurun_out
one() { env ${1:+AIDAX_LIB_DIR=$PWD/tools/bin/$1} timeout 200 python -c "
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'.')
from quick_bench import run
run(378880,16,tile_mode=3,iters=20)" 2>&1 | tail -1 | sed 's/.*N=/N=/'; }
{
echo "default:    $(one '')"
echo "R104/F24:   $(one tc3_r104)"
echo "TS0=12:     $(one tc3_ts0_12)"
echo "TS0=6:      $(one tc3_ts0_6)"
echo "default:    $(one '')"
} | tee $OUT/r02ag_knobs.txt
AIDAX_LIB_DIR=$PWD/tools/bin/tc3_r104 TC3_HS=16 timeout 200 python tools/tc3_check.py check 16 0 2>&1 | tail -1
