#!/bin/bash
# Round 2, GPU call AI: x(t+1) written into the A row before the epilogue instead of after it.
OUT=$PWD/gpurun_out
one() { env ${1:+AIDAX_LIB_DIR=$PWD/tools/bin/$1} timeout 200 python -c "
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'.')
from quick_bench import run
run($3,$2,tile_mode=3,iters=20)" 2>&1 | tail -1 | sed 's/.*N=/N=/'; }
AIDAX_LIB_DIR=$PWD/tools/bin/tc3_earlyx TC3_HS="16 40" timeout 300 python tools/tc3_check.py check 2>&1 | tail -5
{
echo "default 378880:  $(one '' 16 378880)"
echo "early-x 378880:  $(one tc3_earlyx 16 378880)"
echo "default 16384:   $(one '' 16 16384)"
echo "early-x 16384:   $(one tc3_earlyx 16 16384)"
echo "default H40:     $(one '' 40 262144)"
echo "early-x H40:     $(one tc3_earlyx 40 262144)"
} | tee $OUT/r02ai_early_x.txt
