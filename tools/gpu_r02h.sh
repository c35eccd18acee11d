#!/bin/bash
# Round 2, GPU call H: first run of the warp-specialised kernel (generation 3): check against generation 1, then timing.
TAG=${1:-r02h}
OUT=$PWD/gpurun_out
mkdir -p $OUT
echo "== tc3 check (H=16 LSTM first)"
timeout 150 python tools/tc3_check.py check 16 0 2>&1 | tail -3
echo "== tc3 check (all)"
timeout 900 python tools/tc3_check.py check > $OUT/${TAG}_tc3_check.txt 2>&1; echo "exit $?"; cat $OUT/${TAG}_tc3_check.txt | tail -12
echo "== timing"
timeout 600 python tools/tc3_check.py time 16 > $OUT/${TAG}_tc3_time.txt 2>&1; cat $OUT/${TAG}_tc3_time.txt
for v in tc3_96_48; do
  if [ -d tools/bin/$v ]; then
    echo "== variant $v"
    AIDAX_LIB_DIR=$PWD/tools/bin/$v AIDAX_TC_GEN=3 timeout 300 python -c "
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'.')
from quick_bench import run
run(378880,16,tile_mode=3,iters=10); run(454656,16,tile_mode=3,iters=10)" 2>&1 | tee $OUT/${TAG}_$v.txt
  fi
done
