#!/bin/bash
# Round 2, GPU call I: where the warp-specialised kernel's time goes (what-if builds, ncu source view), SPLIT-2 shapes for H 24/32/40.
TAG=${1:-r02i}
OUT=$PWD/gpurun_out
mkdir -p $OUT
for v in tc3_w8 tc3_w1 tc3_w2 tc3_w9; do
  echo "== what-if $v (gen3 then gen1)"
  for gen in 3 1; do
  AIDAX_LIB_DIR=$PWD/tools/bin/$v AIDAX_TC_GEN=$gen timeout 200 python -c "
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'.')
from quick_bench import run
run(378880,16,tile_mode=3,iters=10)" 2>&1 | tail -1 | tee -a $OUT/${TAG}_whatif.txt
  done
done
echo "== SPLIT-2 shapes: check"
TC3_HS="24 32 40" timeout 900 python tools/tc3_check.py check > $OUT/${TAG}_tc3_check_split2.txt 2>&1; echo "exit $?"; tail -8 $OUT/${TAG}_tc3_check_split2.txt
echo "== SPLIT-2 shapes: timing"
timeout 900 python tools/tc3_check.py time 40 32 24 > $OUT/${TAG}_tc3_time_split2.txt 2>&1; cat $OUT/${TAG}_tc3_time_split2.txt
echo "== other small sizes: timing"
timeout 600 python tools/tc3_check.py time 8 12 20 > $OUT/${TAG}_tc3_time_small.txt 2>&1; cat $OUT/${TAG}_tc3_time_small.txt
echo "== ncu source view gen3 H=16"
AIDAX_TC_GEN=3 timeout 400 ncu --set full --clock-control none --import-source on -k regex:chain_kernel_tc3 -s 3 -c 1 -o $OUT/${TAG}_tc3_16 -f \
    python tools/prof_one.py 378880 16 3 2 > $OUT/${TAG}_ncu_tc3_16.log 2>&1; echo "ncu exit $?"
ls -la $OUT | grep $TAG
