#!/bin/bash
# Round 2, GPU call J: 8 filter warps per CTA, coefficient prefetch, sub-block length.
TAG=${1:-r02j}
OUT=$PWD/gpurun_out
mkdir -p $OUT
one() { # lib-dir-or-"" gen extra-env...
  local lib=$1 gen=$2; shift 2
  env AIDAX_TC_GEN=$gen ${lib:+AIDAX_LIB_DIR=$PWD/tools/bin/$lib} "$@" timeout 200 python -c "
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'.')
from quick_bench import run
run(378880,${H:-16},tile_mode=3,iters=10)" 2>&1 | tail -1
}
echo "== check (default build: 8 filter warps)"
TC3_HS="8 12 16 20 24 32 40" timeout 900 python tools/tc3_check.py check > $OUT/${TAG}_tc3_check.txt 2>&1; echo "exit $?"; tail -15 $OUT/${TAG}_tc3_check.txt
{
echo "gen1:            $(one '' 1)"
echo "gen3 f8:         $(one '' 3)"
echo "gen3 f8 ts16:    $(one '' 3 AIDAX_TS_MAX=16)"
echo "gen3 f8 ts24:    $(one '' 3 AIDAX_TS_MAX=24)"
echo "gen3 f8 pf:      $(one tc3_f8_pf 3)"
echo "gen3 f8 pf ts16: $(one tc3_f8_pf 3 AIDAX_TS_MAX=16)"
echo "gen3 f4 48 pf:   $(one tc3_f4_48_pf 3)"
echo "gen3 f8 nofilt:  $(one tc3_f8_w8 3)"
} | tee $OUT/${TAG}_variants.txt
echo "== all sizes"
timeout 900 python tools/tc3_check.py time 8 12 20 24 32 40 > $OUT/${TAG}_tc3_time.txt 2>&1; grep -v "^gen2" $OUT/${TAG}_tc3_time.txt
