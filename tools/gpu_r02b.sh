#!/bin/bash
# Round 2, GPU call B: round-1 fault with a GPU core dump, legacy arms, backtrace of the bench crash, TS-form probe, tc2 check + timing.
TAG=${1:-r02b}
OUT=$PWD/gpurun_out
mkdir -p $OUT
echo "== round-1 build, workaround off, GPU core dump"
( cd tools/bin/r01tree
  for mode in "single20 65536" "single20 65536" "all42 65536" "mid2 65536"; do
    for defer in 0 1; do
      rm -f /tmp/aidax_core*
      AIDAX_DEFER_HETERO=$defer AIDAX_UPLOAD_ORDER=0 AIDAX_TILE_MODE=1 CUDA_ENABLE_COREDUMP_ON_EXCEPTION=1 CUDA_COREDUMP_FILE=/tmp/aidax_core_%p \
        timeout 180 python tools/c4_probe.py $mode > /tmp/probe.out 2>&1
      echo "r01 $mode defer=$defer rc=$? :: $(tail -1 /tmp/probe.out | cut -c1-160)"
      core=$(ls /tmp/aidax_core* 2>/dev/null | head -1)
      if [ -n "$core" ]; then
        echo "core: $core"; ls -la $core
        timeout 200 cuda-gdb -batch -ex "target cudacore $core" -ex "info cuda kernels" -ex "info cuda exception" -ex "bt" -ex "info cuda lanes" -ex "x/12i \$pc-64" -ex "info registers" 2>&1 | grep -v "^\[New\|^warning" | head -150 | cut -c1-300 > $OUT/${TAG}_coredump.txt
        head -60 $OUT/${TAG}_coredump.txt
        break 2
      fi
    done
  done ) 2>&1 | tee $OUT/${TAG}_r01_fault.txt
echo "== legacy arms (table fill restored)"
timeout 900 bash tools/race_probe.sh run 65536 30 > $OUT/${TAG}_race.txt 2>&1; cat $OUT/${TAG}_race.txt
echo "== bench default under the backtrace preload"
PYTHONFAULTHANDLER=1 LD_PRELOAD=$PWD/tools/bin/bt_preload.so timeout 600 python -X faulthandler bench.py --steps 5 --warmup 3 --no-literal --no-cpu-baseline > $OUT/${TAG}_bench_bt.json 2> $OUT/${TAG}_bench_bt.err; echo "exit $?"; tail -60 $OUT/${TAG}_bench_bt.err | cut -c1-250
echo "== TS-form probe"
timeout 120 tools/bin/tc_ts_probe 2>&1 | tee $OUT/${TAG}_ts_probe.txt
echo "== tc2 check"
timeout 900 python tools/tc2_check.py check 2>&1 | tee $OUT/${TAG}_tc2_check.txt | tail -50
echo "== tc2 timing"
timeout 900 python tools/tc2_check.py time 16 40 32 24 20 12 8 2>&1 | tee $OUT/${TAG}_tc2_time.txt | tail -80
