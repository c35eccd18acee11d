#!/bin/bash
# Round 2, GPU call C: round-1 fault (no core-dump env this time), full GPU suite, tc2 shapes, bench, chunk sweep, sanitizer.
TAG=${1:-r02c}
OUT=$PWD/gpurun_out
mkdir -p $OUT
echo "== round-1 build, workaround off"
( cd tools/bin/r01tree
  fault=""
  for mode in "single20 65536" "all42 65536"; do
    for defer in 0 1; do
      for try in 1 2; do
        AIDAX_DEFER_HETERO=$defer AIDAX_UPLOAD_ORDER=0 AIDAX_TILE_MODE=1 timeout 150 python tools/c4_probe.py $mode > /tmp/probe.out 2>&1
        rc=$?
        echo "r01 $mode defer=$defer try=$try rc=$rc :: $(tail -1 /tmp/probe.out | cut -c1-200)"
        if [ $rc -ne 0 ] && [ -z "$fault" ]; then fault="$mode|$defer"; fi
      done
    done
  done
  if [ -n "$fault" ]; then
    mode=${fault%|*}; defer=${fault#*|}
    echo "-- faulting case: $mode defer=$defer; with AIDAX_DEBUG_SYNC=1 (synchronise after every launch, name the launch)"
    AIDAX_DEBUG_SYNC=1 AIDAX_DEFER_HETERO=$defer AIDAX_UPLOAD_ORDER=0 AIDAX_TILE_MODE=1 timeout 150 python tools/c4_probe.py $mode 2>&1 | tail -3 | cut -c1-400
    echo "-- with the round-1 workaround (AIDAX_UPLOAD_ORDER=1)"
    AIDAX_DEFER_HETERO=$defer AIDAX_UPLOAD_ORDER=1 AIDAX_TILE_MODE=1 timeout 150 python tools/c4_probe.py $mode 2>&1 | tail -1 | cut -c1-200
    echo "-- lightweight core dump"
    rm -f /tmp/aidax_core*
    CUDA_ENABLE_COREDUMP_ON_EXCEPTION=1 CUDA_ENABLE_LIGHTWEIGHT_COREDUMP=1 CUDA_COREDUMP_FILE=/tmp/aidax_core_%p AIDAX_DEFER_HETERO=$defer AIDAX_UPLOAD_ORDER=0 AIDAX_TILE_MODE=1 \
      timeout 200 python tools/c4_probe.py $mode 2>&1 | tail -2 | cut -c1-300
    core=$(ls /tmp/aidax_core* 2>/dev/null | head -1)
    if [ -n "$core" ]; then
      ls -la $core
      timeout 200 cuda-gdb -batch -ex "target cudacore $core" -ex "info cuda kernels" -ex "info cuda exception" -ex "bt" -ex "x/12i \$pc-64" 2>&1 | grep -v "^\[New\|^warning" | head -80 | cut -c1-300 > $OUT/${TAG}_coredump.txt
      cat $OUT/${TAG}_coredump.txt
    fi
  else
    echo "-- no fault in 8 runs of the round-1 build on this box"
  fi ) 2>&1 | tee $OUT/${TAG}_r01_fault.txt
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -15 $OUT/${TAG}_pytest_gpu.log
echo "== smoke"
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2
echo "== tc2 check"
timeout 1500 python tools/tc2_check.py check 2>&1 | tee $OUT/${TAG}_tc2_check.txt | tail -60
echo "== tc2 timing"
timeout 600 python tools/tc2_check.py time 40 32 16 2>&1 | tee $OUT/${TAG}_tc2_time.txt | tail -40
echo "== bench (ours)"
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"; head -c 1500 $OUT/${TAG}_bench.json; tail -5 $OUT/${TAG}_bench.err
echo "== bench (reference arm)"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/${TAG}_bench_reference.json 2>/dev/null; echo "ref exit $?"
for chunk in 8192 32768; do
  echo "== e2e at 196608 instances, chunk $chunk"
  AIDAX_BATCH_CHUNK=$chunk timeout 600 python bench.py --steps 5 --warmup 3 --no-literal --no-cpu-baseline --e2e-streams 196608 > $OUT/${TAG}_bench_chunk$chunk.json 2> /dev/null; echo "exit $?"
done
echo "== bench c4"
timeout 600 python bench.py --workload c4 --steps 20 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_c4.json 2> $OUT/${TAG}_bench_c4.err; echo "exit $?"
echo "== compute-sanitizer racecheck / initcheck on the plugin path (2048 instances, one re-binding round)"
timeout 400 compute-sanitizer --tool racecheck --print-limit 5 python tools/race_loop.py 2048 1 16 > $OUT/${TAG}_racecheck.txt 2>&1; echo "racecheck exit $?"; tail -4 $OUT/${TAG}_racecheck.txt
timeout 400 compute-sanitizer --tool initcheck --print-limit 5 python tools/race_loop.py 2048 1 16 > $OUT/${TAG}_initcheck.txt 2>&1; echo "initcheck exit $?"; tail -4 $OUT/${TAG}_initcheck.txt
