#!/bin/bash
# Round 2, GPU call D: fault bisect, tc2 re-check (divergent tcgen05.ld fixed), kernel variants, e2e trace + graded hand-over, C4.
TAG=${1:-r02d}
OUT=$PWD/gpurun_out
mkdir -p $OUT
echo "== fault bisect"
timeout 1500 bash tools/race_probe.sh run > $OUT/${TAG}_race.txt 2>&1; cat $OUT/${TAG}_race.txt
echo "== tc2 check (default build: H=40 shape)"
timeout 600 python tools/tc2_check.py check 40 0 2>&1 | tail -6
timeout 600 python tools/tc2_check.py check 40 1 2>&1 | tail -6
echo "== variant timing"
timeout 900 python tools/variant_time.py fused poly4 fusedpoly4 2>&1 | tee $OUT/${TAG}_variant_time.txt | grep -v "^$" | tail -30
echo "== fused passes: bit-exactness and parity on the variant library"
AIDAX_LIB_DIR=$PWD/tools/bin/fused timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_workloads_gpu.py -m gpu -x -q -k "not full_size" 2>&1 | tail -4
echo "== e2e trace at 196608 instances, graded hand-over"
AIDAX_TRACE=1 AIDAX_TRACE_SKIP=12 timeout 600 python bench.py --steps 5 --warmup 3 --no-literal --no-cpu-baseline --e2e-streams 196608 > $OUT/${TAG}_bench_trace_graded.json 2> $OUT/${TAG}_trace_graded.err; echo "exit $?"; grep -A 24 "aidax trace" $OUT/${TAG}_trace_graded.err | head -60
echo "== e2e trace at 196608 instances, chunk-by-chunk hand-over"
AIDAX_BATCH_GRADED=0 AIDAX_TRACE=1 AIDAX_TRACE_SKIP=12 timeout 600 python bench.py --steps 5 --warmup 3 --no-literal --no-cpu-baseline --e2e-streams 196608 > $OUT/${TAG}_bench_trace_chunks.json 2> $OUT/${TAG}_trace_chunks.err; echo "exit $?"; grep -A 14 "aidax trace" $OUT/${TAG}_trace_chunks.err | head -32
echo "== bench (ours)"
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"; tail -3 $OUT/${TAG}_bench.err
for w in c4 c5; do
  echo "== bench $w"
  timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_$w.json 2> $OUT/${TAG}_bench_$w.err; echo "exit $?"; tail -2 $OUT/${TAG}_bench_$w.err
done
python - <<'PY'
import json
for f in ("trace_graded", "trace_chunks", "", "c4", "c5"):
    name = "gpurun_out/r02d_bench%s.json" % ("_" + f if f else "")
    try:
        d = json.load(open(name)); e = d["e2e"]
        print(name, "dev ms", d["ms_per_step"], "| e2e n", e["streams_per_gpu"], "us", e["us_per_buffer"], "p95", e["us_p95"], e["phase_us"], "meets", e["meets_deadline"])
    except Exception as ex:
        print(name, "unreadable", ex)
PY
