#!/bin/bash
# Round 2, GPU call E: fault bisect (legacy library rebuilt against the current ABI), C4 with more hardware queues, default bench, new tests.
TAG=${1:-r02e}
OUT=$PWD/gpurun_out
mkdir -p $OUT
echo "== fault bisect"
timeout 1500 bash tools/race_probe.sh run > $OUT/${TAG}_race.txt 2>&1; cat $OUT/${TAG}_race.txt
echo "== new tests"
timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -x -q -k "generator or golden" 2>&1 | tail -3
for conn in 8 32; do
  echo "== bench c4, CUDA_DEVICE_MAX_CONNECTIONS=$conn"
  CUDA_DEVICE_MAX_CONNECTIONS=$conn timeout 600 python bench.py --workload c4 --steps 20 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_c4_conn$conn.json 2> $OUT/${TAG}_bench_c4_conn$conn.err; echo "exit $?"; tail -2 $OUT/${TAG}_bench_c4_conn$conn.err
done
echo "== bench (ours)"
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"; tail -3 $OUT/${TAG}_bench.err
echo "== bench c5"
timeout 600 python bench.py --workload c5 --steps 20 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_c5.json 2> $OUT/${TAG}_bench_c5.err; echo "exit $?"
python - <<'PY'
import json, glob
for name in sorted(glob.glob("gpurun_out/r02e_bench*.json")):
    try:
        d = json.load(open(name)); e = d["e2e"]
        print(name, "dev ms", d["ms_per_step"], "| e2e n", e["streams_per_gpu"], "us", e["us_per_buffer"], "p95", e["us_p95"], e["phase_us"], "meets", e["meets_deadline"], "| roofline", d["roofline"]["bound"], d["roofline"]["frac"])
    except Exception as ex:
        print(name, "unreadable", ex)
PY
