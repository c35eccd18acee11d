#!/bin/bash
# Round 2, GPU call S4 (4 GPUs): fixed-total (strong) C4 / C5 at 2 and 4 GPUs with the warp-specialised kernels.
TAG=${1:-r02s}
OUT=$PWD/gpurun_out
mkdir -p $OUT
port=29700
run() { # n workload extra...
  n=$1; w=$2; shift 2
  port=$((port + 1))
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --workload $w --steps 20 --warmup 5 --no-literal --no-cpu-baseline "$@" \
      > $OUT/${TAG}_bench_${w}_n${n}_strong.json 2> $OUT/${TAG}_bench_${w}_n${n}_strong.err
  echo "== $w n=$n $* exit $? :: $(python - <<PY
import json
try:
    d = json.load(open("$OUT/${TAG}_bench_${w}_n${n}_strong.json")); e = d["e2e"]
    print("value %.3e ms %.3f scaling %s streams_total %d | e2e %.3e n/gpu %d us %.0f p95 %.0f meets %s pcie %s chunk %s" % (d["value"], d["ms_per_step"], d["scaling"], d["config"]["streams_total"], e["value"], e["streams_per_gpu"], e["us_per_buffer"], e["us_p95"], e["meets_deadline"], (e.get("pcie") or {}).get("duplex_gbs_per_direction"), e.get("chunk_instances")))
except Exception as ex:
    print("unreadable", ex)
PY
)"
}
for n in 4 2; do run $n c5 --scaling strong --e2e-steps 15; run $n c4 --scaling strong --e2e-steps 15; done
