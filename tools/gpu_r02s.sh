#!/bin/bash
# Round 2, GPU call S (8 GPUs): weak scaling of the default workload with the warp-specialised kernel and the copy-rate-sized
# hand-over chunk; fixed-total (strong) C4 / C5 at 8 GPUs.
TAG=${1:-r02s}
OUT=$PWD/gpurun_out
mkdir -p $OUT
port=29600
run() { # n workload extra...
  n=$1; w=$2; shift 2
  port=$((port + 1))
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --workload $w --steps 20 --warmup 5 --no-literal --no-cpu-baseline "$@" \
      > $OUT/${TAG}_bench_${w}_n${n}$SUFFIX.json 2> $OUT/${TAG}_bench_${w}_n${n}$SUFFIX.err
  echo "== $w n=$n $* exit $? :: $(python - <<PY
import json
try:
    d = json.load(open("$OUT/${TAG}_bench_${w}_n${n}$SUFFIX.json")); e = d["e2e"]
    print("value %.3e ms %.3f scaling %s streams_total %d | e2e %.3e n/gpu %d us %.0f p95 %.0f meets %s pcie %s chunk %s frac_floor %s" % (d["value"], d["ms_per_step"], d["scaling"], d["config"]["streams_total"], e["value"], e["streams_per_gpu"], e["us_per_buffer"], e["us_p95"], e["meets_deadline"], (e.get("pcie") or {}).get("duplex_gbs_per_direction"), e.get("chunk_instances"), (e.get("pcie") or {}).get("frac_of_copy_floor")))
except Exception as ex:
    print("unreadable", ex)
PY
)"
  tail -2 $OUT/${TAG}_bench_${w}_n${n}$SUFFIX.err | cut -c1-200
}
SUFFIX=""
for n in 8 4 2; do run $n c2; done
SUFFIX="_strong"
run 8 c5 --scaling strong --e2e-steps 15
run 8 c4 --scaling strong --e2e-steps 15
