#!/bin/bash
# Round 2, GPU call F (8 GPUs): weak scaling of the default workload and the fixed-total (strong) runs of C4 / C5 at 2, 4, 8 GPUs.
TAG=${1:-r02f}
OUT=$PWD/gpurun_out
mkdir -p $OUT
{ nvidia-smi --query-gpu=index,name,pci.bus_id,pcie.link.gen.current,pcie.link.width.current --format=csv; nvidia-smi topo -m | head -14; lscpu | grep -E "^CPU\(s\)|Model name|NUMA|Socket"; nproc; free -g | head -2;
  for d in /sys/bus/pci/devices/*; do if [ -f $d/class ] && grep -q 0x0302 $d/class; then echo "$d numa=$(cat $d/numa_node) cpus=$(cat $d/local_cpulist)"; fi; done; } > $OUT/${TAG}_box.txt 2>&1
port=29500
run() { # n workload extra...
  n=$1; w=$2; shift 2
  port=$((port + 1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --workload $w --steps 20 --warmup 5 --no-literal --no-cpu-baseline "$@" \
      > $OUT/${TAG}_bench_${w}_n${n}$SUFFIX.json 2> $OUT/${TAG}_bench_${w}_n${n}$SUFFIX.err
  echo "== $w n=$n $* exit $? :: $(python - <<PY
import json
try:
    d = json.load(open("$OUT/${TAG}_bench_${w}_n${n}$SUFFIX.json")); e = d["e2e"]
    print("value %.3e ms %.3f scaling %s streams_total %d | e2e %.3e n/gpu %d us %.0f p95 %.0f meets %s pcie %s" % (d["value"], d["ms_per_step"], d["scaling"], d["config"]["streams_total"], e["value"], e["streams_per_gpu"], e["us_per_buffer"], e["us_p95"], e["meets_deadline"], (e.get("pcie") or {}).get("duplex_gbs_per_direction")))
except Exception as ex:
    print("unreadable", ex)
PY
)"
  tail -2 $OUT/${TAG}_bench_${w}_n${n}$SUFFIX.err | cut -c1-200
}
SUFFIX=""
for n in 8 4 2; do run $n c2; done
SUFFIX="_strong"
for n in 8 4 2; do run $n c4 --scaling strong --e2e-steps 15; run $n c5 --scaling strong --e2e-steps 15; done
