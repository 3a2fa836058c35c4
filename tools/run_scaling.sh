#!/bin/bash
# Scaling runs of bench.py on one node, the way the driver launches them (one rank per GPU): writes gpurun_out/r2_scale_n<N>[_tag].json
# usage: tools/run_scaling.sh "<N list>" [extra bench.py arguments]      env TAG=_thp SDRM_PINNED_MODE=thp for variants
set -u
NS=${1:-"2 4 8"}
shift || true
mkdir -p gpurun_out
for N in $NS; do
  OUT=gpurun_out/r2_scale_n${N}${TAG:-}.json
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 --steps 20 --warmup 5 --full-line $OUT "$@" > /dev/null 2> gpurun_out/r2_scale_n${N}${TAG:-}.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
      bench.py --gpus $N --steps 20 --warmup 5 --full-line $OUT "$@" > /dev/null 2> gpurun_out/r2_scale_n${N}${TAG:-}.err
  fi
  echo "N=$N rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("$OUT"))
    e = d.get("e2e") or {}
    print("N=%d value %.0f ms %.3f | e2e %.0f (%.1f GB/s of ceiling %.1f, frac %.3f) int16 %.0f (frac %.3f) numa %s | configs %s" % (
        d["n_gpus"], d["value"], d["ms_per_step"], e.get("value", 0), e.get("h2d_gbs", 0), e.get("h2d_ceiling_gbs", 0),
        e.get("frac_of_ceiling", 0), e.get("int16_value", 0), e.get("int16_frac_of_ceiling", 0), e.get("pinned_numa_node"),
        [(c.get("name", "")[:12], round(c.get("value", 0))) for c in d.get("configs", [])]))
except Exception as ex:
    print("no line:", ex)
PY
done
