#!/bin/bash
# The driver's scaling run, reproduced: default bench at N = 1, 2, 4, 8 on one box, launched as the driver launches it.
set -u
mkdir -p gpurun_out
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then
    timeout 500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/scale_north_star_${N}gpu.json 2> gpurun_out/scale_north_star_${N}gpu.err
  else
    timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/scale_north_star_${N}gpu.json 2> gpurun_out/scale_north_star_${N}gpu.err
  fi
  echo "N=$N rc=$?"
  python - "$N" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/scale_north_star_{sys.argv[1]}gpu.json").read().strip().splitlines()[-1])
    print("  value %.4e ms/step %.3f frac %.3f e2e %.3e clocks %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["clocks"]))
except Exception as e:
    print("  no line:", e)
PY
done
