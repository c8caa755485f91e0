#!/bin/bash
# Closing pass on N GPUs, launched exactly as the driver does it: both arms of the default bench under torchrun, the multi-GPU tests.
set -u
N=${1:-2}
mkdir -p gpurun_out
echo "== multi-GPU tests"; timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_round2.py -m gpu -q --timeout 300 > gpurun_out/pytest_multi_$N.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_multi_$N.log
P=$((29500 + RANDOM % 500))
echo "== reference arm under torchrun"; timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --impl reference --gpus $N --steps 5 --warmup 1 > gpurun_out/bench_reference_arm_${N}gpu.json 2> gpurun_out/bench_reference_arm_${N}gpu.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_reference_arm_${N}gpu.json
echo "== b200 arm under torchrun"; timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_north_star_${N}gpu.json 2> gpurun_out/bench_north_star_${N}gpu.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_north_star_${N}gpu.json; tail -3 gpurun_out/bench_north_star_${N}gpu.err
python - "$N" <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/bench_north_star_{sys.argv[1]}gpu.json").read().strip().splitlines()[-1])
print("value %.3e frac %.3f e2e %.3e (%s) launches %s clocks %s" % (d["value"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["api"][:40], d["gpu_launches"], d["clocks"]))
PY
