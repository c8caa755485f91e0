#!/bin/bash
# Multi-GPU pass (gpurun --gpus N): group API + IPC gather tests, host-copy ceiling, C5 as written (strong + weak), default bench at N.
# Usage: bash scripts/gpu_r2_multi.sh N [full]
set -u
N=${1:-2}
FULL=${2:-}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" > gpurun_out/lscpu_$N.txt; free -g >> gpurun_out/lscpu_$N.txt
echo "== multi-GPU tests"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 300 > gpurun_out/pytest_multi_$N.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_multi_$N.log
echo "== host copy ceiling"; timeout 200 python -u scripts/e2e_ceiling.py > gpurun_out/e2e_ceiling_$N.jsonl 2> gpurun_out/e2e_ceiling_$N.err; cat gpurun_out/e2e_ceiling_$N.jsonl
timeout 200 python -u scripts/e2e_ceiling.py --numa >> gpurun_out/e2e_ceiling_$N.jsonl 2>> gpurun_out/e2e_ceiling_$N.err; tail -4 gpurun_out/e2e_ceiling_$N.jsonl; tail -2 gpurun_out/e2e_ceiling_$N.err
run() {  # name, gpus, args...
  name=$1; g=$2; shift 2
  if [ $g -eq 1 ]; then
    timeout 500 python -u bench.py --gpus 1 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  else
    timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) \
      bench.py --gpus $g "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  fi
  echo "$name rc=$?"; tail -1 gpurun_out/$name.json | cut -c1-250; python - "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/{sys.argv[1]}.json").read().strip().splitlines()[-1])
    e = d["e2e"]
    print("   value %.3e  ms/step %.3f  frac %.3f  e2e %.3e (%s)  per-rank-e2e %s  gather %s" % (
        d["value"], d["ms_per_step"], d["roofline"]["frac"], e["value"], e["api"][:30],
        e.get("per_rank_processes") and "%.3e" % e["per_rank_processes"]["value"], d.get("gather")))
except Exception as ex:
    print("   (no line)", ex)
PY
}
for g in 1 2 4 8; do
  [ $g -le $N ] || continue
  if [ -n "$FULL" ] || [ $g -eq $N ]; then
    run bench_c5_strong_${g}gpu $g --workload c5 --scaling strong --steps 20 --warmup 5
  fi
done
run bench_c5_weak_${N}gpu $N --workload c5 --steps 20 --warmup 5
run bench_north_star_${N}gpu $N --steps 20 --warmup 5
