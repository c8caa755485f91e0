#!/bin/bash
# Round-2 closing pass on one GPU: smoke, the whole -m gpu suite, one bench line per workload (the default run first, exactly as
# the driver launches it), the reference arm, the north-star line with the optional cuFFT comparator.
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 240 python -u -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== default bench (driver's command)"; timeout 400 python bench.py > gpurun_out/bench_north_star.json 2> gpurun_out/bench_north_star.err; echo "rc=$?"; cut -c1-330 gpurun_out/bench_north_star.json
echo "== reference arm"; timeout 400 python bench.py --impl reference > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err; echo "rc=$?"; cut -c1-330 gpurun_out/bench_reference_arm.json
for wl in c2 c5 c3 c4_2e20 c4_2e24 c1 run_ts_2048 run_ts_4096; do
  echo "== bench $wl"; timeout 400 python -u bench.py --workload $wl > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "rc=$?"; cut -c1-260 gpurun_out/bench_$wl.json; tail -2 gpurun_out/bench_$wl.err
done
echo "== north-star with the cuFFT comparator (context)"; timeout 400 python bench.py --comparator > gpurun_out/bench_north_star_comparator.json 2> gpurun_out/bench_north_star_comparator.err; echo "rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_north_star_comparator.json").read().strip().splitlines()[-1])
print(d.get("library_comparator"))
PY
