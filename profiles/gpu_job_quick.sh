#!/bin/bash
# Quick GPU iteration: parity tests, one bench line, full ncu captures of the iterate and neighbour-cache kernels.
# gpurun --timeout 900 -- 'bash profiles/gpu_job_quick.sh [tag]'
TAG=${1:-q}
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_$TAG.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_$TAG.log
python bench.py --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench exit $?"; python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("ms/frame", d["ms_per_step"], "iter_ms", d["roofline"]["launch_ms"], "frac", d["roofline"]["frac"], "e2e ms", d["e2e"]["ms_per_step"])
print(d["stages_ms"]); print(d["other_math_mode"])
PY
ncu --set full --import-source on --clock-control none -k regex:iterate_tile -s 120 -c 1 -f -o gpurun_out/iter_$TAG \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_iter_$TAG.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:cache_neighbors_sorted -s 4 -c 1 -f -o gpurun_out/cache_$TAG \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_cache_$TAG.log 2>&1
