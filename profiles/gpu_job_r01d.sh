#!/bin/bash
# Round-1 (session d) GPU job: golden fixtures from the reference CUDA kernels, GPU parity tests, bench lines, ncu launch
# list and a full ncu capture (with source) of the dominant kernel.  Run as: gpurun --timeout 1500 -- 'bash profiles/gpu_job_r01d.sh'
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt
python tests/golden/make_golden.py gpurun_out/golden > gpurun_out/golden.log 2>&1
echo "golden exit $?"
( time timeout 700 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_exact.json 2> gpurun_out/bench_exact.err
echo "bench exit $?"; cat gpurun_out/bench_exact.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01d.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:iterate_tile -s 120 -c 1 -f -o gpurun_out/iter_exact_r01d \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_iter.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:cache_neighbors_sorted -s 4 -c 1 -f -o gpurun_out/cache_r01d \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_cache.log 2>&1
( compute-sanitizer --tool memcheck python tests/sanitize_gpu.py; compute-sanitizer --tool racecheck python tests/sanitize_gpu.py ) 2>&1 | grep -E "^ok|SUMMARY|ERROR|hazard" > gpurun_out/sanitizer_r01d.txt
python bench.py --math fast --no-cpu-baseline > gpurun_out/bench_fast.json 2> gpurun_out/bench_fast.err
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
ls -la gpurun_out
