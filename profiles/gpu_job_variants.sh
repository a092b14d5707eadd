#!/bin/bash
# Experiment variants (libvelvet_b200_<variant>.so built by VELVET_VARIANT=... python -m velvet_b200.build): bench only.
# gpurun --timeout 900 -- 'bash profiles/gpu_job_variants.sh v1 v2 ...'
mkdir -p gpurun_out
for V in "$@"; do
  VELVET_VARIANT=$V python bench.py --no-cpu-baseline --steps 20 > gpurun_out/bench_var_$V.json 2> gpurun_out/bench_var_$V.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_var_$V.json"))
    print("$V", "ms/frame %.3f" % d["ms_per_step"], "iter_us %.2f" % (1e3*d["roofline"]["launch_ms"]), "fast iter_us %.2f" % (1e3*d["other_math_mode"]["iterate_launch_ms"]))
except Exception as e:
    print("$V failed", e)
PY
done
