mkdir -p gpurun_out/r2h
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
run() { name=$1; res=$2; shift; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-sub-records --resolution $res > gpurun_out/r2h/bench_$name.json 2> gpurun_out/r2h/bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2h/bench_$name.json"))
    print("$name", "ms/frame", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), "iter_us", round(d["roofline"]["launch_ms"]*1e3,2), "fast", round(d["other_math_mode"]["ms_per_step"],4))
except Exception as e:
    print("$name FAILED", e)
PY
}
run pdl_1m 1023 VELVET_PDL=1
run nopdl_1m 1023 VELVET_PDL=0
run pdl_256 255 VELVET_PDL=1
run nopdl_256 255 VELVET_PDL=0
run pdl_32 31 VELVET_PDL=1
run nopdl_32 31 VELVET_PDL=0
