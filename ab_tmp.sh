mkdir -p gpurun_out/r2f
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f/bench_n1.json 2> gpurun_out/r2f/bench_n1.err
tail -5 gpurun_out/r2f/bench_n1.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2f/bench_n1.json"))
print("ms/frame", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"])
print("ref_cuda", d["ref_cuda"])
print("sub", json.dumps(d["sub_records"], indent=1)[:3000])
print("cpu", d["cpu_baseline"])
PY
