mkdir -p gpurun_out/r2m
timeout 900 python -m pytest tests/test_decomposed_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/_dd_gpu_worker.py gpurun_out/r2m 2047 2 10 peer > /dev/null 2>&1
python -c "
import json
for r in (0,1):
    d=json.load(open(f'gpurun_out/r2m/dd_gpu{r}.json')); print({k:d[k] for k in ('transport','bit_identical','launches_per_frame','single_gpu_ms','decomposed_ms','owned') if k in d})"
