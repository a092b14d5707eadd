mkdir -p gpurun_out/r2g
timeout 900 python -m pytest tests/test_decomposed_gpu.py -m gpu -x -q 2>&1 | tail -15
for dd in strips tiles; do
VELVET_DD=$dd timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/_dd_gpu_worker.py gpurun_out/r2g 2047 2 10 peer 2>&1 | grep '"rank": 0' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$dd', {k:d[k] for k in ('transport','bit_identical','launches_per_frame','single_gpu_ms','decomposed_ms','owned') if k in d})
"
done
