timeout 900 python -m pytest tests/test_dropin_gpu.py tests/test_ref_cuda_gpu.py -m gpu -x -q -s 2>&1 | tail -8
