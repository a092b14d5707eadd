mkdir -p gpurun_out/golden
timeout 900 python tests/golden/make_golden.py gpurun_out/golden --only-hash-1m 2>&1 | tail -3
timeout 1500 python -m pytest tests/test_decomposed_gpu.py tests/test_instances_gpu.py "tests/test_solver_gpu.py::test_headline_1m_one_frame_against_the_oracle" "tests/test_solver_gpu.py::test_iterate_kernel_selection" "tests/test_solver_gpu.py::test_two_grid_cloths_in_one_solver_match_the_oracle" -m gpu -x -q 2>&1 | tail -12
