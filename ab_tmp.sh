mkdir -p gpurun_out/r2i
timeout 900 python -m pytest tests/test_decomposed_gpu.py tests/test_solver_gpu.py -m gpu -x -q 2>&1 | tail -5
python - <<'PY'
import time, numpy as np
import velvet_b200 as vb
for R in (1023, 4095):
    p = vb.default_params(); p.numSubsteps, p.numIterations = 5, 10
    t0=time.perf_counter(); g = vb.build_scene(R, p); t1=time.perf_counter()
    g.UpdateColliders(vb.sphere_plane_colliders()); g.Simulate(); t2=time.perf_counter()
    print(R, "register %.3f s, first Simulate (plans + graph) %.3f s" % (t1-t0, t2-t1), "kernel", g.iterateKernel)
    g.close()
PY
