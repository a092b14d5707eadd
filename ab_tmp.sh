timeout 900 python -m pytest tests/test_hash_gpu.py tests/test_ref_cuda_gpu.py "tests/test_solver_gpu.py::test_headline_1m_one_frame_against_the_oracle" -m gpu -x -q 2>&1 | tail -4
python - <<'PY'
import velvet_b200 as vb
for R in (1023, 4095):
    p = vb.default_params(); p.numSubsteps, p.numIterations = 5, 10
    g = vb.build_scene(R, p); g.UpdateColliders(vb.sphere_plane_colliders())
    for _ in range(8): g.Simulate()
    acc={}
    for _ in range(3):
        for k,v in g.SimulateTimed().items(): acc[k]=acc.get(k,0)+v/3
    print(R, {k:round(v,3) for k,v in acc.items() if k in ("Solver_Total","Solver_HashCache","Solver_Iterate","Solver_CollideParticles","Solver_HashSort")})
    g.close()
PY
