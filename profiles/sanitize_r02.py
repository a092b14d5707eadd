"""Workload for compute-sanitizer (round 2): the fused pipeline with BOTH Jacobi kernels on a small self-colliding drape with
attachments (grid kernel: partially filled tiles, cloth borders, attach CSR; tile kernel: forced), two cloths in one solver,
batched instances (with pinned particles), the 14 x 16 tile shape (cloth side 32), registration on the device (every scene goes
through setup_kernels.cu), the device-side grabber and the NaN guard.
    compute-sanitizer --tool memcheck  python profiles/sanitize_r02.py
    compute-sanitizer --tool racecheck python profiles/sanitize_r02.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import velvet_b200 as vb  # noqa: E402

p = vb.default_params()
p.numSubsteps, p.numIterations = 3, 4
cols = vb.sphere_plane_colliders()
for mode in (vb.ITERATE_AUTO, vb.ITERATE_TILES):
    g = vb.build_scene(37, p, attached=(0, 37))
    g.SetIterateMode(mode)
    g.UpdateColliders(cols)
    for _ in range(2):
        g.Simulate()
    print("ok", mode, g.iterateKernel, float(np.abs(g.download("positions")).sum()))
    g.close()
g = vb.VtClothSolverGPU(p)
for R, pos in ((16, (0, 1.5, 1.0)), (9, (0.1, 1.62, 0.9))):
    v, idx = vb.GenerateClothMesh(R)
    vb.VtClothObjectGPU(R, g).Start(v, idx, vb.TransformMatrix(pos, (90, 0, 0), (1, 1, 1)))
g.UpdateColliders(cols)
g.Simulate()
print("ok two cloths", g.iterateKernel, g.CheckNaN())
g.close()
g = vb.VtClothSolverGPU(p)
v, idx = vb.GenerateClothMesh(20)
g.AddClothInstances(20, v, idx, [vb.TransformMatrix((0, 1.5 + 0.01 * k, 1.0), (90, 0, 0), (1, 1, 1)) for k in range(3)], ())
g.UpdateColliders(cols)
g.Simulate()
print("ok instances", g.iterateKernel, bool(np.isfinite(g.download("positions")).all()))
g.close()
g = vb.build_scene(31, p, attached=(0, 31))  # side 32: the 14 x 16 tile shape of the grid kernel
g.UpdateColliders(cols)
g.Simulate()
idx, dist = g.Grab(np.array([0.3, 3.0, 2.5], np.float32), np.array([-0.1, -0.55, -0.6], np.float32))
g.Drag(np.array([0.3, 3.0, 2.5], np.float32), np.array([-0.1, -0.5, -0.6], np.float32))
g.Simulate()
g.Release()
g.Simulate()
print("ok side 32 + grabber", g.iterateKernel, idx, bool(np.isfinite(g.download("positions")).all()))
g.close()
g = vb.VtClothSolverGPU(p)
v, idx = vb.GenerateClothMesh(15)
g.AddClothInstances(15, v, idx, [vb.TransformMatrix((0, 1.5 + 0.01 * k, 1.0), (90, 0, 0), (1, 1, 1)) for k in range(4)], (0, 15))
g.UpdateColliders(cols)
g.Simulate()
print("ok pinned instances", g.iterateKernel, int((g.download("invMasses") == 0).sum()))
g.close()
