"""Run under compute-sanitizer on the GPU box (not collected by pytest):
    compute-sanitizer --tool memcheck  python tests/sanitize_gpu.py
    compute-sanitizer --tool racecheck python tests/sanitize_gpu.py
Exercises every kernel of both pipelines and both math modes on a small scene."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import velvet_b200 as vb  # noqa: E402
from util import gpu_params  # noqa: E402

for pipeline, mode in ((vb.PIPELINE_FUSED, vb.MATH_EXACT), (vb.PIPELINE_FUSED, vb.MATH_FAST), (vb.PIPELINE_SEAM, vb.MATH_EXACT)):
    p = gpu_params(numSubsteps=3, numIterations=3)
    g = vb.build_scene(20, p, attached=[0, 20], pipeline=pipeline, math_mode=mode, tile_size=128)
    cube = vb.MakeCollider(vb.COLLIDER_CUBE, (0.3, 0.5, 0.2), (1, 1, 1), vb.TransformMatrix((0.3, 0.5, 0.2), (0, 20, 0), (1, 1, 1)))
    g.UpdateColliders(vb.sphere_plane_colliders() + [cube])
    for _ in range(4):
        g.Simulate()
    print("ok", pipeline, mode, float(g.download("positions")[:, 1].min()))
    g.close()
