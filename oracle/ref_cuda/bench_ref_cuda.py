#!/usr/bin/env python
"""Times the REFERENCE's own CUDA solver (oracle/_ref, built by build_ref_cuda.sh) on the GPU box: the denominator of
north_star's ">= 10x the reference CUDA solver" target.  TEST/BASELINE INFRASTRUCTURE, not part of bench.py.

    python oracle/ref_cuda/bench_ref_cuda.py [--resolution 1023] [--frames 10] [--warmup 3]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import numpy as np  # noqa: E402

from oracle import o1, refcuda  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--resolution", type=int, default=1023)
    ap.add_argument("--frames", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    R = a.resolution
    p = o1.default_params()
    p.numSubsteps, p.numIterations = 5, 10
    M = o1.transform_matrix((0, 1.5, 1.0), (90, 0, 0), (1, 1, 1))
    o = o1.O1Solver(p)
    v, idx = o1.generate_cloth_mesh(R)
    o.cloth_object_start(R, v, idx, M, [])
    r = refcuda.RefCudaSolver(p)
    r.register_like(o, R, M, [])
    cols = [o1.make_collider(o1.PLANE, (0, 0, 0), (1, 1, 1)), o1.make_collider(o1.SPHERE, (0, 0.6, 0), (0.6, 0.6, 0.6))]
    r.set_colliders(cols)
    for _ in range(a.warmup):
        r.simulate()
    stages = {}
    t0 = time.perf_counter()
    for _ in range(a.frames):
        r.simulate()
        for k, ms in r.timers().items():
            stages[k] = stages.get(k, 0.0) + ms / a.frames
    wall = (time.perf_counter() - t0) / a.frames
    n = (R + 1) ** 2
    pos = r.buffer("positions")
    print(json.dumps({"impl": "reference-cuda (oracle/_ref, unmodified kernels, sm_100a)", "particles": n,
                      "ms_per_frame_wall": wall * 1e3, "ms_per_frame_solver_total_event": stages.get("Solver_Total"),
                      "particle_substeps_per_s": n * 5 / wall, "stages_ms": {k: round(v, 4) for k, v in stages.items()},
                      "finite": bool(np.isfinite(pos).all())}))


if __name__ == "__main__":
    main()
