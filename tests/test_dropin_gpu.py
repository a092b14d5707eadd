"""The kernel-level drop-in (way A), compiled and run.

oracle/ref_cuda/build_ref_cuda.sh builds oracle/_ref/libvelvet_dropin.so from
  * the reference-side orchestration of VtClothSolverGPU.hpp / SpatialHashGPU.hpp (oracle/ref_cuda/ref_driver.cu) over the
    reference's OWN VtBuffer.hpp, Timer.hpp, VtClothSolverGPU.cuh and SpatialHashGPU.cuh, exactly as for the O3 oracle, and
  * velvet_b200/csrc/dropin/VelvetB200Shim.cpp INSTEAD of the reference's VtClothSolverGPU.cu + SpatialHashGPU.cu: the twelve
    seam functions, defined against the reference's declarations, forwarding to libvelvet_b200.so.
So a maintainer's build with the two .cu files swapped for the shim links and runs; here its results are held against the
same golden vectors (outputs of the reference's own kernels) as the oracle."""
import math
import os

import numpy as np
import pytest

from oracle import o1, refcuda

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not refcuda.available(refcuda.SO_DROPIN), reason="oracle/_ref/libvelvet_dropin.so not built")]

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TOL_1 = 1e-4 * 2.0


def _pair(R, p, position, rotation, attached):
    o = o1.O1Solver(p)
    v, idx = o1.generate_cloth_mesh(R)
    model = o1.transform_matrix(position, rotation, (1, 1, 1))
    o.cloth_object_start(R, v, idx, model, attached)
    d = refcuda.RefCudaSolver(p, so=refcuda.SO_DROPIN)
    d.register_like(o, R, model, attached)
    return o, d


def _masked(nb, n, k=64):
    tab = nb[: n * k].reshape(k, n).copy()
    tab[np.cumsum(tab == 0xFFFFFFFF, axis=0) > 0] = 0xFFFFFFFF
    return tab


def test_dropin_library_links_the_product_not_the_reference_kernels():
    import subprocess
    syms = subprocess.run(["nm", "-D", refcuda.SO_DROPIN], capture_output=True, text=True).stdout
    for name in ("velvet_SolveStretch", "velvet_HashObjects", "velvet_SetSimulationParams", "velvet_ComputeNormal"):
        assert f" U {name}" in syms, f"{name} must be an undefined symbol resolved by libvelvet_b200.so"
    assert "SolveStretch_Kernel" not in syms and "CacheNeighbors_Kernel" not in syms


@pytest.mark.parametrize("R", [31, 63])
def test_dropin_hash_matches_reference_kernel_golden(R):
    g = np.load(os.path.join(GOLDEN, f"refcuda_hash_R{R}.npz"))
    n = (R + 1) ** 2
    o, d = _pair(R, o1.default_params(), (0, 1.5, 1.0), (90, 0, 0), [])
    assert np.float32(d.params.particleDiameter) == g["particleDiameter"]
    d.buffer("initialPositions")[:] = g["initialPositions"]
    d.buffer("predicted")[:] = g["predicted"]
    d.hash_predicted()
    assert np.array_equal(d.buffer("particleHash"), g["particleHash"])
    assert np.array_equal(d.buffer("particleIndex"), g["particleIndex"])
    assert np.array_equal(d.buffer("cellStart"), g["cellStart"])
    valid = g["cellStart"] != 0xFFFFFFFF
    assert np.array_equal(d.buffer("cellEnd")[valid], g["cellEnd"][valid])
    assert np.array_equal(_masked(d.buffer("neighbors"), n), g["neighbors"])


def test_dropin_config1_within_tolerance_of_reference_kernel_golden_and_oracle():
    """BASELINE configs[0] through the reference's Simulate() order over the shim: positions after 1 / 5 / 10 / 15 frames
    against the outputs of the reference's own kernels, and after one frame against the O1 oracle."""
    g = np.load(os.path.join(GOLDEN, "refcuda_cfg1.npz"))
    p = o1.default_params()
    p.numSubsteps, p.numIterations = 5, 10
    o, d = _pair(31, p, (0, 2.5, 0), (0, 0, 0), [0, 31])
    last = o1.transform_matrix((0, 0.6, -1.0), (0, 0, 0), (0.6,) * 3)
    worst = {}
    for f in range(15):
        z = -math.cos(2 * f / 60.0)
        cur = o1.transform_matrix((0, 0.6, z), (0, 0, 0), (0.6,) * 3)
        cols = [o1.make_collider(o1.PLANE, (0, 0, 0), (1, 1, 1)), o1.make_collider(o1.SPHERE, (0, 0.6, z), (0.6,) * 3, cur, last)]
        last = cur
        d.set_colliders(cols)
        d.simulate()
        if f == 0:
            o.set_colliders(cols)
            o.simulate()
            assert np.max(np.abs(d.buffer("positions") - o.buffer("positions"))) <= TOL_1
            assert np.array_equal(d.buffer("invMasses"), g["invMasses"])
            assert np.max(np.abs(d.buffer("normals") - g["normals_1"])) <= 1e-3
        if f + 1 in (1, 5, 10, 15):
            worst[f + 1] = float(np.max(np.abs(d.buffer("positions") - g[f"positions_{f + 1}"])))
            assert worst[f + 1] <= TOL_1, worst
    assert np.isfinite(d.buffer("positions")).all()
    print("drop-in (reference orchestration over the shim) vs reference CUDA kernels, config 1:", worst)
    timers = d.timers()
    assert timers.get("Solver_Total", 0) > 0  # the reference's own GPU timers keep working around the forwarded calls
