"""Registration of a grid cloth on the device (velvet_b200/csrc/setup_kernels.cuh) against the host
generator it replaces (VELVET_HOST_GENERATE=1: GenerateGridConstraints, the restatement of VtClothObjectGPU.hpp L43-148 that
the oracle-parity tests pinned in round 1): every public list bit for bit, the same kernel selection, the same frames."""
import os

import numpy as np
import pytest

import velvet_b200 as vb
from util import gpu_params

pytestmark = pytest.mark.gpu

LISTS = ("positions", "indices", "invMasses", "velocities", "predicted", "stretchIndices", "stretchLengths", "bendIndices",
         "bendAngles", "attachParticleIDs", "attachSlotIDs", "attachDistances", "attachSlotPositions", "initialPositions")


def _build(cloths, host, mutate_indices=None):
    old = os.environ.pop("VELVET_HOST_GENERATE", None)
    if host:
        os.environ["VELVET_HOST_GENERATE"] = "1"
    try:
        g = vb.VtClothSolverGPU(gpu_params(numSubsteps=3, numIterations=5))
        for R, pos, rot, attached in cloths:
            v, idx = vb.GenerateClothMesh(R)
            if mutate_indices is not None:
                idx = mutate_indices(idx)
            o = vb.VtClothObjectGPU(R, g)
            o.SetAttachedIndices(attached)
            o.Start(v, idx, vb.TransformMatrix(pos, rot, (1, 1, 1)))
        g.UpdateColliders(vb.sphere_plane_colliders())
        return g
    finally:
        os.environ.pop("VELVET_HOST_GENERATE", None)
        if old is not None:
            os.environ["VELVET_HOST_GENERATE"] = old


def _same_lists(a, b):
    for name in LISTS:
        x, y = a.download(name), b.download(name)
        assert x.shape == y.shape, name
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32)), f"{name} differs between device and host registration"


CASES = {
    "one cloth": [(33, (0, 1.5, 1.0), (90, 0, 0), ())],
    "attached corners, tilted": [(40, (0.1, 1.3, 0.7), (70, 20, 5), (0, 40))],
    "odd size, three slots": [(14, (0, 1.2, 0.2), (35, 0, 10), (0, 7, 224))],
    "two cloths, second attached": [(30, (0, 1.5, 1.0), (90, 0, 0), ()), (19, (0.1, 1.62, 0.9), (90, 0, 0), (0, 19))],
    "R = 1": [(1, (0, 1.0, 0), (90, 0, 0), ())],
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_device_registration_writes_the_host_generators_lists(case):
    dev, host = _build(CASES[case], host=False), _build(CASES[case], host=True)
    _same_lists(dev, host)
    for _ in range(4):
        dev.Simulate()
        host.Simulate()
    assert dev.iterateKernel == host.iterateKernel == vb.ITERATE_GRID
    for name in ("positions", "velocities", "normals", "predicted", "neighbors"):
        assert np.array_equal(dev.download(name).view(np.uint32), host.download(name).view(np.uint32)), name


def test_another_triangulation_is_noticed_on_the_device():
    """Bending quads come from the mesh's own triangles: with the diagonal of one quad flipped the lists are no longer the
    grid pattern, which the device-side plan builder must notice (tile plan instead), exactly like the host builder."""
    def flip(idx):
        idx = idx.copy()
        q = 6 * 17
        idx[q:q + 6] = idx[q:q + 6][[1, 2, 0, 4, 5, 3]]   # the same two triangles of quad 17, vertices rotated
        return idx
    dev, host = _build(CASES["one cloth"], host=False, mutate_indices=flip), _build(CASES["one cloth"], host=True, mutate_indices=flip)
    _same_lists(dev, host)
    for _ in range(3):
        dev.Simulate()
        host.Simulate()
    assert dev.iterateKernel == host.iterateKernel == vb.ITERATE_TILES
    for name in ("positions", "velocities", "normals"):
        assert np.array_equal(dev.download(name).view(np.uint32), host.download(name).view(np.uint32)), name


def test_constraints_added_by_hand_after_a_generated_cloth_are_honoured():
    """An extra AddStretch after Start: the lists are no longer exactly the generator's, so the plan must come from the lists."""
    out = []
    for host in (False, True):
        g = _build(CASES["one cloth"], host=host)
        g.AddStretch(0, 33 * 34 + 33, 0.5)
        for _ in range(3):
            g.Simulate()
        assert g.iterateKernel == vb.ITERATE_TILES
        out.append(g.download("positions"))
    assert np.array_equal(out[0].view(np.uint32), out[1].view(np.uint32))
