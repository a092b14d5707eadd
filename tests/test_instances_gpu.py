"""Batched independent cloths (BASELINE config 4 / north_star "batched instances shard with no communication"):
K instances of one grid cloth in ONE solver, each with its own hash-table rows and neighbour lists, sharing one
constraint set.  Every instance must evolve exactly like a stand-alone solver of that cloth -- bit-identical to the
oracle when the oracle is given the shared (instance 0) rest lengths, and within north_star's tolerance of the oracle
with its own rest lengths (which differ from instance 0's in the last bit only)."""
import numpy as np
import pytest

import velvet_b200 as vb
from oracle import o1

from util import EXTENT, gpu_params, max_abs_diff, to_o1_collider, to_o1_params, valid_prefix_table

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["grid", "tiles"])
def iterate_kernel_param(request, monkeypatch):
    """Every test runs twice: with the implicit-grid Jacobi kernel (the default for grid cloths) and with the record-driven
    tile kernel forced (VELVET_ITERATE=tiles is read when a solver builds its plans)."""
    if request.param == "tiles":
        monkeypatch.setenv("VELVET_ITERATE", "tiles")
    else:
        monkeypatch.delenv("VELVET_ITERATE", raising=False)
    return request.param
TOL_1 = 1e-4 * EXTENT


def _models(k):
    # config 4 placement: T(0, 1.5 + 0.01 * (i mod 32), 1) * Rx(90); all instances overlap in space on purpose
    return [vb.TransformMatrix((0.003 * i, 1.5 + 0.01 * (i % 32), 1.0), (90, 0, 0), (1, 1, 1)) for i in range(k)]


def _oracles(R, p, models, attached, shared):
    v, idx = o1.generate_cloth_mesh(R)
    outs = []
    for i, M in enumerate(models):
        o = o1.O1Solver(to_o1_params(p))
        o.cloth_object_start(R, v, idx, M, attached)
        if shared and i > 0:
            o.buffer("stretchLengths")[:] = outs[0].buffer("stretchLengths")
            if len(attached):
                o.buffer("attachDistances")[:] = outs[0].buffer("attachDistances")
        outs.append(o)
    return outs


@pytest.mark.parametrize("attached", [(), (0, 24)])
def test_instances_are_bit_identical_to_standalone_oracles(attached):
    R, K = 24, 5
    n = (R + 1) ** 2
    p = gpu_params(numSubsteps=5, numIterations=10)
    models = _models(K)
    v, idx = vb.GenerateClothMesh(R)
    g = vb.VtClothSolverGPU(p)
    g.AddClothInstances(R, v, idx, models, attached)
    assert g.simParams.numParticles == K * n
    oracles = _oracles(R, p, models, list(attached), shared=True)
    loose = _oracles(R, p, models, list(attached), shared=False)
    cols = vb.sphere_plane_colliders()
    g.UpdateColliders(cols)
    for o in oracles + loose:
        o.set_colliders([to_o1_collider(c) for c in cols])
    pos0 = g.download("positions").reshape(K, n, 3)
    for i in range(K):
        assert np.array_equal(pos0[i].reshape(-1), oracles[i].buffer("positions")), "registration"
    frames = 12
    for f in range(frames):
        g.Simulate()
        for o in oracles:
            o.simulate()
        if f == 0:
            for o in loose:
                o.simulate()
            first = g.download("positions").reshape(K, n, 3)
            for i in range(K):
                assert max_abs_diff(first[i], loose[i].buffer("positions")) <= TOL_1, "vs the oracle with its own rest lengths"
    pos = g.download("positions").reshape(K, n, 3)
    vel = g.download("velocities").reshape(K, n, 3)
    nrm = g.download("normals").reshape(K, n, 3)
    for i in range(K):
        assert np.array_equal(pos[i].reshape(-1), oracles[i].buffer("positions")), f"instance {i}"
        assert np.array_equal(vel[i].reshape(-1), oracles[i].buffer("velocities")), f"instance {i}"
        assert np.array_equal(nrm[i].reshape(-1), oracles[i].buffer("normals")), f"instance {i}"
    # integer work: each instance's slice of the global sort is that instance's own stable sort, offset by i*n / i*2n
    ph, pi = g.download("particleHash"), g.download("particleIndex")
    nb = valid_prefix_table(g.download("neighbors"), K * n, 64)
    for i in range(K):
        o = oracles[i]
        assert np.array_equal(ph[i * n:(i + 1) * n], o.buffer("particleHash") + i * 2 * n)
        assert np.array_equal(pi[i * n:(i + 1) * n], o.buffer("particleIndex") + i * n)
        ref = valid_prefix_table(o.buffer("neighbors"), n, 64).astype(np.int64)
        ref = np.where(ref == 0xFFFFFFFF, 0xFFFFFFFF, ref + i * n)
        assert np.array_equal(nb[:, i * n:(i + 1) * n].astype(np.int64), ref), f"neighbour lists of instance {i}"


def test_config4_shape_64x64_cloths_match_their_oracles(iterate_kernel_param):
    """BASELINE configs[3] at its real cloth size: 64 instances of the 64x64 cloth in one solver (the bench runs 4,096 of
    them); a sample of the instances is checked bit for bit against stand-alone oracles, hash buffers included."""
    R, K = 63, 64
    n = (R + 1) ** 2
    p = gpu_params(numSubsteps=5, numIterations=10)
    from velvet_b200.distributed import instance_model_height
    models = [vb.TransformMatrix((0, instance_model_height(i), 1.0), (90, 0, 0), (1, 1, 1)) for i in range(K)]
    v, idx = vb.GenerateClothMesh(R)
    g = vb.VtClothSolverGPU(p)
    g.AddClothInstances(R, v, idx, models, ())
    assert g.iterateKernel == (vb.ITERATE_GRID if iterate_kernel_param == "grid" else vb.ITERATE_TILES)
    sample = [0, 1, 31, 32, 63]
    all_oracles = _oracles(R, p, [models[0]] + [models[i] for i in sample[1:]], [], shared=True)
    oracles = dict(zip(sample, all_oracles))
    cols = vb.sphere_plane_colliders()
    g.UpdateColliders(cols)
    for o in oracles.values():
        o.set_colliders([to_o1_collider(c) for c in cols])
    for _ in range(3):
        g.Simulate()
        for o in oracles.values():
            o.simulate()
    pos = g.download("positions").reshape(K, n, 3)
    nrm = g.download("normals").reshape(K, n, 3)
    ph, pi = g.download("particleHash"), g.download("particleIndex")
    nb = valid_prefix_table(g.download("neighbors"), K * n, 64)
    for i, o in oracles.items():
        assert np.array_equal(pos[i].reshape(-1), o.buffer("positions")), f"instance {i}"
        assert np.array_equal(nrm[i].reshape(-1), o.buffer("normals")), f"instance {i}"
        assert np.array_equal(ph[i * n:(i + 1) * n], o.buffer("particleHash") + i * 2 * n)
        assert np.array_equal(pi[i * n:(i + 1) * n], o.buffer("particleIndex") + i * n)
        ref = valid_prefix_table(o.buffer("neighbors"), n, 64).astype(np.int64)
        ref = np.where(ref == 0xFFFFFFFF, 0xFFFFFFFF, ref + i * n)
        assert np.array_equal(nb[:, i * n:(i + 1) * n].astype(np.int64), ref), f"neighbour lists of instance {i}"
    # instances 0 and 32 share a model height (k mod 32): identical cloths must stay identical
    assert np.array_equal(pos[0], pos[32])


def test_instances_do_not_interact_and_errors():
    R, K = 16, 4
    n = (R + 1) ** 2
    p = gpu_params()
    v, idx = vb.GenerateClothMesh(R)
    same = [vb.TransformMatrix((0, 1.5, 1.0), (90, 0, 0), (1, 1, 1))] * K  # four cloths in exactly the same place
    g = vb.VtClothSolverGPU(p)
    g.AddClothInstances(R, v, idx, same)
    g.UpdateColliders(vb.sphere_plane_colliders())
    for _ in range(15):
        g.Simulate()
    pos = g.download("positions").reshape(K, n, 3)
    for i in range(1, K):
        assert np.array_equal(pos[i], pos[0]), "coincident instances stay coincident: no cross-instance collisions"
    nb = valid_prefix_table(g.download("neighbors"), K * n, 64).astype(np.int64)
    for i in range(K):
        cols = nb[:, i * n:(i + 1) * n]
        valid = cols != 0xFFFFFFFF
        assert np.all((cols[valid] >= i * n) & (cols[valid] < (i + 1) * n))
    with pytest.raises(vb.VelvetError):
        g.AddCloth(v, idx, same[0], 0.1)  # instanced solvers take no further registration
    g2 = vb.VtClothSolverGPU(p, pipeline=vb.PIPELINE_SEAM)
    g2.AddClothInstances(R, v, idx, same)
    with pytest.raises(vb.VelvetError):
        g2.Simulate()  # the reference-order pipeline has no notion of independent instances
