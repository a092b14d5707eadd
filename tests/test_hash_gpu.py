"""GPU parity of the spatial hash and of the hand-written radix sort: integer work, bit-exact."""
import ctypes as C

import numpy as np
import pytest

import velvet_b200 as vb
from velvet_b200 import seam
from oracle import o1

from util import valid_prefix_table

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
f = lambda a: a.ctypes.data_as(C.c_void_p)


def dev(a):
    return torch.from_numpy(np.array(a, copy=True)).cuda()


def host(t):
    torch.cuda.synchronize()
    return t.cpu().numpy()


@pytest.mark.parametrize("n,end_bit", [(1, 8), (33, 5), (1000, 11), (4097, 13), (65536, 17), (300000, 20),
                                       (1 << 20, 21), (2500000, 23), (1 << 20, 32)])
def test_sort_pairs_is_a_stable_sort(n, end_bit):
    rng = np.random.default_rng(n + end_bit)
    keys = rng.integers(0, 1 << min(end_bit, 31), n, dtype=np.int64).astype(np.uint32)
    if end_bit == 32:
        keys = rng.integers(0, 1 << 32, n, dtype=np.int64).astype(np.uint32)
    keys[: n // 3] = keys[0]  # long runs of equal keys exercise stability
    rng.shuffle(keys)
    vals = np.arange(n, dtype=np.uint32)
    order = np.argsort(keys, kind="stable")
    dk, dv = dev(keys), dev(vals)
    seam.SortPairs(dk, dv, n, end_bit)
    assert np.array_equal(host(dk), keys[order])
    assert np.array_equal(host(dv), vals[order])


def test_sort_ignores_bits_above_end_bit():
    rng = np.random.default_rng(0)
    n = 50000
    keys = rng.integers(0, 1 << 20, n, dtype=np.int64).astype(np.uint32)
    vals = np.arange(n, dtype=np.uint32)
    order = np.argsort(keys & 0xFFF, kind="stable")
    dk, dv = dev(keys), dev(vals)
    seam.SortPairs(dk, dv, n, 12)
    assert np.array_equal(host(dv), vals[order])


def _hash_both(pos, init, D, k=64):
    n = len(pos)
    cell = np.float32(np.float32(D) * np.float32(1.5))
    args = (n, k, cell, np.float32(cell * cell), 2 * n, np.float32(np.float32(D) * np.float32(D)))
    ph, pi, cs, ce, nb = (np.zeros(m, np.uint32) for m in (n, n, 2 * n, 2 * n, k * n))
    o1.lib().o1_hash_objects(f(ph), f(pi), f(cs), f(ce), f(nb), f(pos), f(init), o1.HashParams(*args))
    d = [dev(np.zeros(m, np.uint32)) for m in (n, n, 2 * n, 2 * n, k * n)]
    seam.HashObjects(*d, dev(pos), dev(init), vb.VtHashParams(*args))
    g = [host(t) for t in d]
    assert np.array_equal(g[0], ph), "sorted cell keys"
    assert np.array_equal(g[1], pi), "sorted particle order"
    assert np.array_equal(g[2], cs), "cellStart"
    valid = cs != 0xFFFFFFFF
    assert np.array_equal(g[3][valid], ce[valid]), "cellEnd"
    assert np.array_equal(valid_prefix_table(g[4], n, k), valid_prefix_table(nb, n, k)), "neighbor lists"
    return valid_prefix_table(nb, n, k)


@pytest.mark.parametrize("R", [31, 63, 255])
def test_hash_flat_grid_cloth(R):
    v, _ = o1.generate_cloth_mesh(R)
    M = o1.transform_matrix((0, 1.5, 1), (90, 0, 0), (1, 1, 1))
    pos = v.copy().reshape(-1)
    o1.lib().o1_initialize_positions(f(pos), 0, len(v), f(M))
    pos = pos.reshape(-1, 3)
    D = np.float32(np.linalg.norm(v[0] - v[1]).astype(np.float32) * np.float32(1.5))
    tab = _hash_both(pos, pos.copy(), D)
    cnt = (tab != 0xFFFFFFFF).sum(0)
    assert 10 < cnt.mean() < 13


def test_hash_random_cloud_with_negative_coordinates():
    rng = np.random.default_rng(3)
    n = 20000
    init = rng.uniform(-3, 3, (n, 3)).astype(np.float32)
    pos = (init + rng.normal(0, 0.1, (n, 3))).astype(np.float32)
    _hash_both(pos, init, np.float32(0.2))


def test_hash_dense_cluster_overflows_the_64_entry_caps():
    # > 64 particles per bucket and > 64 accepted neighbours: bucket scan truncated at 64 entries
    # (SpatialHashGPU.cu L108), list truncated at 64 with no terminator (L120)
    rng = np.random.default_rng(4)
    n = 3000
    init = rng.uniform(-5, 5, (n, 3)).astype(np.float32)
    pos = rng.uniform(0, 0.3, (n, 3)).astype(np.float32)
    tab = _hash_both(pos, init, np.float32(0.2))
    assert ((tab != 0xFFFFFFFF).sum(0) == 64).any()


def test_hash_small_neighbor_cap_and_single_particle():
    rng = np.random.default_rng(5)
    init = rng.uniform(-1, 1, (500, 3)).astype(np.float32)
    pos = rng.uniform(0, 0.5, (500, 3)).astype(np.float32)
    _hash_both(pos, init, np.float32(0.1), k=8)
    one = np.zeros((1, 3), np.float32)
    _hash_both(one, one.copy(), np.float32(0.1))


def test_spatial_hash_object_matches_oracle():
    rng = np.random.default_rng(6)
    n = 5000
    init = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    pos = (init + rng.normal(0, 0.05, (n, 3))).astype(np.float32)
    D = np.float32(0.08)
    h = vb.SpatialHashGPU(D, n)
    h.SetInitialPositions(init)
    h.Hash(pos)
    cell = np.float32(D * np.float32(1.5))
    ph, pi, cs, ce, nb = (np.zeros(m, np.uint32) for m in (n, n, 2 * n, 2 * n, 64 * n))
    o1.lib().o1_hash_objects(f(ph), f(pi), f(cs), f(ce), f(nb), f(pos), f(init),
                             o1.HashParams(n, 64, cell, np.float32(cell * cell), 2 * n, np.float32(D * D)))
    assert np.array_equal(h.download("particleHash"), ph)
    assert np.array_equal(h.download("particleIndex"), pi)
    assert np.array_equal(valid_prefix_table(h.download("neighbors"), n, 64), valid_prefix_table(nb, n, 64))


# ---------------------------------------------------------------- the fused pipeline's own hash kernels on arbitrary inputs
def _hash_fused_vs_oracle(pos, init, R, k=64):
    """Drives hash_particles / onesweep sort / find_cell_start / reorder + tag-filtered neighbour cache (the kernels
    Simulate() runs) through velvet_solver_hash_fused on a solver of (R+1)^2 particles and compares every buffer with O1."""
    n = (R + 1) ** 2
    assert len(pos) == n
    p = vb.default_params()
    p.maxNumNeighbors = k
    g = vb.build_scene(R, p)
    D = np.float32(g.simParams.particleDiameter)
    g.HashFused()  # builds the fused resources (tile plan from the regular grid) before the buffers are overwritten
    g.upload("initialPositions", init.reshape(-1))
    g.upload("predicted", pos.reshape(-1))
    g.HashFused()
    cell = np.float32(D * np.float32(p.hashCellSizeScalar))
    args = (n, k, cell, np.float32(cell * cell), 2 * n, np.float32(D * D))
    ph, pi, cs, ce, nb = (np.zeros(m, np.uint32) for m in (n, n, 2 * n, 2 * n, k * n))
    o1.lib().o1_hash_objects(f(ph), f(pi), f(cs), f(ce), f(nb), f(pos), f(init), o1.HashParams(*args))
    assert np.array_equal(g.download("particleHash"), ph), "sorted cell keys"
    assert np.array_equal(g.download("particleIndex"), pi), "sorted particle order"
    assert np.array_equal(g.download("cellStart"), cs), "cellStart"
    valid = cs != 0xFFFFFFFF
    assert np.array_equal(g.download("cellEnd")[valid], ce[valid]), "cellEnd"
    tab = valid_prefix_table(nb, n, k)
    assert np.array_equal(valid_prefix_table(g.download("neighbors"), n, k), tab), "neighbor lists"
    return tab, float(D)


def test_fused_hash_on_a_crumpled_cloud_and_far_from_the_origin():
    """The neighbour cache of the fused pipeline rejects candidates by a clamped 10-bit cell tag before the float tests.
    Inputs that stress it: negative coordinates, a cloud hundreds of cells wide (tag coordinates clamp at +-256 cells),
    clusters thousands of cells from the origin (every tag clamped: the filter must degrade to 'pass', never to 'reject')."""
    R = 99
    n = (R + 1) ** 2
    rng = np.random.default_rng(21)
    h = 2.0 / R
    D = 1.5 * h
    cell = 1.5 * D
    # (a) crumpled: particles scattered in a box ~14 cells wide around the origin, registration positions unrelated
    init = rng.uniform(-3, 3, (n, 3)).astype(np.float32)
    pos = rng.uniform(-7 * cell, 7 * cell, (n, 3)).astype(np.float32)
    tab, _ = _hash_fused_vs_oracle(pos, init, R)
    assert (tab != 0xFFFFFFFF).sum() > n
    # (b) a sheet of blobs spanning ~1400 cells in x: tags clamp on both sides
    centers = np.stack([np.linspace(-700 * cell, 700 * cell, 50), np.zeros(50), np.zeros(50)], 1)
    pos = (centers[rng.integers(0, 50, n)] + rng.normal(0, 1.2 * cell, (n, 3))).astype(np.float32)
    tab, _ = _hash_fused_vs_oracle(pos, init, R)
    assert (tab != 0xFFFFFFFF).sum() > n
    # (c) everything far outside the tag range, in all octants
    centers = rng.choice([-1.0, 1.0], (40, 3)) * rng.uniform(3000 * cell, 9000 * cell, (40, 3))
    pos = (centers[rng.integers(0, 40, n)] + rng.normal(0, 1.5 * cell, (n, 3))).astype(np.float32)
    tab, _ = _hash_fused_vs_oracle(pos, init, R)
    assert (tab != 0xFFFFFFFF).sum() > n


def test_fused_hash_dense_cluster_and_small_caps():
    """> 64 particles per bucket and > 64 accepted neighbours (bucket scan and list caps, SpatialHashGPU.cu L108 / L120),
    and a small maxNumNeighbors."""
    R = 49
    n = (R + 1) ** 2
    rng = np.random.default_rng(22)
    h = 2.0 / R
    cell = 2.25 * h
    init = rng.uniform(-5, 5, (n, 3)).astype(np.float32)
    pos = rng.uniform(0, 1.5 * cell, (n, 3)).astype(np.float32)
    tab, _ = _hash_fused_vs_oracle(pos, init, R)
    assert ((tab != 0xFFFFFFFF).sum(0) == 64).any()
    tab, _ = _hash_fused_vs_oracle(pos, init, R, k=7)
    assert ((tab != 0xFFFFFFFF).sum(0) == 7).any()


def test_fused_hash_at_headline_size_matches_reference_kernel_digests():
    """1,048,576 particles (BASELINE configs[2]): the fused pipeline's hash stage against SHA-256 digests of the REFERENCE's
    own kernels' output (tests/golden/refcuda_hash_R1023_digest.npz, produced on a B200 by tests/golden/make_golden.py) and
    the full neighbour lists of 4,096 sampled particles."""
    import hashlib
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", "refcuda_hash_R1023_digest.npz")
    gold = np.load(path)
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    n = 1 << 20
    g = vb.build_scene(1023, vb.default_params())
    assert np.float32(g.simParams.particleDiameter) == gold["particleDiameter"]
    g.HashFused()  # builds the fused resources before the buffers are overwritten
    rng = np.random.default_rng(2024)
    init = g.download("initialPositions").reshape(-1).copy()
    pred = (g.download("positions").reshape(-1) + rng.normal(0, 0.0009, 3 * n)).astype(np.float32)
    assert sha(np.concatenate([init, pred])) == str(gold["inputs_digest"]), "the seeded inputs are not the fixture's"
    g.upload("predicted", pred)
    g.HashFused()
    assert sha(g.download("particleHash")) == str(gold["particleHash"])
    assert sha(g.download("particleIndex")) == str(gold["particleIndex"])
    cs = g.download("cellStart").copy()
    assert sha(cs) == str(gold["cellStart"])
    ce = g.download("cellEnd").copy()
    ce[cs == 0xFFFFFFFF] = 0
    assert sha(ce) == str(gold["cellEnd"])
    tab = valid_prefix_table(g.download("neighbors"), n, 64)
    assert np.array_equal(tab[:, gold["sample"]], gold["sample_neighbors"])
    assert sha(tab) == str(gold["neighbors"])


def test_band_ordered_candidate_walk_builds_the_same_lists(monkeypatch):
    """Cloths of more than 3 M particles walk their sorted slots band by band (bands of 2^18 consecutive particle indices,
    hash_kernels.cuh: band_slots_kernel) so that the candidates of a band stay in L2.  Forced here at a small size, with
    bands that do and do not divide the particle count: lists and frames must equal the unbanded walk's bit for bit."""
    import velvet_b200 as vb
    from util import gpu_params

    def run(band):
        if band is None:
            monkeypatch.delenv("VELVET_WALK_BAND", raising=False)
        else:
            monkeypatch.setenv("VELVET_WALK_BAND", str(band))
        g = vb.build_scene(90, gpu_params(numSubsteps=3, numIterations=4))
        g.UpdateColliders(vb.sphere_plane_colliders())
        for _ in range(10):
            g.Simulate()
        out = (g.download("neighbors").copy(), g.download("positions").copy(), g.download("cellStart").copy())
        g.close()
        return out

    ref = run(0)
    for band in (1000, 4096, 8281, 100000):   # 91 * 91 = 8281 particles: 9 bands, 3 bands (the last one partial), 1 band, 1 oversized band
        got = run(band)
        for a, b, name in zip(ref, got, ("neighbors", "positions", "cellStart")):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (band, name)
