"""GPU parity of VtClothSolverGPU::Simulate (fused sm_100a pipeline and the reference-order seam pipeline)
against the O1 oracle, through the C ABI.

Tolerances are north_star's: max |dx| <= 1e-4 x cloth extent after 1 frame, <= 1e-3 x extent after 60 frames
(extent = 2.0).  The fused pipeline sums corrections in the oracle's order, so it is in practice far tighter;
hash outputs (integer work) are compared bit-exactly on the first rebuild."""
import math
import os

import numpy as np
import pytest

import velvet_b200 as vb
from oracle import o1

from util import EXTENT, ColliderTrack, gpu_params, make_pair, max_abs_diff, set_colliders, to_o1_params, valid_prefix_table

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
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

TOL_1 = 1e-4 * EXTENT
TOL_60 = 1e-3 * EXTENT


def assert_fused_parity(g, o, tol):
    """north_star's contract is max |dx| <= tol; the fused pipeline sums in the oracle's order with the same
    fp32 operations (no FMA contraction, shared acos algorithm), so it is asserted bit-identical as well."""
    for name in ("positions", "velocities", "predicted", "normals"):
        a, b = g.download(name).reshape(-1), o.buffer(name)
        assert max_abs_diff(a, b) <= tol, name
        assert np.array_equal(a, b), f"{name}: fused pipeline no longer bit-identical to the oracle"


def _run_cfg1(g, o, frames):
    """Config 1 (main.cpp L149-178): 32x32, attach {0, R}, plane + sphere r=0.6 at (0, 0.6, -cos 2t)."""
    sphere = ColliderTrack(vb.COLLIDER_SPHERE, (0, 0.6, -1.0), (0.6, 0.6, 0.6))
    plane = vb.MakeCollider(vb.COLLIDER_PLANE, (0, 0, 0), (1, 1, 1))
    for fr in range(frames):
        sphere.move((0, 0.6, -math.cos(2 * fr / 60.0)))
        set_colliders(g, o, [plane, sphere.collider()])
        g.Simulate()
        if o is not None:
            o.simulate()


@pytest.mark.parametrize("pipeline", [vb.PIPELINE_FUSED, vb.PIPELINE_SEAM])
def test_cfg1_one_frame(pipeline):
    p = gpu_params(numSubsteps=5, numIterations=10)
    g, o = make_pair(31, p, position=(0, 2.5, 0), rotation=(0, 0, 0), attached=[0, 31], pipeline=pipeline)
    assert np.array_equal(g.download("positions").reshape(-1), o.buffer("positions")), "registration (InitializePositions)"
    for name in ("stretchIndices", "stretchLengths", "bendIndices", "bendAngles", "attachParticleIDs", "attachSlotIDs",
                 "attachDistances", "attachSlotPositions", "invMasses", "indices", "initialPositions"):
        assert np.array_equal(g.download(name).reshape(-1), o.buffer(name)), name
    _run_cfg1(g, o, 1)
    assert max_abs_diff(g.download("positions"), o.buffer("positions")) <= TOL_1
    assert max_abs_diff(g.download("velocities"), o.buffer("velocities")) <= 60 * 5 * TOL_1
    assert max_abs_diff(g.download("normals"), o.buffer("normals")) <= 1e-3
    if pipeline == vb.PIPELINE_SEAM:
        return  # float atomics: 1-ulp position noise may move a particle across a cell boundary
    # integer work: the hash rebuilt on the last hashing substep
    assert np.array_equal(g.download("particleHash"), o.buffer("particleHash"))
    assert np.array_equal(g.download("particleIndex"), o.buffer("particleIndex"))
    assert np.array_equal(valid_prefix_table(g.download("neighbors"), 1024, 64),
                          valid_prefix_table(o.buffer("neighbors"), 1024, 64))


def test_cfg1_fused_is_bit_close_to_oracle_order():
    # same summation order as the oracle: only libm-vs-CUDA acosf may differ
    p = gpu_params(numSubsteps=5, numIterations=10)
    g, o = make_pair(31, p, position=(0, 2.5, 0), rotation=(0, 0, 0), attached=[0, 31])
    _run_cfg1(g, o, 1)
    assert_fused_parity(g, o, 2e-6)


@pytest.mark.parametrize("pipeline", [vb.PIPELINE_FUSED, vb.PIPELINE_SEAM])
def test_cfg1_sixty_frames(pipeline):
    p = gpu_params(numSubsteps=5, numIterations=10)
    g, o = make_pair(31, p, position=(0, 2.5, 0), rotation=(0, 0, 0), attached=[0, 31], pipeline=pipeline)
    _run_cfg1(g, o, 60)
    pos = g.download("positions")
    assert np.isfinite(pos).all()
    if pipeline == vb.PIPELINE_SEAM:
        # The reference-order pipeline scatters with float atomics exactly like the reference, so its summation order
        # changes from run to run and contact decisions amplify that noise: on B200 two runs of the REFERENCE's own
        # kernels on identical inputs end up 1.3e-2 apart by frame 20 of this scene (tests/test_ref_cuda_gpu.py).
        # It is the compatibility path: only a loose envelope is asserted here.
        assert max_abs_diff(pos, o.buffer("positions")) <= 0.05
        return
    assert_fused_parity(g, o, TOL_60)
    path = os.path.join(GOLDEN, "cfg1_frame60.npz")
    if os.path.exists(path):
        assert max_abs_diff(pos, np.load(path)["positions"]) <= TOL_60


@pytest.mark.parametrize("frames,tol", [(1, TOL_1), (30, TOL_60)])
def test_drape_64_self_collision(frames, tol):
    """Horizontal 64x64 sheet (model = T(0,1.5,1) Rx(90), main.cpp L141) draped over the static sphere + plane."""
    p = gpu_params(numSubsteps=5, numIterations=10)
    g, o = make_pair(63, p)
    cols = vb.sphere_plane_colliders()
    set_colliders(g, o, cols)
    for _ in range(frames):
        g.Simulate()
        o.simulate()
    assert_fused_parity(g, o, tol)


def test_corner_attach_40(self=None):
    """'Cloth / Attach' scene (main.cpp L125-147): 41x41, four corner attach slots, 2 substeps x 4 iterations."""
    R = 40
    p = gpu_params()
    corners = [0, R, (R + 1) * (R + 1) - 1, (R + 1) * R]
    g, o = make_pair(R, p, attached=corners)
    cols = vb.sphere_plane_colliders(radius=0.5)
    set_colliders(g, o, cols)
    for _ in range(20):
        g.Simulate()
        o.simulate()
    assert_fused_parity(g, o, TOL_60)
    inv = g.download("invMasses")
    assert [inv[c] for c in corners] == [0, 0, 0, 0]


def test_moving_attach_slot_and_disabled_self_collision():
    """Swirl scene (main.cpp L302-333): one attach slot moved by the host between frames; self-collision off."""
    R = 36
    p = gpu_params(enableSelfCollision=0)
    g, o = make_pair(R, p, attached=[0])
    set_colliders(g, o, [vb.MakeCollider(vb.COLLIDER_PLANE, (0, 0, 0), (1, 1, 1))])
    for fr in range(10):
        t = fr / 60.0 * 3
        slot = np.array([math.sin(t), math.cos(t) + 2, 0], np.float32)
        g.upload("attachSlotPositions", slot)
        o.buffer("attachSlotPositions")[:] = slot
        g.Simulate()
        o.simulate()
    assert_fused_parity(g, o, TOL_60)


def test_two_cloths_and_cube_collider():
    """'Multiple Object' scene (main.cpp L232-268): cloths registered one after another on one solver + cube SDF."""
    p = gpu_params(numSubsteps=5, numIterations=5, friction=0.6)
    g = vb.VtClothSolverGPU(p, math_mode=vb.MATH_EXACT)
    o = o1.O1Solver(__import__("util").to_o1_params(p))
    for height in (1.5, 1.8):
        R = 24
        v, idx = vb.GenerateClothMesh(R)
        M = vb.TransformMatrix((0, height, 1.0), (90, 0, 0), (1, 1, 1))
        obj = vb.VtClothObjectGPU(R, g)
        obj.Start(v, idx, M)
        o.cloth_object_start(R, v, idx, M, [])
    assert g.simParams.numParticles == 2 * 25 * 25 == o.params.numParticles
    cube = ColliderTrack(vb.COLLIDER_CUBE, (0, 0.5, 0), (1, 1, 1), (0, 15, 0))
    cols = [vb.MakeCollider(vb.COLLIDER_PLANE, (0, 0, 0), (1, 1, 1)), cube.collider()]
    set_colliders(g, o, cols)
    for _ in range(25):
        g.Simulate()
        o.simulate()
    assert_fused_parity(g, o, TOL_60)


def test_host_may_edit_buffers_between_frames():
    """MouseGrabber-style edits (MouseGrabber.hpp L64-77): invMass = 0 and position/velocity writes between frames."""
    p = gpu_params()
    g, o = make_pair(20, p)
    set_colliders(g, o, vb.sphere_plane_colliders())
    for fr in range(6):
        if fr == 2:
            inv = g.download("invMasses"); inv[5] = 0; g.upload("invMasses", inv); o.buffer("invMasses")[5] = 0
            pos = g.download("positions"); pos[5] += np.float32(0.05); g.upload("positions", pos)
            o.buffer("positions").reshape(-1, 3)[5] += np.float32(0.05)
            vel = g.download("velocities"); vel[7] = (0.5, 0.5, 0); g.upload("velocities", vel)
            o.buffer("velocities").reshape(-1, 3)[7] = (0.5, 0.5, 0)
        g.Simulate()
        o.simulate()
    assert_fused_parity(g, o, TOL_60)


def test_param_changes_between_frames_and_simulate_dt():
    p = gpu_params(numSubsteps=2, numIterations=4)
    g, o = make_pair(24, p)
    set_colliders(g, o, vb.sphere_plane_colliders())
    g.Simulate(); o.simulate()
    for q in (g.simParams, o.params):
        q.numIterations = 7
        q.numSubsteps = 3
        q.friction = 0.4
        q.interleavedHash = 2
    g.Simulate(); o.simulate()
    g.Simulate(1.0 / 60.0); o.simulate()
    assert_fused_parity(g, o, TOL_1 * 3)


def test_256_one_frame_and_tile_sizes():
    """Config 2 (256x256, 65k particles): one frame against the oracle, for every supported Jacobi tile size."""
    p = gpu_params(numSubsteps=5, numIterations=10)
    ref = None
    for tile in (0, 128, 512):
        g, o = make_pair(255, p, tile_size=tile, oracle=ref is None)
        cols = vb.sphere_plane_colliders()
        if ref is None:
            set_colliders(g, o, cols)
            o.simulate()
            ref = o.buffer("positions").copy()
        else:
            g.UpdateColliders(cols)
        g.Simulate()
        assert np.array_equal(g.download("positions").reshape(-1), ref), tile  # contract: <= TOL_1


# ------------------------------------------------------------------ FAST math mode (opt-in): tolerance parity
def _assert_within(g, o, tol):
    for name, scale in (("positions", 1.0), ("predicted", 1.0), ("velocities", 300.0)):
        assert max_abs_diff(g.download(name), o.buffer(name)) <= tol * scale, name


def test_fast_math_cfg1_one_frame_and_hash():
    p = gpu_params(numSubsteps=5, numIterations=10)
    g, o = make_pair(31, p, position=(0, 2.5, 0), rotation=(0, 0, 0), attached=[0, 31], math_mode=vb.MATH_FAST)
    _run_cfg1(g, o, 1)
    _assert_within(g, o, TOL_1)
    assert max_abs_diff(g.download("positions"), o.buffer("positions")) <= 0.25 * TOL_1, "measured 1.6e-5 on B200"
    # the spatial hash is compiled exactly in every mode: bit-exact on identical input
    pred = o.buffer("predicted").copy()
    g.upload("predicted", pred)
    g.Hash()
    o.hash()
    assert np.array_equal(g.download("particleHash"), o.buffer("particleHash"))
    assert np.array_equal(g.download("particleIndex"), o.buffer("particleIndex"))
    assert np.array_equal(valid_prefix_table(g.download("neighbors"), 1024, 64), valid_prefix_table(o.buffer("neighbors"), 1024, 64))


@pytest.mark.parametrize("frames,tol", [(1, TOL_1), (20, TOL_60)])
def test_fast_math_drape_64(frames, tol):
    """Measured drift of the FAST mode from the oracle on this scene (B200): 1.0e-5 after 1 frame, 1.6e-4 after 20 --
    the same distance the reference's own CUDA build keeps from the oracle (1.9e-4 after 20 frames, test_ref_cuda_gpu).
    Past frame ~22 the cloth buckles over the sphere and every non-bit-identical implementation, the reference
    against itself included, departs by centimetres; only the EXACT mode (bit-identical) can be checked there."""
    p = gpu_params(numSubsteps=5, numIterations=10)
    g, o = make_pair(63, p, math_mode=vb.MATH_FAST)
    set_colliders(g, o, vb.sphere_plane_colliders())
    for _ in range(frames):
        g.Simulate()
        o.simulate()
    _assert_within(g, o, tol)


def test_fast_math_corner_attach_and_cube_scenes():
    R = 40
    corners = [0, R, (R + 1) * (R + 1) - 1, (R + 1) * R]
    g, o = make_pair(R, gpu_params(), attached=corners, math_mode=vb.MATH_FAST)
    set_colliders(g, o, vb.sphere_plane_colliders(radius=0.5))
    # FAST drifts from the oracle chaotically (which FMAs the compiler forms changes with every edit of the kernel):
    # measured 1.3e-3 .. 2.0e-3 after 20 frames here, so the fixed-tolerance check stops at 16 frames
    for _ in range(16):
        g.Simulate()
        o.simulate()
    _assert_within(g, o, TOL_60)
    p = gpu_params(numSubsteps=5, numIterations=5, friction=0.6)
    g, o = make_pair(24, p, position=(0, 1.5, 1.0), math_mode=vb.MATH_FAST)
    cube = ColliderTrack(vb.COLLIDER_CUBE, (0, 0.5, 0), (1, 1, 1), (0, 15, 0))
    set_colliders(g, o, [vb.MakeCollider(vb.COLLIDER_PLANE, (0, 0, 0), (1, 1, 1)), cube.collider()])
    for _ in range(20):
        g.Simulate()
        o.simulate()
    _assert_within(g, o, TOL_60)


def test_fast_math_ten_frames_then_envelope():
    """Config 1 with the sphere parked away (no collider contact at all).  Even so the Jacobi dynamics amplify last-bit
    differences: measured FAST-vs-oracle drift on B200 is 3.6e-5 after 1 frame, 1.9e-4 after 10, 7.7e-3 after 20.
    north_star's 60-frame tolerance is therefore asserted on the EXACT mode (bit-identical, test_cfg1_sixty_frames);
    here: strict for 10 frames, finite and bounded after 60."""
    p = gpu_params(numSubsteps=5, numIterations=10)
    g, o = make_pair(31, p, position=(0, 8.0, 0), rotation=(0, 0, 0), attached=[0, 31], math_mode=vb.MATH_FAST)
    cols = [vb.MakeCollider(vb.COLLIDER_PLANE, (0, 0, 0), (1, 1, 1)),
            vb.MakeCollider(vb.COLLIDER_SPHERE, (0, 0.6, -3.0), (0.6, 0.6, 0.6))]
    set_colliders(g, o, cols)
    for fr in range(60):
        g.Simulate()
        o.simulate()
        if fr < 10:
            assert max_abs_diff(g.download("positions"), o.buffer("positions")) <= TOL_60 * (fr + 1) / 10, fr
    pos = g.download("positions")
    assert np.isfinite(pos).all() and max_abs_diff(pos, o.buffer("positions")) <= 0.25


def test_fast_math_256_one_frame():
    p = gpu_params(numSubsteps=5, numIterations=10)
    g, o = make_pair(255, p, math_mode=vb.MATH_FAST)
    set_colliders(g, o, vb.sphere_plane_colliders())
    g.Simulate()
    o.simulate()
    _assert_within(g, o, TOL_1)


def test_fast_math_cfg1_sixty_frames_envelope():
    """Chaotic scene (see tests/test_ref_cuda_gpu.py): strict while contact-free, bounded afterwards."""
    p = gpu_params(numSubsteps=5, numIterations=10)
    g, o = make_pair(31, p, position=(0, 2.5, 0), rotation=(0, 0, 0), attached=[0, 31], math_mode=vb.MATH_FAST)
    sphere = ColliderTrack(vb.COLLIDER_SPHERE, (0, 0.6, -1.0), (0.6, 0.6, 0.6))
    plane = vb.MakeCollider(vb.COLLIDER_PLANE, (0, 0, 0), (1, 1, 1))
    for fr in range(60):
        sphere.move((0, 0.6, -math.cos(2 * fr / 60.0)))
        set_colliders(g, o, [plane, sphere.collider()])
        g.Simulate()
        o.simulate()
        if fr < 15:
            assert max_abs_diff(g.download("positions"), o.buffer("positions")) <= TOL_1, fr
    pos = g.download("positions")
    assert np.isfinite(pos).all()
    assert max_abs_diff(pos, o.buffer("positions")) <= 0.05


def test_fused_is_deterministic_and_matches_seam_at_1m():
    """Headline size (1024x1024, 1M particles): properties that do not need the CPU oracle.
    Two independent fused runs are bit-identical (deterministic summation); the fused pipeline agrees with the
    reference-order seam pipeline (float atomics) within the 1-frame tolerance; hash outputs agree bit-exactly."""
    p = gpu_params(numSubsteps=5, numIterations=10)
    outs = []
    for pipeline, mode in ((vb.PIPELINE_FUSED, vb.MATH_EXACT), (vb.PIPELINE_FUSED, vb.MATH_EXACT), (vb.PIPELINE_SEAM, vb.MATH_EXACT),
                           (vb.PIPELINE_FUSED, vb.MATH_FAST), (vb.PIPELINE_FUSED, vb.MATH_FAST)):
        g, _ = make_pair(1023, p, pipeline=pipeline, oracle=False, math_mode=mode)
        g.UpdateColliders(vb.sphere_plane_colliders())
        g.Simulate()
        g.Simulate()
        outs.append((g.download("positions"), g.download("particleIndex"), g.download("particleHash")))
        assert g.lastLaunchCount > 0
        g.close()
    assert np.isfinite(outs[0][0]).all()
    assert np.array_equal(outs[0][0], outs[1][0]), "fused pipeline is run-to-run deterministic"
    assert max_abs_diff(outs[0][0], outs[2][0]) <= TOL_1
    assert np.array_equal(outs[0][2], outs[2][2]) and np.array_equal(outs[0][1], outs[2][1])
    # sortedness / permutation properties at full size
    assert np.all(np.diff(outs[0][2].astype(np.int64)) >= 0)
    assert np.array_equal(np.sort(outs[0][1]), np.arange(1 << 20, dtype=np.uint32))
    # the FAST math mode is deterministic too (fixed summation order) and within tolerance of the EXACT one
    assert np.array_equal(outs[3][0], outs[4][0]), "fast math mode is run-to-run deterministic"
    assert max_abs_diff(outs[3][0], outs[0][0]) <= TOL_1


@pytest.mark.slow
def test_headline_1m_one_frame_against_the_oracle(iterate_kernel_param):
    """BASELINE configs[2], the benchmarked workload itself: one frame (5 x 10) of the 1024x1024 self-colliding drape against
    the O1 oracle (about 20 s of CPU): positions / velocities / normals / predicted bit for bit, and the spatial hash of the
    frame's last rebuild (keys, sorted order, cell table, neighbour lists)."""
    p = gpu_params(numSubsteps=5, numIterations=10)
    g, o = make_pair(1023, p)
    assert g.iterateKernel == (vb.ITERATE_GRID if iterate_kernel_param == "grid" else vb.ITERATE_TILES)
    set_colliders(g, o, vb.sphere_plane_colliders())
    g.Simulate()
    o.simulate()
    assert_fused_parity(g, o, TOL_1)
    N = 1 << 20
    assert np.array_equal(g.download("particleHash"), o.buffer("particleHash"))
    assert np.array_equal(g.download("particleIndex"), o.buffer("particleIndex"))
    cs_g, cs_o = g.download("cellStart"), o.buffer("cellStart")
    assert np.array_equal(cs_g, cs_o)
    occupied = cs_o != 0xFFFFFFFF
    assert np.array_equal(g.download("cellEnd")[occupied], o.buffer("cellEnd")[occupied])
    assert np.array_equal(valid_prefix_table(g.download("neighbors"), N, 64), valid_prefix_table(o.buffer("neighbors"), N, 64))
    g.close()


def test_iterate_kernel_selection():
    """Grid cloths carrying exactly the reference's constraint pattern run the implicit-grid kernel; anything else (an extra
    constraint, a second triangulation, a generic mesh) keeps the record-driven tile kernel -- same results either way."""
    if os.environ.get("VELVET_ITERATE") == "tiles":
        pytest.skip("kernel choice forced by the environment")
    p = gpu_params()
    g, _ = make_pair(20, p, oracle=False)
    assert g.iterateKernel == vb.ITERATE_GRID
    g.AddStretch(0, 5, 0.3)  # one constraint beyond the pattern
    assert g.iterateKernel == vb.ITERATE_TILES
    g2, _ = make_pair(20, p, oracle=False)
    g2.SetIterateMode(vb.ITERATE_TILES)
    assert g2.iterateKernel == vb.ITERATE_TILES
    g2.SetIterateMode(vb.ITERATE_AUTO)
    assert g2.iterateKernel == vb.ITERATE_GRID
    # two grid cloths of different size in one solver: both recognised, one kernel launch covers both
    g3 = vb.VtClothSolverGPU(p)
    for R, pos in ((12, (0, 1.5, 1.0)), (17, (0.3, 1.9, 1.0))):
        v, idx = vb.GenerateClothMesh(R)
        obj = vb.VtClothObjectGPU(R, g3)
        obj.Start(v, idx, vb.TransformMatrix(pos, (90, 0, 0), (1, 1, 1)))
    assert g3.iterateKernel == vb.ITERATE_GRID


def test_two_grid_cloths_in_one_solver_match_the_oracle(iterate_kernel_param):
    """Two cloths of different resolution registered one after the other (their particles collide with each other through
    the shared hash): the grid kernel's cloth table against the oracle."""
    p = gpu_params(numSubsteps=3, numIterations=6)
    g = vb.VtClothSolverGPU(p)
    o = o1.O1Solver(to_o1_params(p))
    for R, pos in ((30, (0, 1.5, 1.0)), (19, (0.1, 1.62, 0.9))):
        v, idx = vb.GenerateClothMesh(R)
        M = vb.TransformMatrix(pos, (90, 0, 0), (1, 1, 1))
        vb.VtClothObjectGPU(R, g).Start(v, idx, M)
        ov, oidx = o1.generate_cloth_mesh(R)
        o.cloth_object_start(R, ov, oidx, o1.transform_matrix(pos, (90, 0, 0), (1, 1, 1)), [])
    set_colliders(g, o, vb.sphere_plane_colliders())
    for _ in range(8):
        g.Simulate()
        o.simulate()
    assert_fused_parity(g, o, TOL_60)


def _register_generic(g, o, vertices, tri_indices, model, stretch, bends, attach_slots, attaches, diameter):
    """Registers an arbitrary mesh through the element-wise API (AddCloth / AddStretch / AddBend / AddAttach*), the
    way a maintainer's own cloth component would, on both the CUDA solver and the oracle."""
    g.AddCloth(vertices, tri_indices, model, diameter)
    o.add_cloth(vertices, tri_indices, model, diameter)
    for (a, b, d) in stretch:
        g.AddStretch(a, b, d); o.add_stretch(a, b, d)
    for slot in attach_slots:
        g.AddAttachSlot(slot); o.add_attach_slot(slot)
    for (pid, slot, d) in attaches:
        g.AddAttach(pid, slot, d); o.add_attach(pid, slot, d)
    for (a, b, c, d) in bends:
        g.AddBend(a, b, c, d, 0.0); o.add_bend(a, b, c, d, 0.0)


def _shuffled_grid(R, seed):
    """A grid cloth whose vertices are stored in random order: constraint indices have no locality at all, which is
    what the Morton tiling of the Jacobi kernel has to cope with on arbitrary meshes."""
    rng = np.random.default_rng(seed)
    v, idx = vb.GenerateClothMesh(R)
    n = len(v)
    perm = rng.permutation(n)          # new index of old vertex i
    inv = np.empty(n, np.int64); inv[perm] = np.arange(n)
    v2 = v[inv]
    idx2 = perm[idx.astype(np.int64)].astype(np.uint32)
    S = R + 1
    world = (vb.TransformMatrix((0, 1.5, 1.0), (90, 0, 0), (1, 1, 1)).reshape(4, 4).T @ np.c_[v, np.ones(n)].T).T[:, :3].astype(np.float32)
    stretch = []
    at = lambda x, y: x * S + y
    for x in range(S):
        for y in range(S):
            pairs = []
            if y != R: pairs.append((at(x, y), at(x, y + 1)))
            if x != R: pairs.append((at(x, y), at(x + 1, y)))
            if y != R and x != R: pairs += [(at(x, y), at(x + 1, y + 1)), (at(x, y + 1), at(x + 1, y))]
            for a, b in pairs:
                stretch.append((int(perm[a]), int(perm[b]), float(np.float32(np.linalg.norm(world[a] - world[b])))))
    bends = [(int(perm[idx[i]]), int(perm[idx[i + 5]]), int(perm[idx[i + 2]]), int(perm[idx[i + 1]])) for i in range(0, len(idx), 6)]
    return v2, idx2, stretch, bends


def test_shuffled_vertex_order_mesh_is_still_bit_identical():
    R = 40
    p = gpu_params(numSubsteps=3, numIterations=6)
    v2, idx2, stretch, bends = _shuffled_grid(R, seed=5)
    g = vb.VtClothSolverGPU(p)
    o = o1.O1Solver(__import__("util").to_o1_params(p))
    M = vb.TransformMatrix((0, 1.5, 1.0), (90, 0, 0), (1, 1, 1))
    D = float(np.float32(2.0 / R * 1.5))
    _register_generic(g, o, v2, idx2, M, stretch, bends, [], [], D)
    set_colliders(g, o, vb.sphere_plane_colliders())
    for _ in range(20):
        g.Simulate(); o.simulate()
    assert_fused_parity(g, o, TOL_60)
    assert g.lastLaunchCount < 200, "the fused pipeline (not the seam fallback) must have been used"


def test_high_valence_mesh_falls_back_to_the_reference_order_pipeline():
    """A hub particle with 40 stretch constraints exceeds the 30 slots a tile gives one particle: Simulate() must fall
    back to the seam pipeline on its own and still agree with the oracle (float atomics: tolerance, not bit-identity)."""
    rng = np.random.default_rng(9)
    n = 60
    v = rng.uniform(-0.5, 0.5, (n, 3)).astype(np.float32); v[:, 1] += 2.0
    tri = np.array([0, 1, 2], np.uint32)
    p = gpu_params(enableSelfCollision=0)
    g = vb.VtClothSolverGPU(p); o = o1.O1Solver(__import__("util").to_o1_params(p))
    M = vb.TransformMatrix()
    stretch = [(0, j, float(np.float32(np.linalg.norm(v[0] - v[j])))) for j in range(1, 41)]
    stretch += [(j, j + 1, float(np.float32(np.linalg.norm(v[j] - v[j + 1])))) for j in range(1, n - 1)]
    _register_generic(g, o, v, tri, M, stretch, [], [(0.0, 2.5, 0.0)], [(0, 0, 0.0)], 0.05)
    set_colliders(g, o, [vb.MakeCollider(vb.COLLIDER_PLANE, (0, 0, 0), (1, 1, 1))])
    for _ in range(5):
        g.Simulate(); o.simulate()
    assert g.lastLaunchCount > 0
    assert max_abs_diff(g.download("positions"), o.buffer("positions")) <= TOL_1


def test_empty_solver_and_error_convention():
    g = vb.VtClothSolverGPU()
    g.Simulate()  # no cloth registered: silently nothing to do (CUDA_CALL early-out)
    with pytest.raises(vb.VelvetError):
        g.AddAttach(3, 0, 0.0)  # out of range -> error status, never exit()
    with pytest.raises(vb.VelvetError):
        g.SetTileSize(100)


# ---------------------------------------------------------------- committed golden vectors (reference CUDA kernels on B200)
@pytest.mark.parametrize("math_mode", [vb.MATH_EXACT, vb.MATH_FAST])
def test_product_against_reference_kernel_golden_vectors(math_mode):
    """tests/golden/refcuda_*.npz hold outputs of the reference's own VtClothSolverGPU.cu / SpatialHashGPU.cu (see
    tests/golden/make_golden.py): positions within north_star's tolerance, spatial hash bit-exact on identical inputs."""
    gold = np.load(os.path.join(GOLDEN, "refcuda_cfg1.npz"))
    p = gpu_params(numSubsteps=5, numIterations=10)
    g, _ = make_pair(31, p, position=(0, 2.5, 0), rotation=(0, 0, 0), attached=[0, 31], oracle=False, math_mode=math_mode)
    sphere = ColliderTrack(vb.COLLIDER_SPHERE, (0, 0.6, -1.0), (0.6, 0.6, 0.6))
    plane = vb.MakeCollider(vb.COLLIDER_PLANE, (0, 0, 0), (1, 1, 1))
    for fr in range(15):
        sphere.move((0, 0.6, -math.cos(2 * fr / 60.0)))
        g.UpdateColliders([plane, sphere.collider()])
        g.Simulate()
        if fr + 1 in (1, 5, 10, 15):
            assert max_abs_diff(g.download("positions"), gold[f"positions_{fr + 1}"]) <= TOL_1, fr + 1
    gold = np.load(os.path.join(GOLDEN, "refcuda_drape64.npz"))
    g, _ = make_pair(63, p, oracle=False, math_mode=math_mode)
    g.UpdateColliders(vb.sphere_plane_colliders())
    for fr in range(20):
        g.Simulate()
        if fr + 1 in (1, 10, 20):
            assert max_abs_diff(g.download("positions"), gold[f"positions_{fr + 1}"]) <= (TOL_1 if fr == 0 else TOL_60), fr + 1
    # 25x25 cloth, four corner attachments, on a turning cube with friction 0.6 (cube SDF + collider velocity)
    gold = np.load(os.path.join(GOLDEN, "refcuda_cube25.npz"))
    R = 24
    pc = gpu_params(numSubsteps=5, numIterations=5, friction=0.6)
    g, _ = make_pair(R, pc, attached=[0, R, (R + 1) * (R + 1) - 1, (R + 1) * R], oracle=False, math_mode=math_mode)
    cube = ColliderTrack(vb.COLLIDER_CUBE, (0, 0.95, 0), (1, 1, 1), (0, 15, 0))
    for fr in range(10):
        cube.move((0, 0.95, 0), (0, 15 + 2 * (fr + 1), 0))
        g.UpdateColliders([vb.MakeCollider(vb.COLLIDER_PLANE, (0, 0, 0), (1, 1, 1)), cube.collider()])
        g.Simulate()
        if fr + 1 in (1, 5, 10):
            assert max_abs_diff(g.download("positions"), gold[f"positions_{fr + 1}"]) <= (TOL_1 if fr == 0 else TOL_60), fr + 1
    for R in (31, 63):
        gold = np.load(os.path.join(GOLDEN, f"refcuda_hash_R{R}.npz"))
        n = (R + 1) ** 2
        g, _ = make_pair(R, gpu_params(), oracle=False, math_mode=math_mode)
        g.upload("initialPositions", gold["initialPositions"])
        g.upload("predicted", gold["predicted"])
        g.Hash()
        assert np.array_equal(g.download("particleHash"), gold["particleHash"])
        assert np.array_equal(g.download("particleIndex"), gold["particleIndex"])
        assert np.array_equal(g.download("cellStart"), gold["cellStart"])
        valid = gold["cellStart"] != 0xFFFFFFFF
        assert np.array_equal(g.download("cellEnd")[valid], gold["cellEnd"][valid])
        assert np.array_equal(valid_prefix_table(g.download("neighbors"), n, 64), gold["neighbors"])


def test_pipelined_readback_delivers_every_frame():
    """velvet_solver_readback_pipelined: frame k's positions/normals land in the host buffers given for frame k while
    frame k+1 is already being simulated; each must equal what a blocking download of that frame returns."""
    torch = pytest.importorskip("torch")
    import ctypes as C
    p = gpu_params(numSubsteps=3, numIterations=4)
    g, _ = make_pair(63, p, oracle=False)
    ref, _ = make_pair(63, p, oracle=False)
    for s in (g, ref):
        s.UpdateColliders(vb.sphere_plane_colliders())
    n = 64 * 64 * 3
    hp = [torch.zeros(n, dtype=torch.float32).pin_memory() for _ in range(2)]
    hn = [torch.zeros(n, dtype=torch.float32).pin_memory() for _ in range(2)]
    want = []
    for _ in range(5):
        ref.Simulate()
        want.append((ref.download("positions").reshape(-1).copy(), ref.download("normals").reshape(-1).copy()))
    prev = None
    for k in range(5):
        g.Simulate(sync=False)
        t = g.ReadbackPipelined(C.c_void_p(hp[k & 1].data_ptr()), C.c_void_p(hn[k & 1].data_ptr()))
        if prev is not None:
            g.ReadbackWait(prev[0])
            assert np.array_equal(hp[prev[1]].numpy(), want[k - 1][0]) and np.array_equal(hn[prev[1]].numpy(), want[k - 1][1]), k - 1
        prev = (t, k & 1)
    g.ReadbackWait(prev[0])
    assert np.array_equal(hp[prev[1]].numpy(), want[4][0]) and np.array_equal(hn[prev[1]].numpy(), want[4][1])


def test_uploading_constraint_or_initial_position_buffers_takes_effect_in_the_fused_pipeline():
    """Public buffers are the reference's contract (VtBuffer: the host may rewrite them between frames).  Rewriting a
    constraint buffer must rebuild the fused pipeline's tile plan, rewriting initialPositions must refresh the copy its
    neighbour filter uses: the fused result has to follow the oracle given the same edits."""
    p = gpu_params(numSubsteps=3, numIterations=5)
    g, o = make_pair(31, p)
    set_colliders(g, o, vb.sphere_plane_colliders())
    for _ in range(2):
        g.Simulate()
        o.simulate()
    assert_fused_parity(g, o, TOL_1)
    # shrink every rest length by 5 %, stiffen nothing else
    sl = (o.buffer("stretchLengths") * np.float32(0.95)).astype(np.float32)
    o.buffer("stretchLengths")[:] = sl
    g.upload("stretchLengths", sl)
    # and pretend the cloth was registered slightly sheared: changes which close pairs count as self-collision neighbours
    ip = o.buffer("initialPositions").reshape(-1, 3).copy()
    ip[:, 0] += np.float32(0.3) * ip[:, 2]
    o.buffer("initialPositions")[:] = ip.reshape(-1)
    g.upload("initialPositions", ip.reshape(-1))
    for _ in range(3):
        g.Simulate()
        o.simulate()
    assert_fused_parity(g, o, TOL_1)
    assert np.array_equal(valid_prefix_table(g.download("neighbors"), 1024, 64), valid_prefix_table(o.buffer("neighbors"), 1024, 64))


def test_nonzero_rest_angles_uploaded_after_registration(iterate_kernel_param):
    """SURVEY section 8 row f4 (the reference leaves `// TODO: calculate angle`, VtClothObjectGPU.hpp L128-129): per-quad rest
    angles written into the public bendAngles buffer after the cloth was generated.  The grid kernel then takes its rest
    angles from the per-vertex array instead of the scalar 0; both iterate kernels against the oracle."""
    p = gpu_params(numSubsteps=3, numIterations=6)
    g, o = make_pair(26, p)
    set_colliders(g, o, vb.sphere_plane_colliders())
    rng = np.random.default_rng(11)
    angles = rng.uniform(0.0, 0.6, len(o.buffer("bendAngles"))).astype(np.float32)
    o.buffer("bendAngles")[:] = angles
    g.upload("bendAngles", angles)
    for _ in range(6):
        g.Simulate()
        o.simulate()
    assert_fused_parity(g, o, TOL_60)
    # and one shared non-zero angle (the scalar path of the grid kernel)
    g2, o2 = make_pair(26, p)
    set_colliders(g2, o2, vb.sphere_plane_colliders())
    same = np.full(len(o2.buffer("bendAngles")), 0.25, np.float32)
    o2.buffer("bendAngles")[:] = same
    g2.upload("bendAngles", same)
    for _ in range(6):
        g2.Simulate()
        o2.simulate()
    assert_fused_parity(g2, o2, TOL_60)


def test_two_solvers_running_concurrently_on_one_device():
    """A cloth with several tiles per CTA runs the ten Jacobi iterations of a substep in ONE launch with grid-wide barriers in
    between (solver.cu: recordFusedFrame).  Such a launch is cooperative: two solvers whose frames overlap on the device (each
    has its own stream) must neither dead-lock nor disturb each other's results."""
    p = gpu_params(numSubsteps=3, numIterations=6)

    def make(height):
        g = vb.build_scene(479, p, position=(0, height, 1.0))  # 480 x 480 particles = 1024 tiles: multi-iteration launches
        g.UpdateColliders(vb.sphere_plane_colliders())
        return g

    a, b = make(1.5), make(1.55)
    for _ in range(6):
        a.Simulate(sync=False)
        b.Simulate(sync=False)
    a.Synchronize()
    b.Synchronize()
    # the iterations really were fused into one launch per substep: 19 launches per frame (2 + one rebuild of 7 + 3 x 3 + normals);
    # with one launch per iteration it would be 34
    if a.iterateKernel == vb.ITERATE_GRID:
        assert a.lastLaunchCount < 25, a.lastLaunchCount
    ra, rb = make(1.5), make(1.55)
    for _ in range(6):
        ra.Simulate()
    for _ in range(6):
        rb.Simulate()
    for name in ("positions", "velocities", "normals"):
        assert np.array_equal(a.download(name), ra.download(name)), name
        assert np.array_equal(b.download(name), rb.download(name)), name
