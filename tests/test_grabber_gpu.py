"""MouseGrabber as device operations (velvet_solver_grab / drag / release, input_kernels.cuh) against its CPU restatement
(oracle/grabber.py, MouseGrabber.hpp L31-110): the pick, the pinned mass, the dragged position / velocity, and the frames
simulated while a vertex is being dragged."""
import numpy as np
import pytest

import velvet_b200 as vb
from oracle.grabber import FLT_MAX, MouseGrabber
from util import gpu_params, make_pair, set_colliders

from test_solver_gpu import TOL_60, assert_fused_parity

pytestmark = pytest.mark.gpu


def _oracle_grabber(o, diameter):
    return MouseGrabber(o.buffer("positions"), o.buffer("velocities"), o.buffer("invMasses"), diameter)


def _ray_through(pos, index, origin):
    d = pos[index].astype(np.float64) - np.asarray(origin, np.float64)
    return np.asarray(origin, np.float32), (d / np.linalg.norm(d)).astype(np.float32)


def test_pick_matches_the_host_loop_on_a_draped_cloth():
    g, o = make_pair(40, gpu_params(numSubsteps=3, numIterations=5))
    set_colliders(g, o, vb.sphere_plane_colliders())
    for _ in range(12):
        g.Simulate()
        o.simulate()
    pos = g.download("positions")
    m = _oracle_grabber(o, g.simParams.particleDiameter)
    rng = np.random.default_rng(3)
    hits = 0
    for trial in range(40):
        origin = (rng.uniform(-2, 2), rng.uniform(0.5, 4), rng.uniform(-2, 3))
        if trial % 4 == 3:  # some rays that miss everything
            ro, rd = np.asarray(origin, np.float32), np.asarray((0, 1, 0), np.float32)
        else:
            ro, rd = _ray_through(pos, int(rng.integers(len(pos))), origin)
        want = m.grab(ro, rd)
        got = g.Grab(ro, rd)
        assert got[0] == want[0], (trial, got, want)
        assert np.float32(got[1]).view(np.uint32) == np.float32(want[1]).view(np.uint32), (trial, got, want)
        hits += want[0] >= 0
        if want[0] >= 0:
            assert g.download("invMasses")[want[0]] == 0
        m.release()
        g.Release()
        assert np.array_equal(g.download("invMasses"), o.buffer("invMasses"))
    assert 20 <= hits < 40


def test_occluded_vertices_and_ties_resolve_like_the_strict_less_than_loop():
    """Three particles on the ray (the nearest wins), two at the same distance (the lower index wins), one behind the origin
    (negative distanceToView is still the smallest)."""
    g = vb.VtClothSolverGPU(gpu_params())
    v = np.array([[0, 0, 5], [0, 0, 3], [0.001, 0, 3], [0, 0, 9], [5, 5, 5], [0, 0, -2]], np.float32)
    idx = np.array([0, 1, 2, 3, 4, 5], np.uint32)
    g.AddCloth(v, idx, np.eye(4, dtype=np.float32), 0.1)
    o, d = np.zeros(3, np.float32), np.array([0, 0, 1], np.float32)
    i, dist = g.Grab(o, d)
    assert (i, float(dist)) == (5, -2.0)   # behind the origin: distanceToView -2 is the minimum (L99-104 has no front test)
    g.Release()
    pos = g.download("positions"); pos[5] = (9, 9, 9); g.upload("positions", pos)
    i, dist = g.Grab(o, d)
    assert (i, float(dist)) == (1, 3.0)    # 1 and 2 tie at distance 3: the first index
    g.Release()
    assert g.Grab(np.array([0, 50, 0], np.float32), np.array([1, 0, 0], np.float32)) == (-1, FLT_MAX)
    assert np.array_equal(g.download("invMasses"), np.ones(6, np.float32))


def test_frames_simulated_while_dragging_match_the_oracle():
    p = gpu_params(numSubsteps=3, numIterations=6)
    g, o = make_pair(24, p, attached=(0, 24))
    set_colliders(g, o, vb.sphere_plane_colliders())
    m = _oracle_grabber(o, g.simParams.particleDiameter)
    for _ in range(3):
        g.Simulate()
        o.simulate()
    pos = g.download("positions")
    origin = (0.3, 3.0, 2.5)
    ro, rd = _ray_through(pos, 24 * 25 + 12, origin)   # a vertex of the free edge
    assert g.Grab(ro, rd)[0] == m.grab(ro, rd)[0] >= 0
    for fr in range(8):
        # the mouse moves: the ray swings a little every frame
        rd2 = rd + np.float32(0.01 * (fr + 1)) * np.array([1, 0.5, 0], np.float32)
        g.Drag(ro, rd2)
        m.drag(ro, rd2)
        g.Simulate()
        o.simulate()
    assert_fused_parity(g, o, TOL_60)
    g.Release()
    m.release()
    for _ in range(3):
        g.Simulate()
        o.simulate()
    assert np.array_equal(g.download("invMasses"), o.buffer("invMasses"))
    assert_fused_parity(g, o, TOL_60)
