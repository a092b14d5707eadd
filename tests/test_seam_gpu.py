"""GPU parity of the kernel seam (one C-ABI entry per reference free function) against the O1 oracle.
Bit-exact wherever a kernel has a fixed summation order; atomics-based scatters (stretch / bend / attach /
normals, nondeterministic order in the reference too) are compared at 2e-6 absolute."""
import ctypes as C

import numpy as np
import pytest

import velvet_b200 as vb
from velvet_b200 import seam
from oracle import o1

from util import gpu_params, to_o1_params, to_o1_collider, ColliderTrack

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

f = lambda a: a.ctypes.data_as(C.c_void_p)


def dev(a):
    return torch.from_numpy(np.array(a, copy=True)).cuda()


def host(t):
    torch.cuda.synchronize()
    return t.cpu().numpy()


@pytest.fixture(scope="module")
def state():
    """A crumpled 40x40 cloth state taken from a few oracle frames: realistic inputs for every kernel."""
    R = 39
    p = gpu_params(numSubsteps=3, numIterations=4)
    s = o1.O1Solver(to_o1_params(p))
    v, idx = o1.generate_cloth_mesh(R)
    s.cloth_object_start(R, v, idx, o1.transform_matrix((0, 1.0, 1.0), (70, 10, 0), (1, 1, 1)), [0, R])
    cols = [vb.MakeCollider(vb.COLLIDER_PLANE, (0, 0, 0), (1, 1, 1)),
            vb.MakeCollider(vb.COLLIDER_SPHERE, (0, 0.5, 0), (0.5, 0.5, 0.5))]
    s.set_colliders([to_o1_collider(c) for c in cols])
    for _ in range(12):
        s.simulate()
    P = vb.VtSimParams()
    C.memmove(C.byref(P), C.byref(s.params), 80)
    seam.SetSimulationParams(P)
    return s, P, cols


def _copy(s, name):
    return s.buffer(name).copy()


def test_initialize_positions(state):
    rng = np.random.default_rng(0)
    pos = rng.uniform(-1, 1, (1000, 3)).astype(np.float32)
    M = vb.TransformMatrix((0.3, 1.5, -1), (33, -20, 71), (1.5, 2, 0.7))
    ref = pos.copy()
    o1.lib().o1_initialize_positions(f(ref), 100, 800, f(M))
    d = dev(pos)
    seam.InitializePositions(d, 100, 800, M)
    assert np.array_equal(host(d), ref)


def test_predict_positions(state):
    s, P, _ = state
    pos, vel = _copy(s, "positions"), _copy(s, "velocities")
    pred = np.zeros_like(pos)
    rv, rp = vel.copy(), pred.copy()
    dt = np.float32(1 / 180)
    o1.lib().o1_predict_positions(C.byref(s.params), f(rp), f(rv), f(pos), dt)
    dv, dp = dev(vel), dev(pred)
    seam.PredictPositions(dp, dv, dev(pos), dt)
    assert np.array_equal(host(dv), rv) and np.array_equal(host(dp), rp)


def _predicted(s):
    pred = _copy(s, "positions")
    rng = np.random.default_rng(1)
    return pred + rng.normal(0, 0.004, pred.shape).astype(np.float32)


def test_solve_stretch_attach_bend_apply(state):
    s, P, _ = state
    n = s.params.numParticles
    pred = _predicted(s)
    inv = _copy(s, "invMasses")
    si, sl = _copy(s, "stretchIndices"), _copy(s, "stretchLengths")
    bi, ba = _copy(s, "bendIndices"), _copy(s, "bendAngles")
    ap, asl, ad, asp = (_copy(s, k) for k in ("attachParticleIDs", "attachSlotIDs", "attachDistances", "attachSlotPositions"))
    dt = np.float32(1 / 180)

    rd, rc, rp = np.zeros(3 * n, np.float32), np.zeros(n, np.int32), pred.copy()
    L = o1.lib()
    L.o1_solve_stretch(f(rp), f(rd), f(rc), f(si), f(sl), f(inv), len(sl))
    L.o1_solve_attachment(C.byref(s.params), f(rp), f(rd), f(rc), f(inv), f(ap), f(asl), f(asp), f(ad), len(ad))
    L.o1_solve_bending(C.byref(s.params), f(rp), f(rd), f(rc), f(bi), f(ba), f(inv), len(ba), dt)
    rd_before = rd.copy()
    rc_before = rc.copy()
    L.o1_apply_deltas(C.byref(s.params), f(rp), f(rd), f(rc))

    dd, dc, dp = dev(np.zeros(3 * n, np.float32)), dev(np.zeros(n, np.int32)), dev(pred)
    dinv = dev(inv)
    seam.SolveStretch(dp, dd, dc, dev(si), dev(sl), dinv, len(sl))
    seam.SolveAttachment(dp, dd, dc, dinv, dev(ap), dev(asl), dev(asp), dev(ad), len(ad))
    seam.SolveBending(dp, dd, dc, dev(bi), dev(ba), dinv, len(ba), dt)
    assert np.array_equal(host(dc), rc_before), "delta counts are integer work: bit-exact"
    assert np.max(np.abs(host(dd) - rd_before)) < 2e-6
    seam.ApplyDeltas(dp, dd, dc)
    assert np.max(np.abs(host(dp) - rp)) < 2e-6
    assert not host(dd).any() and not host(dc).any(), "ApplyDeltas clears deltas/counts where count > 0"


def test_apply_deltas_bit_exact(state):
    s, P, _ = state
    n = s.params.numParticles
    rng = np.random.default_rng(5)
    pred = _predicted(s)
    deltas = rng.normal(0, 0.01, 3 * n).astype(np.float32)
    counts = rng.integers(0, 9, n).astype(np.int32)
    deltas.reshape(-1, 3)[counts == 0] = 0
    rp, rd, rc = pred.copy(), deltas.copy(), counts.copy()
    o1.lib().o1_apply_deltas(C.byref(s.params), f(rp), f(rd), f(rc))
    dp, dd, dc = dev(pred), dev(deltas), dev(counts)
    seam.ApplyDeltas(dp, dd, dc)
    assert np.array_equal(host(dp), rp) and np.array_equal(host(dd), rd) and np.array_equal(host(dc), rc)


@pytest.mark.parametrize("alias", [False, True])
def test_collide_sdf_plane_sphere_cube(state, alias):
    s, P, _ = state
    n = s.params.numParticles
    rng = np.random.default_rng(11)
    pos = rng.uniform(-1.2, 1.2, (n, 3)).astype(np.float32)
    pos[:, 1] = rng.uniform(-0.1, 1.5, n).astype(np.float32)
    pred = pos if alias else (pos + rng.normal(0, 0.02, pos.shape).astype(np.float32))
    sphere = ColliderTrack(vb.COLLIDER_SPHERE, (0.1, 0.5, 0.0), (0.5, 0.5, 0.5))
    sphere.move((0.12, 0.5, 0.03))
    cube = ColliderTrack(vb.COLLIDER_CUBE, (-0.4, 0.5, 0.3), (1.0, 1.0, 0.8), (10, 25, -5))
    cube.move((-0.38, 0.52, 0.3), (12, 27, -5))
    cols = [vb.MakeCollider(vb.COLLIDER_PLANE, (0, 0, 0), (1, 1, 1)), sphere.collider(), cube.collider()]
    ocols = o1.colliders_array([to_o1_collider(c) for c in cols])
    dt = np.float32(1 / 60 if alias else 1 / 180)
    rp = pred.copy().reshape(-1)
    rpos = rp if alias else pos.reshape(-1)
    o1.lib().o1_collide_sdf(C.byref(s.params), f(rp), C.cast(ocols, C.c_void_p), f(rpos), 3, dt)
    raw = np.frombuffer(b"".join(bytes(c) for c in cols), np.uint8)
    dcols = dev(raw)
    dp = dev(pred.reshape(-1))
    dpos = dp if alias else dev(pos.reshape(-1))
    seam.CollideSDF(dp, dcols, dpos, 3, dt)
    out = host(dp)
    moved = np.any(out.reshape(-1, 3) != pred.reshape(-1, 3), axis=1).sum()
    assert moved > n // 10, "test must exercise the colliders"
    assert np.array_equal(out, rp)


def test_hash_and_collide_particles(state):
    s, P, _ = state
    n = s.params.numParticles
    pred = _predicted(s)
    pos, inv, init = _copy(s, "positions"), _copy(s, "invMasses"), _copy(s, "initialPositions")
    D = s.params.particleDiameter
    cell = np.float32(D * np.float32(1.5))
    hp = vb.VtHashParams(n, 64, cell, np.float32(cell * cell), 2 * n, np.float32(D * D))
    ohp = o1.HashParams(n, 64, cell, np.float32(cell * cell), 2 * n, np.float32(D * D))
    ph, pi, cs, ce, nb = (np.zeros(k, np.uint32) for k in (n, n, 2 * n, 2 * n, 64 * n))
    o1.lib().o1_hash_objects(f(ph), f(pi), f(cs), f(ce), f(nb), f(pred), f(init), ohp)
    dph, dpi, dcs, dce, dnb = (dev(np.zeros(k, np.uint32)) for k in (n, n, 2 * n, 2 * n, 64 * n))
    dpred = dev(pred)
    seam.HashObjects(dph, dpi, dcs, dce, dnb, dpred, dev(init), hp)
    assert np.array_equal(host(dph), ph), "cell keys / sorted order"
    assert np.array_equal(host(dpi), pi), "sorted particle order"
    assert np.array_equal(host(dcs), cs), "cellStart"
    valid = cs != 0xFFFFFFFF
    assert np.array_equal(host(dce)[valid], ce[valid]), "cellEnd (only defined where cellStart is)"
    from util import valid_prefix_table
    assert np.array_equal(valid_prefix_table(host(dnb), n, 64), valid_prefix_table(nb, n, 64)), "neighbor lists"
    assert (valid_prefix_table(nb, n, 64) != 0xFFFFFFFF).sum() > 4 * n

    rd, rc, rp = np.zeros(3 * n, np.float32), np.zeros(n, np.int32), pred.copy()
    o1.lib().o1_collide_particles(C.byref(s.params), f(rd), f(rc), f(rp), f(inv), f(nb), f(pos))
    dd, dc = dev(np.zeros(3 * n, np.float32)), dev(np.zeros(n, np.int32))
    seam.CollideParticles(dd, dc, dpred, dev(inv), dnb, dev(pos))
    assert np.any(rp != pred), "some particles must be in contact"
    assert np.array_equal(host(dpred), rp)


def test_finalize_with_speed_clamp(state):
    s, P, _ = state
    pos = _copy(s, "positions")
    rng = np.random.default_rng(2)
    pred = pos + rng.normal(0, 0.1, pos.shape).astype(np.float32)  # some beyond maxSpeed * dt
    dt = np.float32(1 / 180)
    rv, rpos = np.zeros_like(pos), pos.copy()
    o1.lib().o1_finalize(C.byref(s.params), f(rv), f(rpos), f(pred), dt)
    dv, dpos = dev(np.zeros_like(pos)), dev(pos)
    seam.Finalize(dv, dpos, dev(pred), dt)
    speed = np.linalg.norm(rv.reshape(-1, 3), axis=1)
    assert (speed > 0.99 * s.params.maxSpeed * (1 - 0.25 * dt)).any(), "clamp branch exercised"
    assert np.array_equal(host(dv), rv) and np.array_equal(host(dpos), rpos)


def test_compute_normal(state):
    s, P, _ = state
    pos, idx = _copy(s, "positions"), _copy(s, "indices")
    n = s.params.numParticles
    rn = np.zeros(3 * n, np.float32)
    o1.lib().o1_compute_normal(C.byref(s.params), f(rn), f(pos), f(idx), len(idx) // 3)
    dn = dev(np.full(3 * n, 7.0, np.float32))
    seam.ComputeNormal(dn, dev(pos), dev(idx), len(idx) // 3)
    out = host(dn)
    assert np.max(np.abs(out - rn)) < 2e-6
    assert np.allclose(np.linalg.norm(out.reshape(-1, 3), axis=1), 1, atol=1e-5)


def test_zero_sized_calls_are_noops(state):
    # CUDA_CALL returns silently on 0 threads (Common.cuh L24-25)
    seam.SolveStretch(0, 0, 0, 0, 0, 0, 0)
    seam.SolveBending(0, 0, 0, 0, 0, 0, 0, 0.01)
    seam.SolveAttachment(0, 0, 0, 0, 0, 0, 0, 0, 0)
    seam.CollideSDF(0, 0, 0, 0, 0.01)
    seam.HashObjects(0, 0, 0, 0, 0, 0, 0, vb.VtHashParams(0, 64, 0.1, 0.01, 0, 0.01))
    seam.SortPairs(0, 0, 0, 11)
    seam.synchronize()


def test_library_division_is_ieee_for_every_operand_class():
    """vt_div / vec3-by-scalar (vt_math.cuh) replace the compiler's division in every exact kernel: they must return the IEEE-754
    round-to-nearest quotient for all operands -- random bit patterns, zeros of both signs, denormals, huge / tiny exponents,
    inf and NaN -- bit for bit (NaNs compared as a class), against the compiler's x / y and against numpy on the CPU."""
    from velvet_b200 import _capi
    L = _capi.load()
    L.velvet_selftest_division.argtypes = [C.c_void_p] * 2 + [C.c_uint] + [C.c_void_p] * 3
    rng = np.random.default_rng(7)
    n = 1 << 22
    special = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 1e-45, -1e-45, 1e-39, 3e38, -3e38, 2.0 ** -60, 2.0 ** 60,
                        np.nextafter(np.float32(2.0 ** -60), np.float32(0)), np.nextafter(np.float32(2.0 ** 60), np.float32(np.inf)),
                        1e-20, 1e20, 0.1, 3.0], np.float32)

    def operands():
        kind = rng.integers(0, 4, n)
        bits = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32).view(np.float32)       # any bit pattern
        moderate = (rng.standard_normal(n) * np.exp(rng.uniform(-20, 20, n))).astype(np.float32)     # cloth-like magnitudes
        wide = (rng.choice([-1.0, 1.0], n) * np.exp2(rng.uniform(-140, 127, n))).astype(np.float32)  # whole exponent range
        spec = special[rng.integers(0, len(special), n)]
        return np.where(kind == 0, bits, np.where(kind == 1, moderate, np.where(kind == 2, wide, spec))).astype(np.float32)

    x, y = operands(), operands()
    # every pair of specials at least once
    sx, sy = np.meshgrid(special, special)
    x[: sx.size], y[: sy.size] = sx.ravel(), sy.ravel()
    dx, dy = dev(x), dev(y)
    od = torch.zeros(n, dtype=torch.int32, device="cuda")
    ov = torch.zeros(3 * n, dtype=torch.int32, device="cuda")
    op = torch.zeros(n, dtype=torch.int32, device="cuda")
    _capi.check(L.velvet_selftest_division(dx.data_ptr(), dy.data_ptr(), n, od.data_ptr(), ov.data_ptr(), op.data_ptr()))
    got = host(od).view(np.uint32)
    vec = host(ov).view(np.uint32).reshape(n, 3)
    plain = host(op).view(np.uint32)
    with np.errstate(all="ignore"):
        want = [(np.roll(x, -k) / y) for k in range(3)]

    def same(a_bits, ref):
        a = a_bits.view(np.float32)
        return (a_bits == ref.view(np.uint32)) | (np.isnan(a) & np.isnan(ref))

    assert same(plain, want[0]).all(), "the compiler's own division disagrees with the CPU: not an IEEE environment"
    bad = ~same(got, want[0])
    assert not bad.any(), (x[bad][:5], y[bad][:5], got[bad][:5], want[0].view(np.uint32)[bad][:5])
    for k in range(3):
        bad = ~same(np.ascontiguousarray(vec[:, k]), want[k])
        assert not bad.any(), (k, np.roll(x, -k)[bad][:5], y[bad][:5])


def test_checked_fast_constraint_evaluators_match_the_branchy_ones():
    """The Jacobi kernel evaluates constraints with branch-light IEEE sequences guarded by ONE validity predicate
    (stretch_eval_u / bend_eval_u / vt_sqrt_u) and falls back to the plain evaluators when it fails.  Wherever the predicate
    holds the results must be bit-identical -- cloth-like geometry (where it must hold almost always, or the fast path
    would be pointless), degenerate geometry (repeated points, zero weights, rest length hit exactly) and arbitrary bit
    patterns -- and the square root is swept over all 2^32 operands."""
    from velvet_b200 import _capi
    L = _capi.load()
    L.velvet_selftest_constraints.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(11)
    n = 1 << 20
    h = 2.0 / 1023

    def cloth_like(m):
        base = rng.uniform(-1, 1, (m, 1, 3))
        quad = np.array([[0, 0, 0], [h, h, 0], [h, 0, 0], [0, h, 0]])  # wing0, wing1, edge2, edge3
        fold = rng.uniform(-0.6, 0.6, (m, 1)) * h
        pts = base + quad[None] + rng.normal(0, 0.03 * h, (m, 4, 3))
        pts[:, 1, 2] += fold[:, 0]
        w = rng.choice([0.0, 1.0, 1.0, 1.0, 2.5], (m, 4))
        rest = np.linalg.norm(pts[:, 0] - pts[:, 1], axis=1) * rng.uniform(0.9, 1.1, m)
        return np.concatenate([pts.reshape(m, 12), w, rest[:, None], rng.choice([0.0, 1e-3], (m, 1))], axis=1)

    a = cloth_like(n).astype(np.float32)
    b = cloth_like(n).astype(np.float32)
    # degenerate: exactly flat quads on a lattice (zero cross-product components, acos(1)), repeated points, all pinned
    k = np.arange(n)
    b[:, 2::3][:, :4] = 0.0
    b[k % 7 == 0, 3:6] = b[k % 7 == 0, 0:3]
    b[k % 11 == 0, 12:16] = 0.0
    b[k % 13 == 0, 9:12] = b[k % 13 == 0, 6:9]
    exact = (np.linalg.norm(b[:, 0:3].astype(np.float32) - b[:, 3:6].astype(np.float32), axis=1)).astype(np.float32)
    b[k % 5 == 0, 16] = exact[k % 5 == 0]
    c = rng.integers(0, 2 ** 32, (n, 18), dtype=np.uint64).astype(np.uint32).view(np.float32)        # any bit pattern
    d = (rng.choice([-1.0, 1.0], (n, 18)) * np.exp2(rng.uniform(-140, 127, (n, 18)))).astype(np.float32)  # whole exponent range
    counts = {}
    for name, ops in (("cloth", a), ("degenerate", b), ("bits", c), ("wide", d)):
        x = dev(np.ascontiguousarray(ops, np.float32))
        mism = torch.zeros(3, dtype=torch.int64, device="cuda")
        fast = torch.zeros(3, dtype=torch.int64, device="cuda")
        _capi.check(L.velvet_selftest_constraints(x.data_ptr(), n, mism.data_ptr(), fast.data_ptr()))
        m, fs = host(mism), host(fast)
        counts[name] = (m.tolist(), fs.tolist())
        assert m[0] == 0 and m[1] == 0 and m[2] == 0, (name, m, fs)
    print("selftest_constraints (mismatches, fast-path valid):", counts)
    # cloth-like operands: pinned-pinned pairs (1/25 of the stretch operands, 1/625 of the bends) are the only expected fallbacks
    assert counts["cloth"][1][0] >= 0.94 * n and counts["cloth"][1][1] >= 0.98 * n
    assert counts["cloth"][1][2] > 2 ** 30  # sqrt fast path covers [2^-101, 2^128)
