#!/usr/bin/env python
"""Generates the golden fixtures of tests/golden/ by running the REFERENCE's own CUDA kernels (O3: oracle/_ref/
libvelvet_refcuda.so, built from the unmodified VtClothSolverGPU.cu + SpatialHashGPU.cu of /root/reference by
oracle/ref_cuda/build_ref_cuda.sh) on a B200.  Needs a GPU, so it is run on the GPU box:

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'      # then copy gpurun_out/golden/*.npz here

Nothing of this repo's product is involved: inputs come from the O1 registration code only to get the same
constraint lists the reference's VtClothObjectGPU::Start would generate (checked separately against the reference
kernels in tests/test_ref_cuda_gpu.py), all outputs come from the reference kernels.

Fixtures (all small, np.savez_compressed):
  refcuda_hash_R{31,63}.npz   spatial hash on a seeded perturbed sheet: inputs (predicted, initialPositions) and the
                              reference's particleHash, particleIndex, cellStart, cellEnd, neighbour table (entries
                              after each column's terminator masked: they are stale and never read)
  refcuda_hash_R1023_digest.npz  the same at the headline size (1,048,576 particles): SHA-256 of every output buffer and the
                              full neighbour lists of 4,096 sampled particles; inputs are regenerated from seeds
  refcuda_cfg1.npz            BASELINE configs[0] (32x32, 2 attach points, plane + moving sphere, 5 substeps x 10
                              iterations): positions after frames 1, 5, 10, 15 (contact-free, where the reference is
                              self-consistent), velocities + normals after frame 1
  refcuda_drape64.npz         64x64 self-colliding drape over sphere + plane: positions after frames 1, 10, 20
  refcuda_cube25.npz          25x25 cloth with 4 corner attachments (long-range attachment active) falling on a rotated,
                              slowly turning unit cube + plane, friction 0.6, 5 substeps x 5 iterations: positions after
                              frames 1, 5, 10 (cube SDF with rounded edges, collider velocity in the friction term)
"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import o1, refcuda  # noqa: E402


def params(**kw):
    p = o1.default_params()
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def pair(R, p, position, rotation, attached):
    o = o1.O1Solver(p)
    v, idx = o1.generate_cloth_mesh(R)
    model = o1.transform_matrix(position, rotation, (1, 1, 1))
    o.cloth_object_start(R, v, idx, model, attached)
    r = refcuda.RefCudaSolver(p)
    r.register_like(o, R, model, attached)
    return o, r


def masked_table(nb, n, k=64):
    tab = nb[: n * k].reshape(k, n).copy()
    after = np.cumsum(tab == 0xFFFFFFFF, axis=0) > 0
    tab[after] = 0xFFFFFFFF
    return tab


def hash_fixture(R, out):
    p = params()
    o, r = pair(R, p, (0, 1.5, 1.0), (90, 0, 0), [])
    n = (R + 1) ** 2
    rng = np.random.default_rng(1000 + R)
    init = r.buffer("initialPositions").copy()
    pred = (r.buffer("positions") + rng.normal(0, 0.02, 3 * n)).astype(np.float32)
    r.buffer("predicted")[:] = pred
    r.hash_predicted()
    cs = r.buffer("cellStart").copy()
    ce = r.buffer("cellEnd").copy()
    ce[cs == 0xFFFFFFFF] = 0  # never written for empty buckets: undefined in the reference
    np.savez_compressed(os.path.join(out, f"refcuda_hash_R{R}.npz"), resolution=R, predicted=pred, initialPositions=init,
                        particleDiameter=np.float32(r.params.particleDiameter),
                        particleHash=r.buffer("particleHash").copy(), particleIndex=r.buffer("particleIndex").copy(),
                        cellStart=cs, cellEnd=ce, neighbors=masked_table(r.buffer("neighbors"), n))


def hash_inputs_1m(o):
    """Inputs of the headline-size hash fixture, regenerated from seeds by every consumer: the O1 registration of the
    1024x1024 drape (initialPositions) and its positions plus seeded noise of a fifth of a cell (predicted)."""
    n = 1 << 20
    rng = np.random.default_rng(2024)
    init = o.buffer("initialPositions").copy()
    pred = (o.buffer("positions") + rng.normal(0, 0.0009, 3 * n)).astype(np.float32)
    return init, pred


def digest(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def hash_digest_fixture(out):
    """BASELINE configs[2] size: the reference kernels' hash of 1,048,576 particles.  The full outputs are 280 MB, so the
    fixture keeps SHA-256 digests of every buffer plus the complete neighbour lists of 4,096 sampled particles."""
    R = 1023
    p = params()
    o, r = pair(R, p, (0, 1.5, 1.0), (90, 0, 0), [])
    n = (R + 1) ** 2
    init, pred = hash_inputs_1m(o)
    r.buffer("initialPositions")[:] = init
    r.buffer("predicted")[:] = pred
    r.hash_predicted()
    cs = r.buffer("cellStart").copy()
    ce = r.buffer("cellEnd").copy()
    ce[cs == 0xFFFFFFFF] = 0
    tab = masked_table(r.buffer("neighbors"), n)
    sample = np.sort(np.random.default_rng(7).choice(n, 4096, replace=False)).astype(np.uint32)
    np.savez_compressed(os.path.join(out, "refcuda_hash_R1023_digest.npz"), resolution=R,
                        particleDiameter=np.float32(r.params.particleDiameter), inputs_digest=digest(np.concatenate([init, pred])),
                        particleHash=digest(r.buffer("particleHash")), particleIndex=digest(r.buffer("particleIndex")),
                        cellStart=digest(cs), cellEnd=digest(ce), neighbors=digest(tab), neighbor_count=int((tab != 0xFFFFFFFF).sum()),
                        sample=sample, sample_neighbors=tab[:, sample], sample_particleIndex=r.buffer("particleIndex")[::256].copy(),
                        sample_particleHash=r.buffer("particleHash")[::256].copy())


def cfg1_fixture(out):
    p = params(numSubsteps=5, numIterations=10)
    o, r = pair(31, p, (0, 2.5, 0), (0, 0, 0), [0, 31])
    last = o1.transform_matrix((0, 0.6, -1.0), (0, 0, 0), (0.6,) * 3)
    keep = {}
    for f in range(15):
        z = -math.cos(2 * f / 60.0)
        cur = o1.transform_matrix((0, 0.6, z), (0, 0, 0), (0.6,) * 3)
        r.set_colliders([o1.make_collider(o1.PLANE, (0, 0, 0), (1, 1, 1)), o1.make_collider(o1.SPHERE, (0, 0.6, z), (0.6,) * 3, cur, last)])
        last = cur
        r.simulate()
        if f + 1 in (1, 5, 10, 15):
            keep[f"positions_{f + 1}"] = r.buffer("positions").copy()
        if f == 0:
            keep["velocities_1"] = r.buffer("velocities").copy()
            keep["normals_1"] = r.buffer("normals").copy()
            keep["invMasses"] = r.buffer("invMasses").copy()
    np.savez_compressed(os.path.join(out, "refcuda_cfg1.npz"), frames=np.array([1, 5, 10, 15]), **keep)


def drape_fixture(out):
    p = params(numSubsteps=5, numIterations=10)
    o, r = pair(63, p, (0, 1.5, 1.0), (90, 0, 0), [])
    r.set_colliders([o1.make_collider(o1.PLANE, (0, 0, 0), (1, 1, 1)), o1.make_collider(o1.SPHERE, (0, 0.6, 0), (0.6,) * 3)])
    keep = {}
    for f in range(20):
        r.simulate()
        if f + 1 in (1, 10, 20):
            keep[f"positions_{f + 1}"] = r.buffer("positions").copy()
    np.savez_compressed(os.path.join(out, "refcuda_drape64.npz"), frames=np.array([1, 10, 20]), **keep)


def cube_fixture(out):
    R = 24
    p = params(numSubsteps=5, numIterations=5, friction=0.6)
    corners = [0, R, (R + 1) * (R + 1) - 1, (R + 1) * R]
    o, r = pair(R, p, (0, 1.5, 1.0), (90, 0, 0), corners)
    keep = {}
    last = o1.transform_matrix((0, 0.95, 0), (0, 15, 0), (1, 1, 1))
    for f in range(10):
        cur = o1.transform_matrix((0, 0.95, 0), (0, 15 + 2 * (f + 1), 0), (1, 1, 1))
        r.set_colliders([o1.make_collider(o1.PLANE, (0, 0, 0), (1, 1, 1)), o1.make_collider(o1.CUBE, (0, 0.95, 0), (1, 1, 1), cur, last)])
        last = cur
        r.simulate()
        if f + 1 in (1, 5, 10):
            keep[f"positions_{f + 1}"] = r.buffer("positions").copy()
    np.savez_compressed(os.path.join(out, "refcuda_cube25.npz"), frames=np.array([1, 5, 10]), **keep)


if __name__ == "__main__":
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out, exist_ok=True)
    o1.build()
    assert refcuda.available(), "oracle/_ref/libvelvet_refcuda.so missing: run oracle/ref_cuda/build_ref_cuda.sh where /root/reference exists"
    if "--only-hash-1m" in sys.argv:
        hash_digest_fixture(out)
        print("written", sorted(os.listdir(out)))
        sys.exit(0)
    for R in (31, 63):
        hash_fixture(R, out)
    hash_digest_fixture(out)
    cfg1_fixture(out)
    drape_fixture(out)
    cube_fixture(out)
    print("golden fixtures written to", out, sorted(os.listdir(out)))
