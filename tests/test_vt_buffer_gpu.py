"""VtBuffer<T> / VtMergedBuffer<T> / VtRegisteredBuffer<T> (reference: VtBuffer.hpp L7-236) exercised directly from C++:
tests/cpp/vt_buffer_test.cu is compiled with nvcc against velvet_b200/csrc/vt_buffer.hpp and run on the GPU.  The renderer
hand-off (registered arrays + sync), the host-readable hash arrays and the NaN guard through the C ABI."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import velvet_b200 as vb

from util import gpu_params

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_vt_buffer_cpp_unit_test(tmp_path):
    exe = str(tmp_path / "vt_buffer_test")
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "-I", os.path.join(ROOT, "velvet_b200", "csrc"),
           "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "vt_buffer_test.cu"), "-o", exe]
    b = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert b.returncode == 0, b.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "all checks passed" in r.stdout


def _dev_array(L, nbytes):
    p = C.c_void_p()
    L.velvet_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    assert L.velvet_alloc(C.byref(p), nbytes) == 0
    return p


def test_render_targets_receive_each_cloths_range():
    """positions.sync() / normals.sync() of the reference (VtClothSolverGPU.hpp L107-110): two cloths, two pairs of
    caller-owned device arrays standing in for the mapped GL vertex buffers."""
    p = gpu_params(numSubsteps=2, numIterations=4)
    g = vb.VtClothSolverGPU(p)
    sizes = []
    for R, pos in ((15, (0, 1.5, 1.0)), (9, (0.2, 1.8, 1.0))):
        v, idx = vb.GenerateClothMesh(R)
        vb.VtClothObjectGPU(R, g).Start(v, idx, vb.TransformMatrix(pos, (90, 0, 0), (1, 1, 1)))
        sizes.append((R + 1) ** 2)
    g.UpdateColliders(vb.sphere_plane_colliders())
    L = g._L
    targets = [(_dev_array(L, 12 * n), _dev_array(L, 12 * n)) for n in sizes]
    for c, (tp, tn) in enumerate(targets):
        g.SetRenderTargets(c, tp.value, tn.value)
    for _ in range(3):
        g.Simulate(sync=False)
        g.SyncRenderTargets()
    g.Synchronize()
    pos, nrm = g.download("positions"), g.download("normals")
    L.velvet_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    off = 0
    for (tp, tn), n in zip(targets, sizes):
        hp, hn = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
        assert L.velvet_copy(hp.ctypes.data_as(C.c_void_p), tp, hp.nbytes) == 0
        assert L.velvet_copy(hn.ctypes.data_as(C.c_void_p), tn, hn.nbytes) == 0
        assert np.array_equal(hp, pos[off:off + n]) and np.array_equal(hn, nrm[off:off + n])
        off += n
    with pytest.raises(vb.VelvetError):
        g.SetRenderTargets(2, targets[0][0].value, targets[0][1].value)  # no such cloth
    g.SetRenderTargets(1, 0, 0)  # detach
    g.Simulate()
    g.SyncRenderTargets()
    g.Synchronize()
    hp = np.zeros((sizes[1], 3), np.float32)
    L.velvet_copy(hp.ctypes.data_as(C.c_void_p), targets[1][0], hp.nbytes)
    assert np.array_equal(hp, pos[sizes[0]:])  # untouched by the last frame
    for tp, tn in targets:
        L.velvet_free(tp)
        L.velvet_free(tn)


def test_hash_arrays_host_readable_like_the_reference():
    """SpatialHashGPU.hpp L54-60: the five hash VtBuffers can be indexed on the host.  Off by default (plain device memory),
    on request they live in managed memory and the frame is unchanged."""
    p = gpu_params(numSubsteps=2, numIterations=4)
    outs = []
    for managed in (False, True):
        g = vb.VtClothSolverGPU(p)
        if managed:
            g.SetHashHostReadable(True)
        v, idx = vb.GenerateClothMesh(31)
        vb.VtClothObjectGPU(31, g).Start(v, idx, vb.TransformMatrix((0, 1.5, 1.0), (90, 0, 0), (1, 1, 1)))
        g.UpdateColliders(vb.sphere_plane_colliders())
        for _ in range(2):
            g.Simulate()
        if managed:
            ptr, n = g.buffer_ptr("particleIndex")
            direct = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint32)), shape=(n,)).copy()  # host dereference
            assert np.array_equal(direct, g.download("particleIndex"))
            ptr, n = g.buffer_ptr("neighbors")
            direct = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint32)), shape=(n,)).copy()
            assert np.array_equal(direct, g.download("neighbors"))
            with pytest.raises(vb.VelvetError):
                g.SetHashHostReadable(False)  # too late: cloth registered
        outs.append((g.download("positions"), g.download("particleIndex"), g.download("cellStart")))
    for a, b in zip(outs[0], outs[1]):
        assert np.array_equal(a, b)


def test_nan_guard_finds_the_first_non_finite_particle():
    p = gpu_params()
    g = vb.build_scene(20, p)
    g.UpdateColliders(vb.sphere_plane_colliders())
    g.Simulate()
    assert g.CheckNaN() == (0, g.simParams.numParticles)
    vel = g.download("velocities").copy()
    vel[137, 1] = np.nan
    vel[300, 0] = np.inf
    g.upload("velocities", vel.reshape(-1))
    cnt, first = g.CheckNaN()
    assert cnt == 2 and first == 137
    g.Simulate()  # the NaN spreads through the constraints: the guard sees it in positions too
    cnt, first = g.CheckNaN()
    assert cnt > 2 and first <= 137
