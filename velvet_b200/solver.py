"""Python host mirror of the reference's solver interface, over the C ABI (include/velvet_b200.h).

Class and method names follow the reference so that tests read like the reference's own call sites:
  VtClothSolverGPU   (VtClothSolverGPU.hpp L23-231)  AddCloth / AddStretch / AddAttachSlot / AddAttach / AddBend /
                                                     UpdateColliders / Simulate + the public sim buffers
  SpatialHashGPU     (SpatialHashGPU.hpp L15-60)     SetInitialPositions / Hash + neighbors, particleHash, ...
  VtClothObjectGPU   (VtClothObjectGPU.hpp L12-149)  SetAttachedIndices / Start (constraint generation)
numpy arrays are host data; device buffers are exposed as (pointer, count) and copied explicitly.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _capi
from ._capi import (BUFFER_IDS, COLLIDER_CUBE, COLLIDER_PLANE, COLLIDER_SPHERE, ITERATE_AUTO, ITERATE_GRID, ITERATE_TILES, MATH_EXACT, MATH_FAST, PIPELINE_FUSED, PIPELINE_SEAM,
                    VtHashParams, VtSDFCollider, VtSimParams, check)

_BUF_DTYPE = {
    "positions": (np.float32, 3), "normals": (np.float32, 3), "indices": (np.uint32, 1),
    "velocities": (np.float32, 3), "predicted": (np.float32, 3), "deltas": (np.float32, 3),
    "deltaCounts": (np.int32, 1), "invMasses": (np.float32, 1), "stretchIndices": (np.int32, 1),
    "stretchLengths": (np.float32, 1), "bendIndices": (np.uint32, 1), "bendAngles": (np.float32, 1),
    "attachParticleIDs": (np.int32, 1), "attachSlotIDs": (np.int32, 1), "attachDistances": (np.float32, 1),
    "attachSlotPositions": (np.float32, 3), "neighbors": (np.uint32, 1), "initialPositions": (np.float32, 3),
    "particleHash": (np.uint32, 1), "particleIndex": (np.uint32, 1), "cellStart": (np.uint32, 1),
    "cellEnd": (np.uint32, 1),
}


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def default_params() -> VtSimParams:
    p = VtSimParams()
    check(_capi.load().velvet_default_params(C.byref(p)))
    return p


def GenerateClothMesh(resolution: int):
    """Scene.hpp L131-168 -> (vertices float32[(R+1)^2, 3], indices uint32[6 R^2])."""
    v = np.zeros(((resolution + 1) ** 2, 3), np.float32)
    idx = np.zeros(6 * resolution * resolution, np.uint32)
    check(_capi.load().velvet_generate_cloth_mesh(resolution, _ptr(v), _ptr(idx)))
    return v, idx


def TransformMatrix(position=(0, 0, 0), rotation=(0, 0, 0), scale=(1, 1, 1)) -> np.ndarray:
    """Transform::matrix() (Transform.hpp L22-29): column-major float32[16], rotation in degrees."""
    out = np.zeros(16, np.float32)
    check(_capi.load().velvet_transform_matrix(_ptr(np.asarray(position, np.float32)), _ptr(np.asarray(rotation, np.float32)),
                                               _ptr(np.asarray(scale, np.float32)), _ptr(out)))
    return out


def MakeCollider(ctype: int, position, scale, curTransform=None, lastTransform=None, deltaTime: float = 1.0 / 60.0) -> VtSDFCollider:
    """One iteration of the UpdateColliders loop (VtClothSolverGPU.hpp L195-203)."""
    position = np.asarray(position, np.float32)
    scale = np.asarray(scale, np.float32)
    if curTransform is None:
        curTransform = TransformMatrix(position, (0, 0, 0), scale)
    if lastTransform is None:
        lastTransform = curTransform
    out = VtSDFCollider()
    check(_capi.load().velvet_make_collider(ctype, _ptr(position), _ptr(scale), _ptr(np.ascontiguousarray(curTransform, np.float32)),
                                            _ptr(np.ascontiguousarray(lastTransform, np.float32)), float(np.float32(deltaTime)),
                                            C.byref(out)))
    return out


def _collider_array(colliders):
    arr = (VtSDFCollider * max(len(colliders), 1))()
    for i, c in enumerate(colliders):
        arr[i] = c
    return arr


class VtClothSolverGPU:
    def __init__(self, params: VtSimParams | None = None, device: int = -1, pipeline: int = PIPELINE_FUSED,
                 tile_size: int = 0, math_mode: int = MATH_EXACT):
        self._L = _capi.load()
        h = C.c_void_p()
        check(self._L.velvet_solver_create(C.byref(h), device, C.byref(params) if params is not None else None))
        self._h = h
        if pipeline != PIPELINE_FUSED:
            self.SetPipeline(pipeline)
        if tile_size:
            check(self._L.velvet_solver_set_tile_size(self._h, tile_size))
        if math_mode != MATH_EXACT:
            self.SetMathMode(math_mode)

    def SetMathMode(self, mode: int):
        check(self._L.velvet_solver_set_math_mode(self._h, mode))

    def SetIterateMode(self, mode: int):
        """ITERATE_AUTO (default: implicit-grid Jacobi kernel for grid cloths) or ITERATE_TILES (record-driven tile kernel)."""
        check(self._L.velvet_solver_set_iterate_mode(self._h, mode))

    @property
    def iterateKernel(self) -> int:
        """ITERATE_TILES or ITERATE_GRID: the Jacobi kernel the next frame runs."""
        k = C.c_int()
        check(self._L.velvet_solver_iterate_kernel(self._h, C.byref(k)))
        return k.value

    def close(self):
        if getattr(self, "_h", None):
            self._L.velvet_solver_destroy(self._h)
            self._h = None

    __del__ = close

    @property
    def simParams(self) -> VtSimParams:
        return self._L.velvet_solver_params(self._h).contents

    def SetPipeline(self, pipeline: int):
        check(self._L.velvet_solver_set_pipeline(self._h, pipeline))

    def SetTileSize(self, n: int):
        check(self._L.velvet_solver_set_tile_size(self._h, n))

    # ---- registration (VtClothSolverGPU.hpp L114-205)
    def AddCloth(self, vertices, indices, modelMatrix, particleDiameter: float) -> int:
        vertices = np.ascontiguousarray(vertices, np.float32).reshape(-1, 3)
        indices = np.ascontiguousarray(indices, np.uint32).reshape(-1)
        off = C.c_int(0)
        check(self._L.velvet_solver_add_cloth(self._h, _ptr(vertices), len(vertices), _ptr(indices), len(indices),
                                              _ptr(np.ascontiguousarray(modelMatrix, np.float32)),
                                              float(np.float32(particleDiameter)), C.byref(off)))
        return off.value

    def AddStretch(self, idx1: int, idx2: int, distance: float):
        check(self._L.velvet_solver_add_stretch(self._h, idx1, idx2, float(np.float32(distance))))

    def AddAttachSlot(self, pos):
        check(self._L.velvet_solver_add_attach_slot(self._h, _ptr(np.asarray(pos, np.float32))))

    def AddAttach(self, particleIndex: int, slotIndex: int, distance: float):
        check(self._L.velvet_solver_add_attach(self._h, particleIndex, slotIndex, float(np.float32(distance))))

    def AddBend(self, idx1: int, idx2: int, idx3: int, idx4: int, angle: float = 0.0):
        check(self._L.velvet_solver_add_bend(self._h, idx1, idx2, idx3, idx4, float(np.float32(angle))))

    def AddClothInstances(self, resolution: int, vertices, indices, modelMatrices, attached=()):
        """Batched independent cloths: one grid cloth topology, len(modelMatrices) instances (see velvet_b200.h)."""
        vertices = np.ascontiguousarray(vertices, np.float32)
        indices = np.ascontiguousarray(indices, np.uint32)
        models = np.ascontiguousarray(modelMatrices, np.float32).reshape(-1, 16)
        att = np.asarray(list(attached), np.int32)
        check(self._L.velvet_solver_add_cloth_instances(self._h, resolution, _ptr(vertices), _ptr(indices), _ptr(models),
                                                        len(models), _ptr(att), len(att)))

    def UpdateColliders(self, colliders):
        arr = _collider_array(colliders)
        check(self._L.velvet_solver_update_colliders(self._h, C.cast(arr, C.c_void_p), len(colliders)))

    def UpdateCollidersRaw(self, ptr, n: int):
        check(self._L.velvet_solver_update_colliders(self._h, ptr, n))

    # ---- per frame
    def Simulate(self, dt: float | None = None, sync: bool = True):
        if dt is None:
            check(self._L.velvet_solver_simulate(self._h, 1 if sync else 0))
        else:
            check(self._L.velvet_solver_simulate_dt(self._h, float(np.float32(dt)), 1 if sync else 0))

    def SimulateTimed(self) -> dict:
        labels = (C.c_char_p * 32)()
        ms = (C.c_float * 32)()
        n = check(self._L.velvet_solver_simulate_timed(self._h, labels, ms, 32))
        return {labels[i].decode(): ms[i] for i in range(n)}

    def Synchronize(self):
        check(self._L.velvet_solver_synchronize(self._h))

    def Hash(self):
        """m_spatialHash->Hash(predicted) (VtClothSolverGPU.hpp L83)."""
        check(self._L.velvet_solver_hash(self._h))
        self.Synchronize()

    def HashFused(self):
        """The fused pipeline's own hash kernels on the public `predicted` buffer (what Simulate runs internally)."""
        check(self._L.velvet_solver_hash_fused(self._h))

    def ReadbackAsync(self, host_positions_ptr, host_normals_ptr):
        check(self._L.velvet_solver_readback_async(self._h, host_positions_ptr, host_normals_ptr))

    def ReadbackPipelined(self, host_positions_ptr, host_normals_ptr) -> int:
        """Double-buffered read-back on a separate copy stream (overlaps the next Simulate); returns a ticket."""
        t = C.c_int(-1)
        check(self._L.velvet_solver_readback_pipelined(self._h, host_positions_ptr, host_normals_ptr, C.byref(t)))
        return t.value

    def ReadbackWait(self, ticket: int):
        check(self._L.velvet_solver_readback_wait(self._h, ticket))

    @property
    def stream(self) -> int:
        return self._L.velvet_solver_stream(self._h) or 0

    @property
    def lastLaunchCount(self) -> int:
        return self._L.velvet_solver_last_launch_count(self._h)

    # ---- buffers (VtClothSolverGPU.hpp L209-231, SpatialHashGPU.hpp L54-60)
    def SetHashHostReadable(self, on: bool = True):
        """Before AddCloth: the five hash arrays in managed memory, host-indexable like the reference's VtBuffers."""
        self._L.velvet_solver_set_hash_host_readable.argtypes = [C.c_void_p, C.c_int]
        check(self._L.velvet_solver_set_hash_host_readable(self._h, 1 if on else 0))

    def SetRenderTargets(self, cloth_index: int, positions_dev: int, normals_dev: int):
        """Device arrays (raw pointers) that SyncRenderTargets mirrors cloth `cloth_index`'s positions / normals into."""
        self._L.velvet_solver_set_render_targets.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        check(self._L.velvet_solver_set_render_targets(self._h, cloth_index, C.c_void_p(positions_dev), C.c_void_p(normals_dev)))

    def SyncRenderTargets(self):
        self._L.velvet_solver_sync_render_targets.argtypes = [C.c_void_p]
        check(self._L.velvet_solver_sync_render_targets(self._h))

    def CheckNaN(self):
        """(number of non-finite components in positions / velocities / predicted, first offending particle)."""
        self._L.velvet_solver_check_nan.argtypes = [C.c_void_p, C.POINTER(C.c_uint), C.POINTER(C.c_uint)]
        cnt, first = C.c_uint(), C.c_uint()
        check(self._L.velvet_solver_check_nan(self._h, C.byref(cnt), C.byref(first)))
        return cnt.value, first.value

    def Grab(self, rayOrigin, rayDirection):
        """MouseGrabber::HandleMouseInteraction, mouse-down branch (MouseGrabber.hpp L40-55), on the device: returns
        (index of the picked particle or -1, its distance along the ray)."""
        o = np.ascontiguousarray(rayOrigin, np.float32)
        d = np.ascontiguousarray(rayDirection, np.float32)
        idx, dist = C.c_int(-1), C.c_float(0)
        self._L.velvet_solver_grab.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float)]
        check(self._L.velvet_solver_grab(self._h, _ptr(o), _ptr(d), C.byref(idx), C.byref(dist)))
        return idx.value, np.float32(dist.value)

    def Drag(self, rayOrigin, rayDirection):
        """MouseGrabber::UpdateGrappedVertex (L66-79) for this frame's ray."""
        o = np.ascontiguousarray(rayOrigin, np.float32)
        d = np.ascontiguousarray(rayDirection, np.float32)
        self._L.velvet_solver_drag.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        check(self._L.velvet_solver_drag(self._h, _ptr(o), _ptr(d)))

    def Release(self):
        """Mouse-up branch (L57-62)."""
        self._L.velvet_solver_release.argtypes = [C.c_void_p]
        check(self._L.velvet_solver_release(self._h))

    def buffer_ptr(self, name: str):
        p = C.c_void_p()
        n = C.c_size_t()
        check(self._L.velvet_solver_buffer(self._h, BUFFER_IDS[name], C.byref(p), C.byref(n)))
        return p.value or 0, n.value

    def download(self, name: str) -> np.ndarray:
        _, n = self.buffer_ptr(name)
        if name == "sdfColliders":
            out = np.zeros(n * 196, np.uint8)
        else:
            dt, w = _BUF_DTYPE[name]
            out = np.zeros((n, w) if w > 1 else n, dt)
        if out.nbytes:
            check(self._L.velvet_solver_download(self._h, BUFFER_IDS[name], _ptr(out), out.nbytes))
        return out

    def upload(self, name: str, data):
        dt, _ = _BUF_DTYPE[name]
        data = np.ascontiguousarray(data, dt)
        check(self._L.velvet_solver_upload(self._h, BUFFER_IDS[name], _ptr(data), data.nbytes))


class VtClothObjectGPU:
    """Constraint generation for a grid cloth (VtClothObjectGPU.hpp L43-148)."""

    def __init__(self, resolution: int, solver: VtClothSolverGPU):
        self.resolution = resolution
        self.solver = solver
        self.attached = []
        self.indexOffset = 0

    def SetAttachedIndices(self, indices):
        self.attached = list(indices)

    def Start(self, vertices, indices, modelMatrix):
        vertices = np.ascontiguousarray(vertices, np.float32)
        indices = np.ascontiguousarray(indices, np.uint32)
        att = np.asarray(self.attached, np.int32)
        off = C.c_int(0)
        check(self.solver._L.velvet_cloth_object_start(self.solver._h, self.resolution, _ptr(vertices), _ptr(indices),
                                                       _ptr(np.ascontiguousarray(modelMatrix, np.float32)), _ptr(att),
                                                       len(att), C.byref(off)))
        self.indexOffset = off.value
        return off.value


class SpatialHashGPU:
    """SpatialHashGPU.hpp L15-60 as a stand-alone object (positions are device pointers or numpy arrays)."""

    def __init__(self, particleDiameter: float, maxNumObjects: int, hashCellSizeScalar: float = 1.5,
                 maxNumNeighbors: int = 64):
        self._L = _capi.load()
        h = C.c_void_p()
        check(self._L.velvet_hash_create(C.byref(h), float(np.float32(particleDiameter)), maxNumObjects,
                                         float(np.float32(hashCellSizeScalar)), maxNumNeighbors))
        self._h = h
        self.maxNumObjects = maxNumObjects
        self.maxNumNeighbors = maxNumNeighbors

    def close(self):
        if getattr(self, "_h", None):
            self._L.velvet_hash_destroy(self._h)
            self._h = None

    __del__ = close

    def SetInitialPositions(self, positions):
        positions = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
        check(self._L.velvet_hash_set_initial_positions(self._h, _ptr(positions), len(positions)))

    def Hash(self, positions):
        """positions: numpy float32 [n,3] (copied to the device) or (device_ptr, n)."""
        if isinstance(positions, tuple):
            check(self._L.velvet_hash_hash(self._h, positions[0], positions[1]))
            return
        positions = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
        dev = C.c_void_p()
        check(self._L.velvet_alloc(C.byref(dev), positions.nbytes))
        try:
            check(self._L.velvet_copy(dev, _ptr(positions), positions.nbytes))
            check(self._L.velvet_hash_hash(self._h, dev, len(positions)))
        finally:
            self._L.velvet_free(dev)

    def download(self, name: str) -> np.ndarray:
        p = C.c_void_p()
        n = C.c_size_t()
        check(self._L.velvet_hash_buffer(self._h, BUFFER_IDS[name], C.byref(p), C.byref(n)))
        dt, w = _BUF_DTYPE[name]
        out = np.zeros((n.value, w) if w > 1 else n.value, dt)
        if out.nbytes:
            check(self._L.velvet_copy(_ptr(out), p, out.nbytes))
        return out


def build_scene(resolution: int, params: VtSimParams | None = None, position=(0, 1.5, 1.0), rotation=(90, 0, 0),
                attached=(), pipeline: int = PIPELINE_FUSED, device: int = -1, tile_size: int = 0,
                math_mode: int = MATH_EXACT) -> VtClothSolverGPU:
    """SpawnCloth + Initialize + VtClothObjectGPU::Start for one grid cloth (Scene.hpp L251-291, main.cpp L141)."""
    solver = VtClothSolverGPU(params, device=device, pipeline=pipeline, tile_size=tile_size, math_mode=math_mode)
    v, idx = GenerateClothMesh(resolution)
    obj = VtClothObjectGPU(resolution, solver)
    obj.SetAttachedIndices(attached)
    obj.Start(v, idx, TransformMatrix(position, rotation, (1, 1, 1)))
    return solver


def sphere_plane_colliders(t: float | None = None, radius: float = 0.6):
    """Plane at the origin + sphere: static at (0, r, 0) (main.cpp L294-296) or, when t is given, moving as
    (0, r, -cos 2t) (main.cpp L162-167)."""
    plane = MakeCollider(COLLIDER_PLANE, (0, 0, 0), (1, 1, 1))
    z = 0.0 if t is None else -math.cos(2.0 * t)
    sphere = MakeCollider(COLLIDER_SPHERE, (0, radius, z), (radius, radius, radius))
    return [plane, sphere]
