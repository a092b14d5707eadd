"""ctypes binding of the O1 parity oracle (oracle/ref_jacobi_cpu.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never imported by the product package velvet_b200.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class SimParams(C.Structure):
    """VtSimParams, Common.hpp L19-47 (80 bytes)."""
    _fields_ = [
        ("numSubsteps", C.c_int32), ("numIterations", C.c_int32), ("maxNumNeighbors", C.c_int32),
        ("maxSpeed", C.c_float), ("gravity", C.c_float * 3), ("bendCompliance", C.c_float),
        ("damping", C.c_float), ("relaxationFactor", C.c_float), ("longRangeStretchiness", C.c_float),
        ("collisionMargin", C.c_float), ("friction", C.c_float), ("enableSelfCollision", C.c_uint8),
        ("_pad", C.c_uint8 * 3), ("interleavedHash", C.c_int32), ("numParticles", C.c_uint32),
        ("particleDiameter", C.c_float), ("deltaTime", C.c_float), ("particleDiameterScalar", C.c_float),
        ("hashCellSizeScalar", C.c_float),
    ]


class SDFCollider(C.Structure):
    """SDFCollider, VtClothSolverGPU.cuh L8-18 (196 bytes)."""
    _fields_ = [
        ("type", C.c_int32), ("position", C.c_float * 3), ("scale", C.c_float * 3),
        ("deltaTime", C.c_float), ("curTransform", C.c_float * 9), ("invCurTransform", C.c_float * 16),
        ("lastTransform", C.c_float * 16),
    ]


class HashParams(C.Structure):
    """HashParams, SpatialHashGPU.cuh L7-15 (24 bytes)."""
    _fields_ = [
        ("numObjects", C.c_uint32), ("maxNumNeighbors", C.c_uint32), ("cellSpacing", C.c_float),
        ("cellSpacing2", C.c_float), ("tableSize", C.c_int32), ("particleDiameter2", C.c_float),
    ]


assert C.sizeof(SimParams) == 80 and C.sizeof(SDFCollider) == 196 and C.sizeof(HashParams) == 24

SPHERE, PLANE, CUBE = 0, 1, 2

BUF = {name: i for i, name in enumerate([
    "positions", "normals", "indices", "velocities", "predicted", "deltas", "deltaCounts", "invMasses",
    "stretchIndices", "stretchLengths", "bendIndices", "bendAngles", "attachParticleIDs", "attachSlotIDs",
    "attachDistances", "attachSlotPositions", "neighbors", "initialPositions", "particleHash",
    "particleIndex", "cellStart", "cellEnd"])}
_DTYPE = {"indices": np.uint32, "deltaCounts": np.int32, "stretchIndices": np.int32, "bendIndices": np.uint32,
          "attachParticleIDs": np.int32, "attachSlotIDs": np.int32, "neighbors": np.uint32,
          "particleHash": np.uint32, "particleIndex": np.uint32, "cellStart": np.uint32, "cellEnd": np.uint32}


def _host_has_fma() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return " fma " in (line + " ")
    except OSError:
        pass
    return False


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (seconds). Returns the .so path: the -mfma build (explicit fmaf() calls inlined) when the
    host CPU has FMA, else the portable one -- both produce the same bits."""
    name = "libo1_ref_jacobi_cpu_fma.so" if _host_has_fma() else "libo1_ref_jacobi_cpu.so"
    so = os.path.join(_HERE, name)
    src = os.path.join(_HERE, "ref_jacobi_cpu.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, name], stdout=subprocess.DEVNULL)
    return so


def _fp(a):
    return a.ctypes.data_as(C.c_void_p)


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.o1_solver_create.restype = C.c_void_p
        L.o1_solver_create.argtypes = [C.c_void_p]
        L.o1_solver_params.restype = C.POINTER(SimParams)
        L.o1_solver_params.argtypes = [C.c_void_p]
        L.o1_solver_buffer.restype = C.c_void_p
        L.o1_solver_buffer.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint64)]
        L.o1_solver_destroy.argtypes = [C.c_void_p]
        L.o1_solver_simulate.argtypes = [C.c_void_p]
        L.o1_solver_hash.argtypes = [C.c_void_p]
        L.o1_solver_add_cloth.restype = C.c_int
        L.o1_solver_add_cloth.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_float]
        L.o1_solver_add_stretch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float]
        L.o1_solver_add_attach_slot.argtypes = [C.c_void_p, C.c_void_p]
        L.o1_solver_add_attach.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float]
        L.o1_solver_add_bend.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float]
        L.o1_solver_set_colliders.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.o1_cloth_object_start.restype = C.c_int
        L.o1_cloth_object_start.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.o1_generate_cloth_mesh.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.o1_transform_matrix.argtypes = [C.c_void_p] * 4
        L.o1_mat4_inverse.argtypes = [C.c_void_p] * 2
        L.o1_make_collider.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
        L.o1_default_params.argtypes = [C.POINTER(SimParams)]
        L.o1_hash_objects.argtypes = [C.c_void_p] * 7 + [HashParams]
        L.o1_hash_position.restype = C.c_int
        L.o1_hash_position.argtypes = [C.c_void_p, C.c_float, C.c_int]
        P = C.POINTER(SimParams)
        v = C.c_void_p
        L.o1_initialize_positions.argtypes = [v, C.c_int, C.c_int, v]
        L.o1_predict_positions.argtypes = [P, v, v, v, C.c_float]
        L.o1_solve_stretch.argtypes = [v, v, v, v, v, v, C.c_uint32]
        L.o1_solve_bending.argtypes = [P, v, v, v, v, v, v, C.c_uint32, C.c_float]
        L.o1_solve_attachment.argtypes = [P, v, v, v, v, v, v, v, v, C.c_int]
        L.o1_apply_deltas.argtypes = [P, v, v, v]
        L.o1_collide_sdf.argtypes = [P, v, v, v, C.c_uint32, C.c_float]
        L.o1_collide_particles.argtypes = [P, v, v, v, v, v, v]
        L.o1_finalize.argtypes = [P, v, v, v, C.c_float]
        L.o1_compute_normal.argtypes = [P, v, v, v, C.c_uint32]
        _LIB = L
    return _LIB


def default_params() -> SimParams:
    p = SimParams()
    lib().o1_default_params(C.byref(p))
    return p


def generate_cloth_mesh(resolution: int):
    """Scene.hpp L131-168 -> (vertices float32 [(R+1)^2,3], indices uint32 [6 R^2])."""
    n = (resolution + 1) ** 2
    v = np.zeros((n, 3), np.float32)
    idx = np.zeros(6 * resolution * resolution, np.uint32)
    lib().o1_generate_cloth_mesh(resolution, _fp(v), _fp(idx))
    return v, idx


def transform_matrix(position=(0, 0, 0), rotation_deg=(0, 0, 0), scale=(1, 1, 1)) -> np.ndarray:
    """Transform::matrix(), column-major float32[16]."""
    out = np.zeros(16, np.float32)
    lib().o1_transform_matrix(_fp(np.asarray(position, np.float32)), _fp(np.asarray(rotation_deg, np.float32)),
                              _fp(np.asarray(scale, np.float32)), _fp(out))
    return out


def mat4_inverse(m16) -> np.ndarray:
    out = np.zeros(16, np.float32)
    lib().o1_mat4_inverse(_fp(np.ascontiguousarray(m16, np.float32)), _fp(out))
    return out


def make_collider(ctype: int, position, scale, cur16=None, last16=None, dt: float = 1.0 / 60.0) -> SDFCollider:
    """UpdateColliders body (VtClothSolverGPU.hpp L195-203)."""
    position = np.asarray(position, np.float32)
    scale = np.asarray(scale, np.float32)
    if cur16 is None:
        cur16 = transform_matrix(position, (0, 0, 0), scale)
    if last16 is None:
        last16 = cur16
    c = SDFCollider()
    lib().o1_make_collider(ctype, _fp(position), _fp(scale), _fp(np.ascontiguousarray(cur16, np.float32)),
                           _fp(np.ascontiguousarray(last16, np.float32)), np.float32(dt), C.byref(c))
    return c


def colliders_array(cols):
    arr = (SDFCollider * max(len(cols), 1))()
    for i, c in enumerate(cols):
        arr[i] = c
    return arr


class O1Solver:
    """VtClothSolverGPU restated on the CPU (sequential Jacobi)."""

    def __init__(self, params: SimParams | None = None):
        self._L = lib()
        self._h = self._L.o1_solver_create(C.byref(params) if params is not None else None)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.o1_solver_destroy(self._h)
            self._h = None

    @property
    def params(self) -> SimParams:
        return self._L.o1_solver_params(self._h).contents

    def buffer(self, name: str) -> np.ndarray:
        n = C.c_uint64(0)
        p = self._L.o1_solver_buffer(self._h, BUF[name], C.byref(n))
        if not p or n.value == 0:
            return np.zeros(0, _DTYPE.get(name, np.float32))
        arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(n.value,))
        return arr.view(_DTYPE.get(name, np.float32))

    def add_cloth(self, vertices, indices, model16, diameter) -> int:
        vertices = np.ascontiguousarray(vertices, np.float32)
        indices = np.ascontiguousarray(indices, np.uint32)
        return self._L.o1_solver_add_cloth(self._h, _fp(vertices), len(vertices), _fp(indices), len(indices),
                                           _fp(np.ascontiguousarray(model16, np.float32)), np.float32(diameter))

    def add_stretch(self, i, j, d):
        self._L.o1_solver_add_stretch(self._h, i, j, np.float32(d))

    def add_attach_slot(self, pos):
        self._L.o1_solver_add_attach_slot(self._h, _fp(np.asarray(pos, np.float32)))

    def add_attach(self, particle, slot, dist):
        self._L.o1_solver_add_attach(self._h, particle, slot, np.float32(dist))

    def add_bend(self, a, b, c, d, angle=0.0):
        self._L.o1_solver_add_bend(self._h, a, b, c, d, np.float32(angle))

    def cloth_object_start(self, resolution, vertices, indices, model16, attached=()):
        vertices = np.ascontiguousarray(vertices, np.float32)
        indices = np.ascontiguousarray(indices, np.uint32)
        att = np.asarray(list(attached), np.int32)
        return self._L.o1_cloth_object_start(self._h, resolution, _fp(vertices), _fp(indices),
                                             _fp(np.ascontiguousarray(model16, np.float32)), _fp(att), len(att))

    def set_colliders(self, cols):
        arr = colliders_array(cols)
        self._L.o1_solver_set_colliders(self._h, C.cast(arr, C.c_void_p), len(cols))

    def simulate(self):
        self._L.o1_solver_simulate(self._h)

    def hash(self):
        self._L.o1_solver_hash(self._h)
