"""ctypes binding of O3 (oracle/_ref/libvelvet_refcuda.so): the reference's own CUDA kernels built by
oracle/ref_cuda/build_ref_cuda.sh.  TEST INFRASTRUCTURE ONLY; needs a GPU."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import o1

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libvelvet_refcuda.so")
# the same reference-side driver with the reference's two .cu files REPLACED by the drop-in shim over libvelvet_b200.so
# (velvet_b200/csrc/dropin/VelvetB200Shim.cpp; built by oracle/ref_cuda/build_ref_cuda.sh; tests/test_dropin_gpu.py)
SO_DROPIN = os.path.join(_HERE, "_ref", "libvelvet_dropin.so")
_LIBS = {}

BUF = dict(o1.BUF)
_DTYPE = {"indices": np.uint32, "deltaCounts": np.int32, "stretchIndices": np.int32, "bendIndices": np.uint32,
          "attachParticleIDs": np.int32, "attachSlotIDs": np.int32, "neighbors": np.uint32, "particleHash": np.uint32,
          "particleIndex": np.uint32, "cellStart": np.uint32, "cellEnd": np.uint32}


def available(so: str = SO) -> bool:
    return os.path.exists(so)


def lib(so: str = SO):
    if so not in _LIBS:
        L = C.CDLL(so)
        v = C.c_void_p
        L.refcuda_create.restype = v
        L.refcuda_create.argtypes = [v]
        L.refcuda_destroy.argtypes = [v]
        L.refcuda_params.restype = C.POINTER(o1.SimParams)
        L.refcuda_params.argtypes = [v]
        L.refcuda_add_cloth.restype = C.c_int
        L.refcuda_add_cloth.argtypes = [v, v, C.c_int, v, C.c_int, v, C.c_float]
        L.refcuda_add_stretch_bulk.argtypes = [v, v, v, C.c_size_t]
        L.refcuda_add_bend_bulk.argtypes = [v, v, v, C.c_size_t]
        L.refcuda_add_attach_slot.argtypes = [v, v]
        L.refcuda_add_attach_bulk.argtypes = [v, v, v, v, C.c_size_t]
        L.refcuda_set_colliders.argtypes = [v, v, C.c_int]
        L.refcuda_simulate.argtypes = [v]
        L.refcuda_hash_predicted.argtypes = [v]
        L.refcuda_label.restype = C.c_char_p
        L.refcuda_label.argtypes = [C.c_int]
        L.refcuda_label_ms.restype = C.c_double
        L.refcuda_label_ms.argtypes = [C.c_int]
        L.refcuda_buffer.restype = v
        L.refcuda_buffer.argtypes = [v, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        _LIBS[so] = L
    return _LIBS[so]


def _fp(a):
    return a.ctypes.data_as(C.c_void_p)


class RefCudaSolver:
    """VtClothSolverGPU driven through the reference's own kernels."""

    def __init__(self, params: o1.SimParams, so: str = SO):
        self._L = lib(so)
        self._h = self._L.refcuda_create(C.byref(params))

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.refcuda_destroy(self._h)
            self._h = None

    @property
    def params(self) -> o1.SimParams:
        return self._L.refcuda_params(self._h).contents

    def register_like(self, o: "o1.O1Solver", resolution, model16, attached=()):
        """Registers the same grid cloth: AddCloth with model-space vertices (the device applies the transform), then the
        constraint lists exactly as the oracle generated them (same values, same order)."""
        v, idx = o1.generate_cloth_mesh(resolution)
        d = np.float32(np.linalg.norm(v[0] - v[1]).astype(np.float32) * np.float32(self.params.particleDiameterScalar))
        self._L.refcuda_add_cloth(self._h, _fp(v), len(v), _fp(idx), len(idx), _fp(np.ascontiguousarray(model16, np.float32)), d)
        si, sl = o.buffer("stretchIndices").copy(), o.buffer("stretchLengths").copy()
        self._L.refcuda_add_stretch_bulk(self._h, _fp(si), _fp(sl), len(sl))
        nslots = len(o.buffer("attachSlotPositions")) // 3
        ap, asl, ad = (o.buffer(k).copy() for k in ("attachParticleIDs", "attachSlotIDs", "attachDistances"))
        slots = o.buffer("attachSlotPositions").reshape(-1, 3).copy()
        per = len(ad) // nslots if nslots else 0
        for s in range(nslots):  # AddAttachSlot then its AddAttach calls, slot by slot (VtClothObjectGPU.hpp L134-148)
            self._L.refcuda_add_attach_slot(self._h, _fp(slots[s]))
            a, b = s * per, (s + 1) * per
            self._L.refcuda_add_attach_bulk(self._h, _fp(ap[a:b].copy()), _fp(asl[a:b].copy()), _fp(ad[a:b].copy()), b - a)
        bi, ba = o.buffer("bendIndices").copy(), o.buffer("bendAngles").copy()
        self._L.refcuda_add_bend_bulk(self._h, _fp(bi), _fp(ba), len(ba))

    def set_colliders(self, cols):
        arr = o1.colliders_array(cols)
        self._L.refcuda_set_colliders(self._h, C.cast(arr, C.c_void_p), len(cols))

    def simulate(self):
        self._L.refcuda_simulate(self._h)

    def hash_predicted(self):
        self._L.refcuda_hash_predicted(self._h)

    def timers(self) -> dict:
        return {self._L.refcuda_label(i).decode(): self._L.refcuda_label_ms(i) for i in range(self._L.refcuda_num_labels())}

    def buffer(self, name: str) -> np.ndarray:
        """Managed memory viewed from the host (the reference's VtBuffer contract)."""
        n, es = C.c_size_t(), C.c_size_t()
        self._L.refcuda_sync()
        p = self._L.refcuda_buffer(self._h, BUF[name], C.byref(n), C.byref(es))
        words = n.value * es.value // 4
        if not p or not words:
            return np.zeros(0, _DTYPE.get(name, np.float32))
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(words,)).view(_DTYPE.get(name, np.float32))
