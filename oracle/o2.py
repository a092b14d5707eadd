"""ctypes binding of O2 (oracle/ref_gs_cpu.c): the reference's own CPU solver restated headless.
TEST / BASELINE INFRASTRUCTURE ONLY (bench.py cpu_baseline and --impl reference, tests/)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import o1

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libo2_ref_gs_cpu.so")
        src = os.path.join(_HERE, "ref_gs_cpu.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE, "libo2_ref_gs_cpu.so"], stdout=subprocess.DEVNULL)
        L = C.CDLL(so)
        v = C.c_void_p
        L.o2_create.restype = v
        L.o2_create.argtypes = [v, C.c_int, v, v, v, v, C.c_int]
        L.o2_destroy.argtypes = [v]
        L.o2_set_colliders.argtypes = [v, v, v, v, C.c_int]
        L.o2_simulate.argtypes = [v]
        L.o2_positions.restype = v
        L.o2_positions.argtypes = [v]
        L.o2_normals.restype = v
        L.o2_normals.argtypes = [v]
        L.o2_num_particles.argtypes = [v]
        _LIB = L
    return _LIB


class O2Solver:
    """VtClothSolverCPU (single-thread Gauss-Seidel)."""

    def __init__(self, params: o1.SimParams, resolution: int, model16, attached=()):
        self._L = lib()
        v, idx = o1.generate_cloth_mesh(resolution)
        att = np.asarray(list(attached), np.int32)
        f = lambda a: a.ctypes.data_as(C.c_void_p)
        self._h = self._L.o2_create(C.byref(params), resolution, f(v), f(idx), f(np.ascontiguousarray(model16, np.float32)),
                                    f(att), len(att))
        self.n = self._L.o2_num_particles(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.o2_destroy(self._h)
            self._h = None

    def set_colliders(self, types, positions, scales_x):
        t = np.asarray(types, np.int32)
        p = np.ascontiguousarray(positions, np.float32)
        s = np.asarray(scales_x, np.float32)
        f = lambda a: a.ctypes.data_as(C.c_void_p)
        self._L.o2_set_colliders(self._h, f(t), f(p), f(s), len(t))

    def simulate(self):
        self._L.o2_simulate(self._h)

    @property
    def positions(self) -> np.ndarray:
        p = self._L.o2_positions(self._h)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(self.n, 3))
