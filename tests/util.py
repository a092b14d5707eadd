"""Shared helpers: build the same scene for the CUDA solver (through the C ABI) and for the O1 oracle."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

import velvet_b200 as vb
from oracle import o1

EXTENT = 2.0  # cloth size (Scene.hpp L137); tolerances are stated as a fraction of it


def gpu_params(**kw) -> vb.VtSimParams:
    p = vb.default_params()
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def to_o1_params(p: vb.VtSimParams) -> o1.SimParams:
    q = o1.SimParams()
    C.memmove(C.byref(q), C.byref(p), 80)
    return q


def to_o1_collider(c: vb.VtSDFCollider) -> o1.SDFCollider:
    q = o1.SDFCollider()
    C.memmove(C.byref(q), C.byref(c), 196)
    return q


class ColliderTrack:
    """Collider::FixedUpdate (Collider.hpp L33-41): lastTransform <- cur, cur <- matrix()."""

    def __init__(self, ctype, position, scale, rotation=(0, 0, 0)):
        self.ctype = ctype
        self.scale = scale
        self.cur = vb.TransformMatrix(position, rotation, scale)
        self.last = self.cur.copy()
        self.position = position

    def move(self, position, rotation=(0, 0, 0)):
        self.last = self.cur
        self.cur = vb.TransformMatrix(position, rotation, self.scale)
        self.position = position

    def collider(self) -> vb.VtSDFCollider:
        return vb.MakeCollider(self.ctype, self.position, self.scale, self.cur, self.last)


def make_pair(resolution, params=None, position=(0, 1.5, 1.0), rotation=(90, 0, 0), attached=(), pipeline=vb.PIPELINE_FUSED,
              tile_size=0, oracle=True, math_mode=vb.MATH_EXACT):
    """Same grid cloth registered in the CUDA solver and (optionally) in the O1 oracle.  EXACT math (the library default)
    is bit-identical to the oracle; the opt-in FAST mode is tested at tolerance."""
    params = params or gpu_params()
    g = vb.build_scene(resolution, params, position, rotation, attached, pipeline=pipeline, tile_size=tile_size,
                       math_mode=math_mode)
    o = None
    if oracle:
        o = o1.O1Solver(to_o1_params(params))
        v, idx = o1.generate_cloth_mesh(resolution)
        o.cloth_object_start(resolution, v, idx, o1.transform_matrix(position, rotation, (1, 1, 1)), attached)
    return g, o


def set_colliders(g, o, colliders):
    g.UpdateColliders(colliders)
    if o is not None:
        o.set_colliders([to_o1_collider(c) for c in colliders])


def max_abs_diff(a, b) -> float:
    a = np.asarray(a, np.float64).reshape(-1)
    b = np.asarray(b, np.float64).reshape(-1)
    assert a.shape == b.shape
    return float(np.max(np.abs(a - b))) if a.size else 0.0


def neighbor_lists(neighbors: np.ndarray, n: int, k: int):
    """Column-major table [i + n*k] -> list of per-particle arrays, cut at the 0xffffffff terminator."""
    tab = neighbors[: n * k].reshape(k, n)
    out = []
    for i in range(n):
        col = tab[:, i]
        stop = np.nonzero(col == 0xFFFFFFFF)[0]
        out.append(col[: stop[0]] if len(stop) else col)
    return out


def valid_prefix_table(neighbors: np.ndarray, n: int, k: int) -> np.ndarray:
    """Neighbour table with everything after each column's terminator masked (stale entries are never read)."""
    tab = neighbors[: n * k].reshape(k, n).copy()
    term = tab == 0xFFFFFFFF
    after = np.cumsum(term, axis=0) > 0
    tab[after] = 0xFFFFFFFF
    return tab
