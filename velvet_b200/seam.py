"""The kernel seam with the reference's function names (VtClothSolverGPU.cuh L99-168, SpatialHashGPU.cuh L17-25).

Arguments are device pointers (int / ctypes.c_void_p) or any object with a ``data_ptr()`` method (torch CUDA
tensors); layouts are the reference's (packed float3).  Calls are asynchronous on the seam stream.
"""
from __future__ import annotations

import ctypes as C

from . import _capi
from ._capi import VtHashParams, VtSimParams, check


def _p(x):
    if x is None:
        return None
    if hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr())
    if isinstance(x, C.c_void_p):
        return x
    return C.c_void_p(int(x))


def SetSimulationParams(params: VtSimParams):
    check(_capi.load().velvet_SetSimulationParams(C.byref(params)))


def InitializePositions(positions, start: int, count: int, modelMatrix):
    import numpy as np
    m = np.ascontiguousarray(modelMatrix, np.float32)
    check(_capi.load().velvet_InitializePositions(_p(positions), start, count, m.ctypes.data_as(C.c_void_p)))


def PredictPositions(predicted, velocities, positions, deltaTime: float):
    check(_capi.load().velvet_PredictPositions(_p(predicted), _p(velocities), _p(positions), deltaTime))


def SolveStretch(predicted, deltas, deltaCounts, stretchIndices, stretchLengths, invMasses, numConstraints: int):
    check(_capi.load().velvet_SolveStretch(_p(predicted), _p(deltas), _p(deltaCounts), _p(stretchIndices),
                                           _p(stretchLengths), _p(invMasses), numConstraints))


def SolveBending(predicted, deltas, deltaCounts, bendingIndices, bendingAngles, invMass, numConstraints: int,
                 deltaTime: float):
    check(_capi.load().velvet_SolveBending(_p(predicted), _p(deltas), _p(deltaCounts), _p(bendingIndices),
                                           _p(bendingAngles), _p(invMass), numConstraints, deltaTime))


def SolveAttachment(predicted, deltas, deltaCounts, invMass, attachParticleIDs, attachSlotIDs, attachSlotPositions,
                    attachDistances, numConstraints: int):
    check(_capi.load().velvet_SolveAttachment(_p(predicted), _p(deltas), _p(deltaCounts), _p(invMass),
                                              _p(attachParticleIDs), _p(attachSlotIDs), _p(attachSlotPositions),
                                              _p(attachDistances), numConstraints))


def ApplyDeltas(predicted, deltas, deltaCounts):
    check(_capi.load().velvet_ApplyDeltas(_p(predicted), _p(deltas), _p(deltaCounts)))


def CollideSDF(predicted, colliders, positions, numColliders: int, deltaTime: float):
    check(_capi.load().velvet_CollideSDF(_p(predicted), _p(colliders), _p(positions), numColliders, deltaTime))


def CollideParticles(deltas, deltaCounts, predicted, invMasses, neighbors, positions):
    check(_capi.load().velvet_CollideParticles(_p(deltas), _p(deltaCounts), _p(predicted), _p(invMasses),
                                               _p(neighbors), _p(positions)))


def Finalize(velocities, positions, predicted, deltaTime: float):
    check(_capi.load().velvet_Finalize(_p(velocities), _p(positions), _p(predicted), deltaTime))


def ComputeNormal(normals, positions, indices, numTriangles: int):
    check(_capi.load().velvet_ComputeNormal(_p(normals), _p(positions), _p(indices), numTriangles))


def HashObjects(particleHash, particleIndex, cellStart, cellEnd, neighbors, positions, originalPositions,
                params: VtHashParams):
    check(_capi.load().velvet_HashObjects(_p(particleHash), _p(particleIndex), _p(cellStart), _p(cellEnd),
                                          _p(neighbors), _p(positions), _p(originalPositions), params))


def SortPairs(keys, values, numItems: int, endBit: int):
    check(_capi.load().velvet_SortPairs(_p(keys), _p(values), numItems, endBit))


def set_stream(stream):
    check(_capi.load().velvet_seam_set_stream(_p(stream) if stream else None))


def synchronize():
    check(_capi.load().velvet_device_synchronize())
