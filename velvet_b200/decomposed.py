"""One large cloth decomposed over several GPUs (north_star mode 2): one process per GPU (torchrun), every rank registers
the same cloth, rank r runs the Jacobi iterations of its own tile range, the boundary particles move over NVLink once
per iteration and the predicted positions are all-gathered once per substep.  Two transports:

* "peer" (default): the library's own kernels store boundary particles straight into the peers' arrays (CUDA IPC mappings)
  and order them with epoch flags; a frame is one CUDA graph per rank (velvet_solver_dd_simulate).  torch.distributed is
  used once, to all-gather the 216-byte IPC blobs.
* "nccl": the stepped C ABI (velvet_solver_dd_step) with torch.distributed send/recv + all_gather between the steps, on
  tensors that alias the solver's device buffers.  Fallback when IPC mappings are unavailable, and the cross-check.

Both are bit-identical to the single-GPU solver.
"""
from __future__ import annotations

import ctypes as C

from . import _capi
from ._capi import check

DD_FRAME_BEGIN, DD_SUBSTEP_BEGIN, DD_ITERATE_OWNED, DD_ITERATE_FINISH = 0, 1, 2, 3
DD_GATHER_PACK, DD_GATHER_UNPACK, DD_SUBSTEP_END, DD_FRAME_END = 4, 5, 6, 7


class VelvetDDInfo(C.Structure):
    _fields_ = [("rank", C.c_int), ("world", C.c_int), ("tileBegin", C.c_uint), ("tileEnd", C.c_uint), ("numTiles", C.c_uint),
                ("ownedCount", C.c_uint), ("maxOwnedCount", C.c_uint), ("sendTotal", C.c_uint), ("recvTotal", C.c_uint),
                ("sendBuf", C.c_void_p), ("recvBuf", C.c_void_p), ("gatherSend", C.c_void_p), ("gatherRecv", C.c_void_p)]


def plan_grid(resolution: int, rank: int, world: int, tile_size: int = 0):
    """Host-only exchange lists of `rank` (no GPU): (send {peer: ids}, recv {peer: ids}, (ownedBegin, ownedEnd))."""
    import numpy as np
    L = _capi.load()
    L.velvet_dd_plan_grid.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    counts = np.zeros(2 * world, np.uint32)
    check(L.velvet_dd_plan_grid(resolution, tile_size, rank, world, counts.ctypes.data_as(C.c_void_p), None, None, None))
    send = np.zeros(max(int(counts[:world].sum()), 1), np.uint32)
    recv = np.zeros(max(int(counts[world:].sum()), 1), np.uint32)
    owned = np.zeros(2, np.uint32)
    check(L.velvet_dd_plan_grid(resolution, tile_size, rank, world, counts.ctypes.data_as(C.c_void_p), send.ctypes.data_as(C.c_void_p),
                                recv.ctypes.data_as(C.c_void_p), owned.ctypes.data_as(C.c_void_p)))
    so = np.concatenate([[0], np.cumsum(counts[:world])]).astype(int)
    ro = np.concatenate([[0], np.cumsum(counts[world:])]).astype(int)
    return ({q: send[so[q]:so[q + 1]].copy() for q in range(world)}, {q: recv[ro[q]:ro[q + 1]].copy() for q in range(world)},
            (int(owned[0]), int(owned[1])))


class _DevicePtr:
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can alias it (no copy)."""

    def __init__(self, ptr: int, nfloats: int):
        self.__cuda_array_interface__ = {"shape": (nfloats,), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


class DecomposedCloth:
    def __init__(self, solver, device_index: int, transport: str = "peer"):
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised (backend nccl) before decomposing a cloth")
        self.torch, self.dist = torch, dist
        self.solver = solver
        self._L = _capi.load()
        self._L.velvet_solver_dd_setup.argtypes = [C.c_void_p, C.c_int, C.c_int]
        self._L.velvet_solver_dd_info.argtypes = [C.c_void_p, C.POINTER(VelvetDDInfo)]
        self._L.velvet_solver_dd_offsets.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        self._L.velvet_solver_dd_step.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float]
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        check(self._L.velvet_solver_dd_setup(solver._h, self.rank, self.world))
        self.info = VelvetDDInfo()
        check(self._L.velvet_solver_dd_info(solver._h, C.byref(self.info)))
        self._dev = torch.device("cuda", device_index)
        self.stream = torch.cuda.ExternalStream(solver.stream, device=self._dev)
        self._stepped_ready = False
        self.halo_bytes_per_iteration = 16 * (self.info.sendTotal + self.info.recvTotal)
        if transport not in ("peer", "nccl"):
            raise ValueError("transport must be 'peer' or 'nccl'")
        self.transport = "nccl"
        self.peer_error = None
        if transport == "peer" and self.world > 1:
            self._map_peers(self._dev)

    def _map_peers(self, dev):
        """All-gather the IPC blobs and map the peers' arrays; every rank agrees on the outcome (peer, or nccl fallback)."""
        torch, dist, L = self.torch, self.dist, self._L
        L.velvet_dd_peer_blob_bytes.restype = C.c_size_t
        L.velvet_solver_dd_peer_export.argtypes = [C.c_void_p, C.c_void_p]
        L.velvet_solver_dd_peer_import.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.velvet_solver_dd_peer_close.argtypes = [C.c_void_p]
        L.velvet_solver_dd_simulate.argtypes = [C.c_void_p, C.c_float, C.c_int]
        nbytes = int(L.velvet_dd_peer_blob_bytes())
        blob = (C.c_ubyte * nbytes)()
        check(L.velvet_solver_dd_peer_export(self.solver._h, blob))
        mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
        everyone = torch.empty(nbytes * self.world, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(everyone, mine)
        raw = everyone.cpu().numpy().tobytes()
        ok = 1
        try:
            check(L.velvet_solver_dd_peer_import(self.solver._h, raw, nbytes))
        except _capi.VelvetError as e:
            self.peer_error = str(e)
            ok = 0
        agreed = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(agreed, op=dist.ReduceOp.MIN)
        if int(agreed.item()) == 1:
            self.transport = "peer"
        else:
            check(L.velvet_solver_dd_peer_close(self.solver._h))

    def _ensure_stepped(self):
        """Exchange lists and staging buffers of the tile form (a strip-decomposed cloth sets them up only when the stepped
        schedule is really used), aliased as torch tensors for the NCCL calls."""
        if self._stepped_ready:
            return
        torch = self.torch
        self._L.velvet_solver_dd_prepare_stepped.argtypes = [C.c_void_p]
        check(self._L.velvet_solver_dd_prepare_stepped(self.solver._h))
        check(self._L.velvet_solver_dd_info(self.solver._h, C.byref(self.info)))
        so = (C.c_uint * (self.world + 1))()
        ro = (C.c_uint * (self.world + 1))()
        check(self._L.velvet_solver_dd_offsets(self.solver._h, so, ro))
        self.send_off, self.recv_off = list(so), list(ro)
        alias = lambda ptr, n: torch.as_tensor(_DevicePtr(ptr, 4 * max(n, 1)), device=self._dev)
        self.send = alias(self.info.sendBuf, self.info.sendTotal)
        self.recv = alias(self.info.recvBuf, self.info.recvTotal)
        self.gather_send = alias(self.info.gatherSend, self.info.maxOwnedCount)
        self.gather_recv = alias(self.info.gatherRecv, self.info.maxOwnedCount * self.world)
        self.halo_bytes_per_iteration = 16 * (self.info.sendTotal + self.info.recvTotal)
        self._stepped_ready = True

    def close(self):
        """Unmap the peers (collective: every rank must call it before any rank destroys its solver)."""
        if self.transport == "peer":
            self.solver.Synchronize()
            check(self._L.velvet_solver_dd_peer_close(self.solver._h))
            self.dist.barrier()
            self.transport = "closed"

    def _step(self, op: int, arg: int = 0, farg: float = 0.0):
        check(self._L.velvet_solver_dd_step(self.solver._h, op, arg, farg))

    def _exchange_halo(self):
        dist = self.dist
        ops = []
        for q in range(self.world):
            if q == self.rank:
                continue
            s0, s1 = 4 * self.send_off[q], 4 * self.send_off[q + 1]
            r0, r1 = 4 * self.recv_off[q], 4 * self.recv_off[q + 1]
            if s1 > s0:
                ops.append(dist.P2POp(dist.isend, self.send[s0:s1], q))
            if r1 > r0:
                ops.append(dist.P2POp(dist.irecv, self.recv[r0:r1], q))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()  # orders the solver stream after the transfer; does not block the host

    def Simulate(self, dt: float = 1.0 / 60.0, sync: bool = True):
        if self.transport == "peer":
            check(self._L.velvet_solver_dd_simulate(self.solver._h, dt, 1 if sync else 0))
            return
        if self.transport == "closed":
            raise RuntimeError("DecomposedCloth.close() was called")
        self._ensure_stepped()
        P = self.solver.simParams
        with self.torch.cuda.stream(self.stream):
            self._step(DD_FRAME_BEGIN, 0, dt)
            for s in range(P.numSubsteps):
                self._step(DD_SUBSTEP_BEGIN, s)  # hash + collide of the owned particles, boundary packed
                self._exchange_halo()
                self._step(DD_ITERATE_FINISH)
                for _ in range(P.numIterations):
                    self._step(DD_ITERATE_OWNED)
                    self._exchange_halo()
                    self._step(DD_ITERATE_FINISH)
                if self.world > 1:
                    self._step(DD_GATHER_PACK)
                    self.dist.all_gather_into_tensor(self.gather_recv, self.gather_send)
                    self._step(DD_GATHER_UNPACK)
                self._step(DD_SUBSTEP_END, s)
            self._step(DD_FRAME_END)
        if sync:
            self.solver.Synchronize()


class LocalShards:
    """The decomposed cloth with all G shards on ONE device in ONE process: G solvers carrying the same cloth, solver r set
    up as rank r of G, driven through the same stepped C ABI as the NCCL transport, with the halo exchange and the
    per-substep all-gather done as device-to-device copies between the solvers' buffers.  No throughput to be had -- it
    exists so that the decomposition logic (ownership, exchange lists, owned-only collide / neighbour walk, the stepped
    schedule) is checked against the single-GPU solver on a one-GPU box (tests/test_decomposed_gpu.py)."""

    def __init__(self, solvers, device_index: int = 0):
        import torch
        self.torch = torch
        self.solvers = list(solvers)
        self.world = len(self.solvers)
        self._L = _capi.load()
        self._L.velvet_solver_dd_setup.argtypes = [C.c_void_p, C.c_int, C.c_int]
        self._L.velvet_solver_dd_info.argtypes = [C.c_void_p, C.POINTER(VelvetDDInfo)]
        self._L.velvet_solver_dd_offsets.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        self._L.velvet_solver_dd_step.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float]
        dev = torch.device("cuda", device_index)
        alias = lambda ptr, n: torch.as_tensor(_DevicePtr(ptr, 4 * max(n, 1)), device=dev)
        self.info, self.send_off, self.recv_off = [], [], []
        self.send, self.recv, self.gather_send, self.gather_recv = [], [], [], []
        self._L.velvet_solver_dd_prepare_stepped.argtypes = [C.c_void_p]
        for r, s in enumerate(self.solvers):
            check(self._L.velvet_solver_dd_setup(s._h, r, self.world))
            check(self._L.velvet_solver_dd_prepare_stepped(s._h))
            info = VelvetDDInfo()
            check(self._L.velvet_solver_dd_info(s._h, C.byref(info)))
            so = (C.c_uint * (self.world + 1))()
            ro = (C.c_uint * (self.world + 1))()
            check(self._L.velvet_solver_dd_offsets(s._h, so, ro))
            self.info.append(info)
            self.send_off.append(list(so))
            self.recv_off.append(list(ro))
            self.send.append(alias(info.sendBuf, info.sendTotal))
            self.recv.append(alias(info.recvBuf, info.recvTotal))
            self.gather_send.append(alias(info.gatherSend, info.maxOwnedCount))
            self.gather_recv.append(alias(info.gatherRecv, info.maxOwnedCount * self.world))

    def _all(self, op, arg=0, farg=0.0):
        for s in self.solvers:
            check(self._L.velvet_solver_dd_step(s._h, op, arg, farg))

    def _sync(self):
        for s in self.solvers:
            s.Synchronize()
        self.torch.cuda.synchronize()

    def _exchange_halo(self):
        self._sync()
        for r in range(self.world):
            for q in range(self.world):
                if q == r:
                    continue
                s0, s1 = 4 * self.send_off[r][q], 4 * self.send_off[r][q + 1]
                r0, r1 = 4 * self.recv_off[q][r], 4 * self.recv_off[q][r + 1]
                assert s1 - s0 == r1 - r0, "exchange lists must be symmetric"
                if s1 > s0:
                    self.recv[q][r0:r1].copy_(self.send[r][s0:s1])
        self._sync()

    def _all_gather(self):
        self._sync()
        m = 4 * self.info[0].maxOwnedCount
        for q in range(self.world):
            for r in range(self.world):
                self.gather_recv[q][r * m:(r + 1) * m].copy_(self.gather_send[r][:m])
        self._sync()

    def Simulate(self, dt: float = 1.0 / 60.0):
        P = self.solvers[0].simParams
        self._all(DD_FRAME_BEGIN, 0, dt)
        for sub in range(P.numSubsteps):
            self._all(DD_SUBSTEP_BEGIN, sub)
            self._exchange_halo()
            self._all(DD_ITERATE_FINISH)
            for _ in range(P.numIterations):
                self._all(DD_ITERATE_OWNED)
                self._exchange_halo()
                self._all(DD_ITERATE_FINISH)
            if self.world > 1:
                self._all(DD_GATHER_PACK)
                self._all_gather()
                self._all(DD_GATHER_UNPACK)
            self._all(DD_SUBSTEP_END, sub)
        self._all(DD_FRAME_END)
        self._sync()
