#!/usr/bin/env python
"""bench.py -- headline benchmark of the XPBD cloth hot path (BASELINE.json: particle-substeps/sec, 10 iterations,
self-collision, 1M particles; ms/frame).

    python bench.py --gpus N --steps K --warmup W             # this repo's sm_100a solver
    python bench.py --impl reference --gpus N --steps K ...    # the reference's own CPU solver (port, oracle/), SAME config

One "step" = one Simulate() frame (1/60 s: 5 substeps x 10 Jacobi iterations, hash rebuilt every 3rd substep)
of BASELINE.json configs[2]: a 1024x1024 cloth (1,048,576 particles) draped over an SDF sphere + plane with
particle self-collision.  Under torchrun (N > 1) every rank simulates its own independent cloth on its own
GPU -- the batched-independent-instances mode: no data-path collective, weak scaling.  Every line also carries
`sub_records` for the two multi-GPU configurations of BASELINE.json: configs[3] (4,096 independent 64x64 cloths sharded over
the N GPUs, strong scaling) and configs[4] (ONE 4096x4096 cloth decomposed over the N GPUs with NVLink halo exchange,
strong scaling; a plain single-GPU solver at N = 1), and at N = 1 a `ref_cuda` block: the reference's own CUDA kernels
(oracle/_ref, unmodified, compiled for sm_100a) timed on the same box and the same frames -- north_star's ">= 10x" denominator.

Timing: W >= 3 warm-up frames; K frames bracketed by barrier + synchronize on both sides, timed with CUDA
events recorded on the solver's stream, max over ranks.  `value` has every input resident in HBM; `e2e` goes
through the C-ABI object surface with HOST buffers: each frame uploads the collider block from pinned host
memory (UpdateColliders) and reads positions + normals back into pinned host memory (double-buffered, on a copy
stream, so the transfer of one frame overlaps the simulation of the next).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle-substeps/sec"
UNIT = "particle-substeps/s"
SUBSTEPS, ITERATIONS = 5, 10


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on fd 1 at
# communicator creation), so fd 1 is pointed at stderr for the whole run and the result line goes to the saved descriptor.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict):
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def algorithmic_bytes(N, S, B, A, rebuilds_per_frame, nbar):
    """bytes per particle per stage (fp32, 12-byte vec3, every array touched once)."""
    per_iter = 28 + 12 * S / N + 20 * B / N + 12 * A / N
    hash_rebuild = 20 + (4 + 16 * 3) + 20 + (44 + 4 * (nbar + 1))
    collide = 40 + 4 * (nbar + 1)
    per_substep = 48 + collide + 36 + ITERATIONS * per_iter + 48
    per_frame = SUBSTEPS * per_substep + rebuilds_per_frame * hash_rebuild + 24 + (24 + 12 * 2)
    return {"iterate_per_launch": N * per_iter, "frame": N * per_frame, "per_particle_iter": per_iter}


def workload_config(R, world, interleaved_hash=3):
    """The ONE description of the headline workload, shared verbatim by both arms (`--impl reference` prints the same dict)."""
    n = (R + 1) * (R + 1)
    return {"workload": (f"{R + 1}x{R + 1} cloth ({n} particles) self-colliding drape over SDF sphere + plane, "
                         f"{SUBSTEPS} substeps x {ITERATIONS} iterations, hash every {interleaved_hash} substeps"
                         + (f"; {world} independent cloths, one per GPU, no communication" if world > 1 else "")),
            "particles_per_gpu": n, "stretch": 4 * R * R + 2 * R, "bend": R * R, "attach": 0, "substeps": SUBSTEPS,
            "iterations": ITERATIONS}


# ---------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  NVML in a thread of this process (one sample every
    2 ms: a 60 ms timed region still gets ~30; `nvidia-smi -lms` needs longer than that just to start on an 8-GPU box), with
    nvidia-smi as the fallback when the NVML binding is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4),
               ("hw_power_brake_slowdown", 0x80))

    def __init__(self, gpu_index, torch=None):
        self.gpu = gpu_index
        self.p = self.f = None
        self.nvml = self.handle = self.thread = None
        self.samples, self.masks, self.running = [], [], False
        try:
            import pynvml
            pynvml.nvmlInit()
            handle = None
            if torch is not None:  # the device this rank computes on, whatever CUDA_VISIBLE_DEVICES did to the numbering
                try:
                    pr = torch.cuda.get_device_properties(gpu_index)
                    bus = "%08X:%02X:%02X.0" % (getattr(pr, "pci_domain_id", 0), pr.pci_bus_id, pr.pci_device_id)
                    handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
                except Exception:
                    handle = None
            self.handle = handle if handle is not None else pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _loop(self):
        nv = self.nvml
        while self.running:
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                self.masks.append(int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nvml is not None:
            import threading
            self.running = True
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
            return
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.nvml is not None:
            self.running = False
            if self.thread is not None:
                self.thread.join(timeout=2)
            if not self.samples:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"], "source": "nvml"}
            sm = sorted(self.samples)
            mask = 0
            for m in self.masks:
                mask |= m
            return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.smax, "reasons": sorted(n for n, bit in self.REASONS if mask & bit),
                    "samples": len(sm), "sm_min_mhz": sm[0], "source": "nvml, one sample per 2 ms inside the timed region"}
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, smax, reasons = [], [], set()
        for r in rows:
            try:
                r = [c.strip() for c in r]
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"], "source": "nvidia-smi"}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(smax), "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


# ---------------------------------------------------------------------------------------------- CPU reference arm
def run_cpu_reference(resolution, frames, warm_frames=0, budget_s=None):
    """Times O2 = the reference's CPU solver (VtClothSolverCPU, single-thread Gauss-Seidel) restated in oracle/ on this
    box's host cores.  Returns (particle_substeps_per_s, seconds_per_frame, n)."""
    from oracle import o1, o2
    p = o1.default_params()
    p.numSubsteps, p.numIterations = SUBSTEPS, ITERATIONS
    s = o2.O2Solver(p, resolution, o1.transform_matrix((0, 1.5, 1.0), (90, 0, 0), (1, 1, 1)))
    s.set_colliders([1, 0], [[0, 0, 0], [0, 0.6, 0]], [1.0, 0.6])
    # budget_s bounds the whole run (a 1024x1024 frame takes ~12 s on one core and cannot be split: the reference hashes once
    # per frame): the first frame is timed, then as many warm-up / timed frames as fit, at least one timed.  Throughput
    # (particle-substeps/s) does not depend on how many frames are averaged.
    done_warm = 0
    if budget_s is not None and warm_frames > 0:
        t0 = time.perf_counter()
        s.simulate()
        first = time.perf_counter() - t0
        done_warm = 1
        can = max(1, int(budget_s / max(first, 1e-9)) - 1)  # frames that still fit
        if warm_frames - 1 + frames > can:
            warm_frames = 1
            frames = max(1, min(frames, can))
    elif budget_s is not None:
        t0 = time.perf_counter()
        s.simulate()  # an untimed probe frame to size the run
        first = time.perf_counter() - t0
        frames = max(1, min(frames, int(budget_s / max(first, 1e-9)) - 1))
    for _ in range(warm_frames - done_warm):
        s.simulate()
    t0 = time.perf_counter()
    for _ in range(frames):
        s.simulate()
    dt = time.perf_counter() - t0
    if budget_s is not None:
        return s.n * SUBSTEPS * frames / dt, dt / frames, s.n, frames, max(warm_frames, 1)
    return s.n * SUBSTEPS * frames / dt, dt / frames, s.n


def reference_main(args, rank, world):
    if rank != 0:
        return 0
    # The reference's CPU solver (VtClothSolverCPU: single-threaded Gauss-Seidel, restated in oracle/ref_gs_cpu.c) on the SAME
    # workload as the velvet arm: the same 1024x1024 drape, W warm-up frames, K timed frames (~12 s each on one host core --
    # the algorithm is sequential, so one core is all it can use).  Under torchrun this is one of the N identical cloths.
    res = args.cpu_resolution if args.cpu_resolution is not None else args.resolution
    asked_steps, asked_warm = max(1, args.steps), max(0, args.warmup)
    budget = float(os.environ.get("VELVET_REF_BUDGET_S", "150"))
    value, sec_per_frame, n, steps, warm = run_cpu_reference(res, asked_steps, asked_warm, budget_s=budget)
    same = res == args.resolution
    cfg = workload_config(args.resolution, world)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": sec_per_frame * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg, "same_config": same,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": ((f"--steps {asked_steps} --warmup {asked_warm} bounded to " if (steps, warm) != (asked_steps, asked_warm) else "")
                                    + f"{steps} timed + {warm} warm-up frame(s) "
                                    + (f"(a frame takes {sec_per_frame:.1f} s and cannot be split; budget {budget:.0f} s, VELVET_REF_BUDGET_S) "
                                       if (steps, warm) != (asked_steps, asked_warm) else "") + "of "
                                    + ("the same workload (one cloth)" if same else f"a {res + 1}x{res + 1} sample of the workload")
                                    + "; VtClothSolverCPU restated (oracle/ref_gs_cpu.c), single thread like the reference "
                                      f"(Gauss-Seidel is sequential); host has {os.cpu_count()} cores")},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ---------------------------------------------------------------------------------------------- extra records
def timed_frames(torch, simulate, synchronize, stream, barrier, group, frames, warm):
    """W untimed + K timed frames bracketed by barrier + synchronize, CUDA events on the solver stream, max over ranks."""
    for _ in range(warm):
        simulate()
    synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(frames):
        simulate()
    e1.record(stream)
    synchronize()
    barrier()
    return group.max(e0.elapsed_time(e1)) / frames


def sub_record_batch64(torch, vb, group, rank, world, local_rank, barrier, instances=4096, frames=5, warm=2):
    """BASELINE configs[3]: `instances` independent 64x64 cloths in total, sharded over the ranks with no communication."""
    from velvet_b200.distributed import instance_model_height, shard_instances
    p = vb.default_params()
    p.numSubsteps, p.numIterations = SUBSTEPS, ITERATIONS
    R = 63
    mine = shard_instances(instances, world, rank)
    g = vb.VtClothSolverGPU(p, device=local_rank)
    v, idx = vb.GenerateClothMesh(R)
    g.AddClothInstances(R, v, idx, [vb.TransformMatrix((0, instance_model_height(k), 1.0), (90, 0, 0), (1, 1, 1)) for k in mine])
    g.UpdateColliders(vb.sphere_plane_colliders())
    stream = torch.cuda.ExternalStream(g.stream, device=torch.device("cuda", local_rank))
    ms = timed_frames(torch, lambda: g.Simulate(sync=False), g.Synchronize, stream, barrier, group, frames, warm)
    particles = group.sum(float(g.simParams.numParticles))
    kernel = "grid" if g.iterateKernel == vb.ITERATE_GRID else "tiles"
    finite = bool(__import__("numpy").isfinite(g.download("positions")).all())
    g.close()
    return {"config": f"{instances} independent 64x64 cloths ({int(particles)} particles) sharded over {world} GPU(s), no communication, "
                      f"self-collision + SDF sphere + plane, {SUBSTEPS} substeps x {ITERATIONS} iterations",
            "scaling": "strong", "n_gpus": world, "ms_per_step": ms, "value": particles * SUBSTEPS / (ms * 1e-3), "unit": UNIT,
            "steps": frames, "warmup": warm, "instances_this_rank": len(mine), "iterate_kernel": kernel, "result_finite": finite}


def sub_record_decomposed(torch, vb, group, rank, world, local_rank, barrier, resolution=4095, frames=5, warm=2):
    """BASELINE configs[4]: ONE cloth decomposed over the ranks (NVLink peer-memory halo exchange once per Jacobi iteration,
    cross-partition self-collision); at world == 1 the plain single-GPU solver of the same cloth (the strong-scaling base)."""
    import numpy as np
    p = vb.default_params()
    p.numSubsteps, p.numIterations = SUBSTEPS, ITERATIONS
    g = vb.build_scene(resolution, p, device=local_rank)
    g.UpdateColliders(vb.sphere_plane_colliders())
    n = g.simParams.numParticles
    stream = torch.cuda.ExternalStream(g.stream, device=torch.device("cuda", local_rank))
    rec = {"config": f"one {resolution + 1}x{resolution + 1} cloth ({n} particles) "
                     + (f"domain-decomposed over {world} GPUs, halo exchange once per Jacobi iteration, cross-partition self-collision"
                        if world > 1 else "on one GPU (strong-scaling base of the decomposed runs)")
                     + f", drape over SDF sphere + plane, {SUBSTEPS} substeps x {ITERATIONS} iterations",
           "scaling": "strong", "n_gpus": world, "unit": UNIT, "steps": frames, "warmup": warm}
    if world == 1:
        ms = timed_frames(torch, lambda: g.Simulate(sync=False), g.Synchronize, stream, barrier, group, frames, warm)
        rec.update({"transport": "none", "iterate_kernel": "grid" if g.iterateKernel == vb.ITERATE_GRID else "tiles",
                    "launches_per_step": g.lastLaunchCount})
    else:
        from velvet_b200.decomposed import DecomposedCloth
        dd = DecomposedCloth(g, local_rank, transport="peer")
        ms = timed_frames(torch, lambda: dd.Simulate(sync=False), g.Synchronize, stream, barrier, group, frames, warm)
        rec.update({"transport": dd.transport, "owned_particles_this_rank": int(dd.info.ownedCount),
                    "halo_bytes_per_iteration_this_rank": int(dd.halo_bytes_per_iteration),
                    "launches_per_step": int(g._L.velvet_solver_last_launch_count(g._h))})
    # every rank must hold the same full state: checksum of the positions, compared across ranks
    pos = g.download("positions")
    chk = float(np.sum(pos.astype(np.float64)))
    rec["positions_checksum"] = chk
    rec["ranks_agree"] = len(set(group.gather_all(chk))) == 1
    rec["result_finite"] = bool(np.isfinite(pos).all())
    rec["ms_per_step"] = ms
    rec["value"] = n * SUBSTEPS / (ms * 1e-3)
    if world > 1:
        dd.close()
    g.close()
    return rec


def run_ref_cuda(R, frames, warm):
    """O3: the reference's own VtClothSolverGPU.cu / SpatialHashGPU.cu, unmodified, compiled for sm_100a on stand-in headers
    (oracle/ref_cuda/build_ref_cuda.sh -> oracle/_ref), on the same scene and frames.  A reported baseline like cpu_baseline:
    test infrastructure, timed after every velvet measurement is finished."""
    from oracle import o1, refcuda
    if not refcuda.available():
        return {"unavailable": "oracle/_ref/libvelvet_refcuda.so not built (needs /root/reference at build time)"}
    p = o1.default_params()
    p.numSubsteps, p.numIterations = SUBSTEPS, ITERATIONS
    M = o1.transform_matrix((0, 1.5, 1.0), (90, 0, 0), (1, 1, 1))
    o = o1.O1Solver(p)
    v, idx = o1.generate_cloth_mesh(R)
    o.cloth_object_start(R, v, idx, M, [])
    r = refcuda.RefCudaSolver(p)
    r.register_like(o, R, M, [])
    r.set_colliders([o1.make_collider(o1.PLANE, (0, 0, 0), (1, 1, 1)), o1.make_collider(o1.SPHERE, (0, 0.6, 0), (0.6, 0.6, 0.6))])
    for _ in range(warm):
        r.simulate()
    total, t0 = 0.0, time.perf_counter()
    for _ in range(frames):
        r.simulate()
        total += r.timers().get("Solver_Total", 0.0)
    wall_ms = (time.perf_counter() - t0) * 1e3 / frames
    n = (R + 1) ** 2
    ms = total / frames  # its own cudaEvent pair around the Simulate() body
    return {"impl": "reference CUDA kernels (VtClothSolverGPU.cu + SpatialHashGPU.cu unmodified, nvcc sm_100a, oracle/_ref)",
            "ms_per_step": ms, "wall_ms_per_step": wall_ms, "value": n * SUBSTEPS / (ms * 1e-3), "unit": UNIT, "steps": frames,
            "warmup": warm, "how": "the reference's own Solver_Total event timer per Simulate(), same scene, same frame indices"}


# ---------------------------------------------------------------------------------------------- our arm
def velvet_main(args, rank, world, local_rank):
    import torch

    import velvet_b200 as vb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the velvet_b200 arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    from velvet_b200.distributed import Group, instance_model_height
    group = Group("nccl", device=torch.device("cuda", local_rank))

    def barrier():
        group.barrier()
        torch.cuda.synchronize()

    R = args.resolution
    p = vb.default_params()
    p.numSubsteps, p.numIterations = SUBSTEPS, ITERATIONS
    torch.empty(16, dtype=torch.uint8).pin_memory()  # torch's one-off pinned-allocator / context set-up stays out of setup_s
    t0 = time.perf_counter()
    math_mode = vb.MATH_FAST if args.math == "fast" else vb.MATH_EXACT
    batch = args.workload == "batch64"
    if batch:
        # BASELINE config 4: 4 096 independent 64x64 cloths in total, sharded over the ranks with no communication
        from velvet_b200.distributed import shard_instances
        R = 63
        mine = shard_instances(args.instances, world, rank)
        g = vb.VtClothSolverGPU(p, device=local_rank, tile_size=args.tile, math_mode=math_mode)
        v, idx = vb.GenerateClothMesh(R)
        g.AddClothInstances(R, v, idx, [vb.TransformMatrix((0, instance_model_height(k), 1.0), (90, 0, 0), (1, 1, 1)) for k in mine])
    else:
        g = vb.build_scene(R, p, position=(0, instance_model_height(rank), 1.0), rotation=(90, 0, 0), device=local_rank,
                           tile_size=args.tile, math_mode=math_mode)
    cols = vb.sphere_plane_colliders()
    raw = b"".join(bytes(c) for c in cols)
    pinned_cols = torch.empty(len(raw), dtype=torch.uint8).pin_memory()
    pinned_cols.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
    g.UpdateCollidersRaw(C.c_void_p(pinned_cols.data_ptr()), len(cols))
    g.Simulate()  # builds the tile plan, captures the graph
    setup_s = time.perf_counter() - t0
    N = g.simParams.numParticles
    ninst = len(mine) if batch else 1
    S = g.buffer_ptr("stretchLengths")[1] * ninst
    B = g.buffer_ptr("bendAngles")[1] * ninst
    A = g.buffer_ptr("attachDistances")[1] * ninst
    launches = g.lastLaunchCount
    iterate_kernel = "iterate_grid_kernel" if g.iterateKernel == vb.ITERATE_GRID else "iterate_tile_kernel"
    log(f"[rank {rank}] setup {setup_s:.2f}s  N={N} S={S} B={B} A={A}  launches/frame={launches}")

    stream = torch.cuda.ExternalStream(g.stream, device=torch.device("cuda", local_rank))
    W = max(args.warmup, 3)

    # The drape evolves (contacts and active constraints grow from frame to frame), so every timed region below replays
    # the SAME frames: the state right after registration is restored and W warm-up frames are run before each region.
    g.Synchronize()
    state0 = {k: g.download(k).copy() for k in ("positions", "velocities", "predicted")}

    def restart(warm):
        for k, v in state0.items():
            g.upload(k, v)
        for _ in range(warm):
            g.UpdateCollidersRaw(C.c_void_p(pinned_cols.data_ptr()), len(cols))
            g.Simulate(sync=False)
        g.Synchronize()

    restart(W)

    # ---- timed region 1: device-resident (value)
    sampler = ClockSampler(local_rank, torch)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    wall0 = time.perf_counter()
    ev0.record(stream)
    for _ in range(args.steps):
        g.Simulate(sync=False)
    ev1.record(stream)
    g.Synchronize()
    wall = time.perf_counter() - wall0
    barrier()
    clocks = sampler.stop()
    ms = group.max(ev0.elapsed_time(ev1))
    ms_per_step = ms / args.steps
    total_particles = group.sum(float(N))
    value = total_particles * SUBSTEPS * args.steps / (ms * 1e-3)

    # ---- timed region 2: end to end through the C ABI with host buffers.  Every frame uploads the collider block from
    # pinned host memory and reads positions + normals back into pinned host memory; the read-back is double-buffered
    # (velvet_solver_readback_pipelined: copy stream + two host buffers), so frame k travels over PCIe while frame k+1 is
    # simulated, and the host consumes frame k-1's result while frame k runs.  Every frame's result is read inside the
    # timed region (the last one is waited for before the closing event is recorded).
    host_pos = [torch.empty(N * 3, dtype=torch.float32).pin_memory() for _ in range(2)]
    host_nrm = [torch.empty(N * 3, dtype=torch.float32).pin_memory() for _ in range(2)]
    h2d = len(raw)
    d2h = 2 * N * 12
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for slot in range(2):  # warm-up of the read-back path (allocates the two device staging buffers)
        g.ReadbackWait(g.ReadbackPipelined(C.c_void_p(host_pos[slot].data_ptr()), C.c_void_p(host_nrm[slot].data_ptr())))
    restart(W)
    barrier()
    e0.record(stream)
    checksum = 0.0
    prev = None
    for k in range(args.steps):
        g.UpdateCollidersRaw(C.c_void_p(pinned_cols.data_ptr()), len(cols))  # H2D from pinned host memory
        g.Simulate(sync=False)
        ticket = g.ReadbackPipelined(C.c_void_p(host_pos[k & 1].data_ptr()), C.c_void_p(host_nrm[k & 1].data_ptr()))
        if prev is not None:
            g.ReadbackWait(prev[0])
            checksum += float(host_pos[prev[1]][1])  # the host consumes the previous step's result
        prev = (ticket, k & 1)
    g.ReadbackWait(prev[0])
    checksum += float(host_pos[prev[1]][1])
    e1.record(stream)
    g.Synchronize()
    barrier()
    e2e_ms = group.max(e0.elapsed_time(e1))
    e2e_value = total_particles * SUBSTEPS * args.steps / (e2e_ms * 1e-3)
    finite = bool(torch.isfinite(host_pos[prev[1]]).all())

    # the serial form of the same loop (simulate, read back, wait, repeat), for reference
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    serial_steps = min(args.steps, 10)
    restart(W)
    s0.record(stream)
    for _ in range(serial_steps):
        g.UpdateCollidersRaw(C.c_void_p(pinned_cols.data_ptr()), len(cols))
        g.Simulate(sync=False)
        g.ReadbackAsync(C.c_void_p(host_pos[0].data_ptr()), C.c_void_p(host_nrm[0].data_ptr()))
        g.Synchronize()
        checksum += float(host_pos[0][1])
    s1.record(stream)
    g.Synchronize()
    e2e_serial_ms = s0.elapsed_time(s1) / serial_steps

    # ---- the two multi-GPU configurations of BASELINE.json as sub-records (collective: every rank takes part)
    sub_records = None
    if not args.no_sub_records and not batch:
        sub_records = {}
        for name, fn in (("batch64_4096_cloths", sub_record_batch64), ("decomposed_4096x4096", sub_record_decomposed)):
            try:
                log(f"[rank {rank}] sub-record {name} ...")
                sub_records[name] = fn(torch, vb, group, rank, world, local_rank, barrier)
            except Exception as e:
                sub_records[name] = {"failed": f"{type(e).__name__}: {e}"}
                log(f"[rank {rank}] sub-record {name} failed: {e}")
            barrier()

    if rank != 0:
        group.barrier()  # rank 0 is still measuring stages / the CPU baseline
        group.close()
        return 0

    # ---- roofline of the dominant kernel (tile-fused Jacobi iteration), CUDA events around each stage
    stage_frames = 3
    stages = {}
    restart(W + max(0, args.steps // 2 - 1))  # the middle frames of the timed region
    for _ in range(stage_frames):
        for k, v in g.SimulateTimed().items():
            stages[k] = stages.get(k, 0.0) + v / stage_frames
    nb = g.download("neighbors")[: 64 * N].reshape(64, N)
    nbar = float(((nb != 0xFFFFFFFF).cumprod(0)).sum(0).mean())
    del nb
    rebuilds = len([s for s in range(SUBSTEPS) if s % p.interleavedHash == 0])
    alg = algorithmic_bytes(N, S, B, A, rebuilds, nbar)
    # the grid kernel runs all iterations of a substep in ONE launch (grid-wide barriers in between): a launch is then
    # `iters_per_launch` iterations, with that many times the algorithmic bytes of one iteration
    frame_launches = int(g.lastLaunchCount)
    iters_per_launch = ITERATIONS if (iterate_kernel.startswith("iterate_grid") and frame_launches < SUBSTEPS * ITERATIONS) else 1
    iteration_ms = stages.get("Solver_Iterate", 0.0) / (SUBSTEPS * ITERATIONS)
    iter_launch_ms = iteration_ms * iters_per_launch
    peak, peak_src = measured_peak_gbs()
    achieved = alg["iterate_per_launch"] * iters_per_launch / (iter_launch_ms * 1e-3) / 1e9 if iter_launch_ms > 0 else 0.0
    traffic, traffic_src = None, None
    try:  # static: one `ncu --set full` capture of this kernel on this workload, committed with its summary under profiles/
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        traffic = tj.get(iterate_kernel + "_dram_bytes_per_launch")
        traffic_src = tj.get("source")
    except Exception:
        pass
    frame_gbs = alg["frame"] / (ms_per_step * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": iterate_kernel + " (SolveStretch+SolveAttach+SolveBending+ApplyDeltas fused)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": ("static, not measured in this run: " + str(traffic_src)) if traffic is not None else None,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg["iterate_per_launch"] * iters_per_launch,
                "algorithmic_bytes_per_particle": alg["per_particle_iter"], "launch_ms": iter_launch_ms,
                "iterations_per_launch": iters_per_launch, "iteration_ms": iteration_ms,
                "how": f"CUDA events per stage on the solver stream, un-graphed pass over the middle frames of the timed region, "
                       f"mean of {SUBSTEPS * ITERATIONS // iters_per_launch} launches x {stage_frames} frames",
                "share_of_frame": stages.get("Solver_Iterate", 0.0) / max(stages.get("Solver_Total", 1e-9), 1e-9),
                "frame": {"algorithmic_bytes": alg["frame"], "achieved": frame_gbs, "frac": frame_gbs / peak}}

    # ---- the other math mode, same solver, same timing method (reported, not the headline)
    other = {}
    other_mode, other_name = (vb.MATH_EXACT, "exact") if args.math == "fast" else (vb.MATH_FAST, "fast")
    g.SetMathMode(other_mode)
    restart(W)
    o0, o1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    o0.record(stream)
    for _ in range(args.steps):
        g.Simulate(sync=False)
    o1_.record(stream)
    g.Synchronize()
    oms = o0.elapsed_time(o1_) / args.steps
    restart(W + max(0, args.steps // 2 - 1))
    ostage = g.SimulateTimed()
    other = {"math": other_name, "ms_per_step": oms, "value": N * SUBSTEPS / (oms * 1e-3),
             "iteration_ms": ostage.get("Solver_Iterate", 0.0) / (SUBSTEPS * ITERATIONS)}
    g.SetMathMode(math_mode)

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline and not batch:
        log("timing the CPU reference (VtClothSolverCPU restated, 1 thread) on 1 frame of the same workload ...")
        v, spf, n = run_cpu_reference(R, 1)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                        "sample": f"1 frame of the same {R + 1}x{R + 1} workload ({spf:.1f} s); VtClothSolverCPU restated in "
                                  f"oracle/ref_gs_cpu.c, single thread like the reference; host has {os.cpu_count()} cores"}

    ref_cuda = None
    if world == 1 and not args.no_sub_records and not batch:
        log("timing the reference's own CUDA kernels (oracle/_ref) on the same frames ...")
        try:
            ref_cuda = run_ref_cuda(R, args.steps, W)
            if "value" in ref_cuda:
                ref_cuda["velvet_over_ref_cuda"] = value / ref_cuda["value"]
        except Exception as e:  # a baseline must never take the product line down
            ref_cuda = {"unavailable": f"{type(e).__name__}: {e}"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if batch else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": ({"workload": f"{args.instances} independent {R + 1}x{R + 1} cloths ({int(total_particles)} particles in total) batched in one "
                               f"solver per GPU, sharded over {world} GPU(s) with no communication, self-collision + SDF sphere + plane, "
                               f"{SUBSTEPS} substeps x {ITERATIONS} iterations", "particles_per_gpu": N, "stretch": S, "bend": B, "attach": A,
                    "substeps": SUBSTEPS, "iterations": ITERATIONS} if batch else workload_config(R, world, p.interleavedHash)),
        "implementation": {"pipeline": "fused", "iterate_kernel": iterate_kernel, "mean_neighbors": nbar, "tile": args.tile or 256,
                           "math": args.math + (" (bit-identical to the CPU oracle)" if args.math == "exact" else " (approximate div/sqrt)"),
                           "l2": "per-frame working set (SoA state 64 MB + rest lengths 17 MB + neighbor table ~60 MB + hash / "
                                 "packed-float3 buffers ~100 MB) exceeds the 126 MB L2; no flush between frames"},
        "ms_per_frame": ms_per_step, "wall_ms_per_step": wall * 1e3 / args.steps,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / args.steps, "result_finite": finite,
                "how": "double-buffered read-back on a copy stream (frame k over PCIe while frame k+1 is simulated)",
                "serial_ms_per_step": e2e_serial_ms},
        "gpu_launches": launches * args.steps, "gpu_launches_per_step": launches,
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline, "ref_cuda": ref_cuda, "sub_records": sub_records,
        "stages_ms": {k: round(v, 4) for k, v in stages.items()}, "setup_s": setup_s,
        "other_math_mode": other,
    }
    emit(line)
    group.barrier()
    group.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="velvet", choices=["velvet", "reference"])
    ap.add_argument("--resolution", type=int, default=1023, help="cloth resolution R (particles = (R+1)^2)")
    ap.add_argument("--cpu-resolution", type=int, default=None,
                    help="--impl reference: resolution of a smaller CPU sample (default: the same cloth as --resolution)")
    ap.add_argument("--no-sub-records", action="store_true", help="skip the configs[3] / configs[4] sub-records and the ref_cuda block")
    ap.add_argument("--tile", type=int, default=0, help="particles per Jacobi tile (0 = default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="drape1m", choices=["drape1m", "batch64"],
                    help="drape1m: BASELINE configs[2], the headline (default); batch64: configs[3], batched independent 64x64 cloths")
    ap.add_argument("--instances", type=int, default=4096, help="batch64: total number of cloths over all ranks")
    ap.add_argument("--math", default="exact", choices=["exact", "fast"],
                    help="float kernels: exact (library default, bit-identical to the oracle) or fast (opt-in)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.impl == "reference":
        return reference_main(args, rank, world)
    return velvet_main(args, rank, world, local_rank)


if __name__ == "__main__":
    sys.exit(main())
