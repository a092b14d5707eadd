#!/usr/bin/env python
"""bench.py -- headline benchmark of the XPBD cloth hot path (BASELINE.json: particle-substeps/sec, 10 iterations,
self-collision, 1M particles; ms/frame).

    python bench.py --gpus N --steps K --warmup W             # this repo's sm_100a solver
    python bench.py --impl reference --gpus N --steps K ...    # the reference's own CPU solver (port, oracle/)

One "step" = one Simulate() frame (1/60 s: 5 substeps x 10 Jacobi iterations, hash rebuilt every 3rd substep)
of BASELINE.json configs[2]: a 1024x1024 cloth (1,048,576 particles) draped over an SDF sphere + plane with
particle self-collision.  Under torchrun (N > 1) every rank simulates its own independent cloth on its own
GPU -- the batched-independent-instances mode: no data-path collective, weak scaling.

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


# ---------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
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
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(smax), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- CPU reference arm
def run_cpu_reference(resolution, frames, warm_frames=0):
    """Times O2 = the reference's CPU solver (VtClothSolverCPU, single-thread Gauss-Seidel) restated in oracle/ on this
    box's host cores.  Returns (particle_substeps_per_s, seconds_per_frame, n)."""
    from oracle import o1, o2
    p = o1.default_params()
    p.numSubsteps, p.numIterations = SUBSTEPS, ITERATIONS
    s = o2.O2Solver(p, resolution, o1.transform_matrix((0, 1.5, 1.0), (90, 0, 0), (1, 1, 1)))
    s.set_colliders([1, 0], [[0, 0, 0], [0, 0.6, 0]], [1.0, 0.6])
    for _ in range(warm_frames):
        s.simulate()
    t0 = time.perf_counter()
    for _ in range(frames):
        s.simulate()
    dt = time.perf_counter() - t0
    return s.n * SUBSTEPS * frames / dt, dt / frames, s.n


def reference_main(args, rank, world):
    if rank != 0:
        return 0
    # Each step is one frame of the same scene on a bounded sample: the single-threaded CPU solver needs ~25 s per
    # 1M-particle frame, so the default sample is a 256x256 cloth (65,536 particles, ~1.4 s per frame).
    res = args.cpu_resolution
    steps = max(1, min(args.steps, 5))
    warm = 1 if args.warmup > 0 else 0
    value, sec_per_frame, n = run_cpu_reference(res, steps, warm)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": sec_per_frame * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{res + 1}x{res + 1} cloth ({n} particles) self-colliding drape over SDF sphere + plane, "
                               f"{SUBSTEPS} substeps x {ITERATIONS} iterations (bounded sample of the 1024x1024 headline workload)",
                   "substeps": SUBSTEPS, "iterations": ITERATIONS},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": f"{steps} frame(s) of a {res + 1}x{res + 1} cloth; VtClothSolverCPU restated "
                                   f"(oracle/ref_gs_cpu.c), single thread like the reference (Gauss-Seidel is sequential)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


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
    sampler = ClockSampler(local_rank)
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
    iter_launch_ms = stages.get("Solver_Iterate", 0.0) / (SUBSTEPS * ITERATIONS)
    peak, peak_src = measured_peak_gbs()
    achieved = alg["iterate_per_launch"] / (iter_launch_ms * 1e-3) / 1e9 if iter_launch_ms > 0 else 0.0
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get("iterate_tile_kernel_dram_bytes_per_launch")
    except Exception:
        pass
    frame_gbs = alg["frame"] / (ms_per_step * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "iterate_tile_kernel (SolveStretch+SolveAttach+SolveBending+ApplyDeltas fused)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg["iterate_per_launch"],
                "algorithmic_bytes_per_particle": alg["per_particle_iter"], "launch_ms": iter_launch_ms,
                "how": f"CUDA events per stage on the solver stream, un-graphed pass over the middle frames of the timed region, "
                       f"mean of {SUBSTEPS * ITERATIONS} launches x {stage_frames} frames",
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
             "iterate_launch_ms": ostage.get("Solver_Iterate", 0.0) / (SUBSTEPS * ITERATIONS)}
    g.SetMathMode(math_mode)

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline and not batch:
        log("timing the CPU reference (VtClothSolverCPU restated, 1 thread) on 1 frame of the same workload ...")
        v, spf, n = run_cpu_reference(R, 1)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                        "sample": f"1 frame of the same {R + 1}x{R + 1} workload ({spf:.1f} s); VtClothSolverCPU restated in "
                                  f"oracle/ref_gs_cpu.c, single thread like the reference; host has {os.cpu_count()} cores"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if batch else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": (f"{args.instances} independent {R + 1}x{R + 1} cloths ({int(total_particles)} particles in total) batched in one "
                                f"solver per GPU, sharded over {world} GPU(s) with no communication, self-collision + SDF sphere + plane, "
                                f"{SUBSTEPS} substeps x {ITERATIONS} iterations" if batch else
                                f"{R + 1}x{R + 1} cloth ({N} particles) self-colliding drape over SDF sphere + plane, "
                                f"{SUBSTEPS} substeps x {ITERATIONS} iterations, hash every {p.interleavedHash} substeps"
                                + (f"; {world} independent cloths, one per GPU, no communication" if world > 1 else "")),
                   "particles_per_gpu": N, "stretch": S, "bend": B, "attach": A, "substeps": SUBSTEPS,
                   "iterations": ITERATIONS, "mean_neighbors": nbar, "pipeline": "fused", "tile": args.tile or 256,
                   "math": args.math + (" (bit-identical to the CPU oracle)" if args.math == "exact" else " (FMA + approximate div/sqrt)"),
                   "l2": "per-frame working set (SoA state 64 MB + constraints ~50 MB + neighbor table ~60 MB + hash / "
                         "packed-float3 buffers ~100 MB) exceeds the 126 MB L2; no flush between frames"},
        "ms_per_frame": ms_per_step, "wall_ms_per_step": wall * 1e3 / args.steps,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / args.steps, "result_finite": finite,
                "how": "double-buffered read-back on a copy stream (frame k over PCIe while frame k+1 is simulated)",
                "serial_ms_per_step": e2e_serial_ms},
        "gpu_launches": launches * args.steps, "gpu_launches_per_step": launches,
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
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
    ap.add_argument("--cpu-resolution", type=int, default=255, help="--impl reference: resolution of the bounded CPU sample")
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
