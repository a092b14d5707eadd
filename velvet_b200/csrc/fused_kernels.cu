// fused_kernels.cu -- the B200-native substep pipeline behind VtClothSolverGPU::Simulate.
//
// State lives in SoA float4 arrays (x, y, z, invMass) so that one 16-byte load brings a particle's position
// and its weight; the packed-float3 public buffers are touched once per frame (import / export).  Per frame:
//
//   begin_frame      import + CollideSDF pre-stabilisation (frame dt) + PredictPositions(substep 0)
//   per substep      [hash rebuild]  collide (particles + apply + SDF)  I x iterate  end_substep
//   normals          per-vertex gather
//
// All kernels are HBM/L2-bound gather-scatter work: no tensor cores.  Design points:
//   * every buffer that a kernel both gathers from and writes is double-buffered (predIn -> predOut), so no
//     kernel needs atomics or a separate delta array;
//   * the Jacobi iteration (reference: 3 constraint kernels with 6-16 global atomics per constraint + an
//     averaging kernel, VtClothSolverGPU.cu L65-264) is ONE kernel per iteration: a CTA owns a tile of
//     particles, stages tile + halo positions in shared memory, evaluates each constraint once, drops the
//     per-endpoint corrections into private shared-memory slots and lets each particle sum its slots in
//     constraint-id order -- deterministic, atomic-free, and ~0.9x the algorithmic bytes thanks to 16-bit
//     tile-local indices;
//   * colliders are prepared once per frame (lastTransform * invCurTransform hoisted) and staged per CTA.
#include "fused_kernels.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <utility>

#include "hash_kernels.cuh"
#include "vt_buffer.hpp"

#ifndef VT_FAST_MATH
#define VT_FAST_MATH 0
#endif
#if VT_FAST_MATH
#define VT_MATH_NS fast_math
#else
#define VT_MATH_NS exact_math
#endif

namespace velvet {
namespace VT_MATH_NS {

namespace {

constexpr int PB = 256;  // threads per CTA for per-particle kernels
inline unsigned pgrid(unsigned n) { return (n + PB - 1) / PB; }

// Launch with programmatic stream serialization (see vt_math.cuh: vt_pdl_wait): inside the frame's CUDA graph the edge to
// the previous kernel becomes a programmatic dependency, which hides most of the ~4 us a full kernel-to-kernel dependency
// costs (81 nodes per frame: it was two thirds of a 65k-particle frame).  VELVET_PDL=0 launches the ordinary way.
inline bool pdl_enabled()
{
    static const bool on = [] {
        const char* e = getenv("VELVET_PDL");
        return !(e && e[0] == '0');
    }();
    return on;
}
template <class... KArgs, class... Args>
void launch_kernel_ex(bool cooperative, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args);
// cooperative: the grid synchronises inside the kernel (grid-wide barriers), so the driver must place ALL its CTAs at once --
// with a plain launch two such grids on one device (two solvers on their own streams) could each hold part of the SMs and
// wait for the other for ever.  A cooperative launch is not combined with programmatic dependent launch.
template <class... KArgs, class... Args>
void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args)
{
    launch_kernel_ex(false, kernel, grid, block, smem, stream, std::forward<Args>(args)...);
}
template <class... KArgs, class... Args>
void launch_kernel_ex(bool cooperative, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    if (cooperative) {
        attr[0].id = cudaLaunchAttributeCooperative;
        attr[0].val.cooperative = 1;
        cfg.numAttrs = 1;
    } else {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.numAttrs = pdl_enabled() ? 1 : 0;
    }
    cfg.attrs = attr;
    VT_CUDA(cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...));
}

__device__ __forceinline__ float4 F4(vec3 v, float w) { return make_float4(v.x, v.y, v.z, w); }

__device__ __forceinline__ void stage_colliders(PreparedCollider* s_col, const PreparedCollider* __restrict__ g_col,
                                                unsigned n)
{
    // 196-byte structs copied as 49 words each (callers clamp n to the VT_MAX_COLLIDERS entries of the stage)
    const unsigned words = n * (unsigned)(sizeof(PreparedCollider) / 4);
    const unsigned* src = reinterpret_cast<const unsigned*>(g_col);
    unsigned* dst = reinterpret_cast<unsigned*>(s_col);
    for (unsigned i = threadIdx.x; i < words; i += blockDim.x) dst[i] = __ldg(src + i);
    __syncthreads();
}

__global__ void prepare_inputs_kernel(const VtSDFCollider* __restrict__ colliders, PreparedCollider* prepared,
                                      const float* __restrict__ slotPositions, float* __restrict__ slotPositionsOut,
                                      unsigned numSlotFloats, const FrameParams* __restrict__ fp)
{
    vt_pdl_trigger();
    vt_pdl_wait();
    const unsigned i = threadIdx.x;
    if (i < min(fp->numColliders, VT_MAX_COLLIDERS)) prepare_collider(colliders[i], prepared[i]);
    for (unsigned k = i; k < numSlotFloats; k += blockDim.x) slotPositionsOut[k] = slotPositions[k];
}

__global__ void __launch_bounds__(PB) begin_frame_kernel(const float* __restrict__ positions,
                                                         const float* __restrict__ velocities,
                                                         const float* __restrict__ invMasses, float4* __restrict__ pos4,
                                                         float4* __restrict__ pred,
                                                         const PreparedCollider* __restrict__ colliders,
                                                         const FrameParams* __restrict__ fp, unsigned n)
{
    vt_pdl_trigger();
    vt_pdl_wait();
    __shared__ PreparedCollider s_col[VT_MAX_COLLIDERS];
    const unsigned nc = min(fp->numColliders, VT_MAX_COLLIDERS);
    stage_colliders(s_col, colliders, nc);
    const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n) return;
    const VtSimParams& P = fp->P;
    vec3 pos = load3(positions, id);
    const float w = invMasses[id];
    // pre-stabilisation: CollideSDF(positions, colliders, positions, frameTime), VtClothSolverGPU.hpp L73
    pos = collide_sdf_point(s_col, nc, pos, pos, P.collisionMargin, P.friction, fp->frameTime);
    // PredictPositions, VtClothSolverGPU.cu L51-52
    const float dt = fp->substepTime;
    const vec3 vel = load3(velocities, id) + V3(P.gravity[0], P.gravity[1], P.gravity[2]) * dt;
    pos4[id] = F4(pos, w);
    pred[id] = F4(pos + vel * dt, w);
}

__global__ void __launch_bounds__(PB, 5) collide_kernel(const float4* __restrict__ predIn, float4* __restrict__ predOut,
                                                     const float4* __restrict__ pos4,
                                                     const unsigned* __restrict__ neighbors,
                                                     const PreparedCollider* __restrict__ colliders,
                                                     const FrameParams* __restrict__ fp, unsigned N, int selfCollision,
                                                     const unsigned* __restrict__ subset, unsigned subsetCount,
                                                     unsigned rangeBegin)
{
    vt_pdl_trigger();
    vt_pdl_wait();
    __shared__ PreparedCollider s_col[VT_MAX_COLLIDERS];
    const unsigned nc = min(fp->numColliders, VT_MAX_COLLIDERS);
    stage_colliders(s_col, colliders, nc);
    // domain-decomposed mode: only the particles this rank owns -- a list (subset != nullptr) or the contiguous range
    // [rangeBegin, rangeBegin + subsetCount) of a strip
    const unsigned tidx = blockIdx.x * blockDim.x + threadIdx.x;
    if (tidx >= ((subset || subsetCount) ? subsetCount : N)) return;
    const unsigned id = subset ? __ldg(subset + tidx) : rangeBegin + tidx;
    const VtSimParams& P = fp->P;
    const float4 pi4 = predIn[id];
    const float4 xi4 = pos4[id];
    vec3 pred_i = V3(pi4);
    const vec3 pos_i = V3(xi4);
    const float w_i = pi4.w;

    if (selfCollision) {
        // CollideParticles_Kernel, VtClothSolverGPU.cu L339-372 (gather over the cached neighbour column)
        vec3 positionDelta = V3(0, 0, 0);
        int deltaCount = 0;
        const vec3 vel_i = pred_i - pos_i;
        const float D = P.particleDiameter;
        const float farEnough2 = D * D * 1.00001f;
        const unsigned maxK = (unsigned)P.maxNumNeighbors;
        // The column walk is a chain of dependent loads (id -> predicted[id]); four ids and four gathers are kept in
        // flight per trip.  Contributions are still accumulated in list order.
        const unsigned* col = neighbors + id;
        bool done = false;
        for (unsigned k0 = 0; k0 < maxK && !done; k0 += 4, col += 4 * (size_t)N) {
            unsigned j[4];
#pragma unroll
            for (int i = 0; i < 4; i++) j[i] = (k0 + i < maxK) ? __ldg(col + (size_t)i * N) : 0xffffffffu;
            float4 pj[4];
#pragma unroll
            for (int i = 0; i < 4; i++) pj[i] = (j[i] <= N) ? __ldg(predIn + (j[i] < N ? j[i] : 0)) : make_float4(0, 0, 0, 0);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                if (done) break;
                if (j[i] > N) {  // terminator (or end of the table): the reference's `if (j > numParticles) break`
                    done = true;
                    break;
                }
                const float4 pj4 = pj[i];
                const float denom = w_i + pj4.w;
                if (denom <= 0) continue;
                const vec3 pred_j = V3(pj4);
                const vec3 diff = pred_i - pred_j;
                // `distance >= D` is certain when the squared distance exceeds D^2 by more than any rounding can undo
                // (sqrt is monotone and correctly rounded): nearly every listed neighbour of a cloth that is not folded
                // leaves here, without the IEEE square root (22 % of this kernel's instructions before)
                const float distance2 = dot(diff, diff);
                if (distance2 > farEnough2) continue;
                const float distance = sqrtf(distance2);
                if (distance >= D) continue;
                const vec3 gradient = diff / (distance + VT_EPSILON);
                const float lambda = vt_div(distance - D, denom);
                const vec3 common = lambda * gradient;
                deltaCount++;
                positionDelta -= w_i * common;
                const vec3 relativeVelocity = vel_i - (pred_j - V3(__ldg(pos4 + j[i])));
                positionDelta += w_i * compute_friction(P.friction, common, relativeVelocity);
            }
        }
        // ApplyDeltas_Kernel, L257-263
        const float count = (float)deltaCount;
        if (count > 0) pred_i += positionDelta / count * P.relaxationFactor;
    }
    // CollideSDF_Kernel with the substep dt, L298-313
    pred_i = collide_sdf_point(s_col, nc, pred_i, pos_i, P.collisionMargin, P.friction, fp->substepTime);
    predOut[id] = F4(pred_i, w_i);
}

// ---------------------------------------------------------------- tile-fused Jacobi iteration
//
// Shared memory: sp[maxLocals] float4            tile + halo predicted positions, w = invMass
//                slots[maxK][tileSize] float4    xyz = correction, w = 1 if the constraint was active
//                s_brec[maxBendPerTile] uint4, s_srec[maxStretchPerTile] uint2   the tile's constraint records
// Slot (ordinal k, particle l) lives at slots[k * tileSize + l]: the per-particle sums read consecutive 16-byte words
// (no bank conflicts); constraint threads scatter, but consecutive constraints touch neighbouring particles.
// Every global load of a tile is issued before the first barrier so that the latencies overlap; the constraint records go
// global -> shared with cp.async and never occupy registers (prefetching them into registers spilled under the 64-register
// cap, and a spill store has to wait for its load: ncu showed the long-scoreboard stalls on exactly those STL).

// slot index of an encoded endpoint: row = ordinal (low 5 bits), column = local particle index
template <int LOG2T>
__device__ __forceinline__ unsigned slot_of(unsigned e)
{
    return ((e & 31u) << LOG2T) + ((e >> TP_ORD_BITS) & ((1u << LOG2T) - 1u));  // halo columns wrap inside the dump row
}

__device__ __forceinline__ void cp_async_16(void* smemDst, const void* gmemSrc)
{
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smemDst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gmemSrc));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// One constraint: positions from the staged tile, corrections into the endpoints' private slots.  An inactive constraint
// stores +0 vectors with flag 0, so the per-particle sums need no select.
// EXACT build: the checked-fast evaluators of vt_math.cuh; when their validity predicate fails (rare: pinned pair,
// degenerate or non-finite geometry) the constraint is redone by an out-of-line copy of the branchy evaluator, so the hot
// loop stays small and straight.  FAST build: the predicate is constant true.
struct StretchOut {
    vec3 c1, c2;
    float flag;
};
// first half: positions, distance, active flag; second half (only for active constraints): the divisions
struct StretchWork {
    float4 pa, pb;
#if !VT_FAST_MATH
    StretchHalf h;
#endif
    bool active;
};
__device__ __forceinline__ StretchWork stretch_begin(const uint2 r, const float4* __restrict__ sp, bool& ok)
{
    const unsigned ea = r.x & 0xffffu, eb = r.x >> 16;
    StretchWork w;
    w.pa = sp[ea >> TP_ORD_BITS];
    w.pb = sp[eb >> TP_ORD_BITS];
#if VT_FAST_MATH
    w.active = true;
#else
    w.h = stretch_begin_u(V3(w.pa), V3(w.pb), w.pa.w, w.pb.w, __uint_as_float(r.y), ok);
    w.active = w.h.active;
#endif
    return w;
}
__device__ __forceinline__ StretchOut stretch_finish(const uint2 r, const StretchWork& w, bool& ok)
{
    StretchOut o;
#if VT_FAST_MATH
    const bool active = stretch_eval_flagged(V3(w.pa), V3(w.pb), w.pa.w, w.pb.w, __uint_as_float(r.y), o.c1, o.c2);
    if (!active) o.c1 = o.c2 = V3(0, 0, 0);
    o.flag = active ? 1.0f : 0.0f;
#else
    stretch_finish_u(w.h, w.pa.w, w.pb.w, __uint_as_float(r.y), o.c1, o.c2, ok);
    o.flag = 1.0f;
#endif
    return o;
}
__device__ __forceinline__ StretchOut stretch_inactive()
{
    StretchOut o;
    o.c1 = o.c2 = V3(0, 0, 0);
    o.flag = 0.0f;
    return o;
}
template <int LOG2T>
__device__ __forceinline__ void stretch_store(const uint2 r, const StretchOut& o, float4* __restrict__ slots, unsigned dumpRow)
{
    const unsigned ea = r.x & 0xffffu, eb = r.x >> 16;
    // halo endpoints carry the dump-row ordinal and are simply not stored (a store into a shared dump row would be a
    // benign write-write race, but it would drown compute-sanitizer racecheck in false positives)
    if ((ea & 31u) != dumpRow) slots[slot_of<LOG2T>(ea)] = F4(o.c1, o.flag);
    if ((eb & 31u) != dumpRow) slots[slot_of<LOG2T>(eb)] = F4(o.c2, o.flag);
}

struct BendOut {
    vec3 c0, c1, c2, c3;
    float flag;
};
template <int LOG2T>
__device__ __forceinline__ void bend_store(const uint4 r, const BendOut& o, float4* __restrict__ slots, unsigned dumpRow)
{
    const unsigned e0 = r.x & 0xffffu, e1 = r.x >> 16, e2 = r.y & 0xffffu, e3 = r.y >> 16;
    if ((e0 & 31u) != dumpRow) slots[slot_of<LOG2T>(e0)] = F4(o.c0, o.flag);
    if ((e1 & 31u) != dumpRow) slots[slot_of<LOG2T>(e1)] = F4(o.c1, o.flag);
    if ((e2 & 31u) != dumpRow) slots[slot_of<LOG2T>(e2)] = F4(o.c2, o.flag);
    if ((e3 & 31u) != dumpRow) slots[slot_of<LOG2T>(e3)] = F4(o.c3, o.flag);
}

#if !VT_FAST_MATH
template <int LOG2T>
__device__ __noinline__ void stretch_slow_to_slots(const uint2 r, const float4* __restrict__ sp, float4* __restrict__ slots, unsigned dumpRow)
{
    const unsigned ea = r.x & 0xffffu, eb = r.x >> 16;
    const float4 pa = sp[ea >> TP_ORD_BITS], pb = sp[eb >> TP_ORD_BITS];
    StretchOut o;
    o.c1 = o.c2 = V3(0, 0, 0);
    o.flag = stretch_eval(V3(pa), V3(pb), pa.w, pb.w, __uint_as_float(r.y), o.c1, o.c2) ? 1.0f : 0.0f;
    stretch_store<LOG2T>(r, o, slots, dumpRow);
}
template <int LOG2T>
__device__ __noinline__ void bend_slow_to_slots(const uint4 r, const float4* __restrict__ sp, float4* __restrict__ slots,
                                                float xpbd_bend, unsigned dumpRow)
{
    const unsigned e0 = r.x & 0xffffu, e1 = r.x >> 16, e2 = r.y & 0xffffu, e3 = r.y >> 16;
    const float4 p0 = sp[e0 >> TP_ORD_BITS], p1 = sp[e1 >> TP_ORD_BITS], p2 = sp[e2 >> TP_ORD_BITS], p3 = sp[e3 >> TP_ORD_BITS];
    BendOut o;
    o.c0 = o.c1 = o.c2 = o.c3 = V3(0, 0, 0);
    o.flag = bend_eval(V3(p0), V3(p1), V3(p2), V3(p3), p0.w, p1.w, p2.w, p3.w, __uint_as_float(r.z), xpbd_bend, o.c0, o.c1,
                       o.c2, o.c3) ? 1.0f : 0.0f;
    bend_store<LOG2T>(r, o, slots, dumpRow);
}
#endif

template <int LOG2T>
__device__ __forceinline__ void bend_to_slots(const uint4 r, const float4* __restrict__ sp, float4* __restrict__ slots,
                                              float xpbd_bend, unsigned dumpRow)
{
    const unsigned e0 = r.x & 0xffffu, e1 = r.x >> 16, e2 = r.y & 0xffffu, e3 = r.y >> 16;
    const float4 p0 = sp[e0 >> TP_ORD_BITS], p1 = sp[e1 >> TP_ORD_BITS], p2 = sp[e2 >> TP_ORD_BITS], p3 = sp[e3 >> TP_ORD_BITS];
    BendOut o;
#if VT_FAST_MATH
    o.c0 = o.c1 = o.c2 = o.c3 = V3(0, 0, 0);
    const bool active = bend_eval(V3(p0), V3(p1), V3(p2), V3(p3), p0.w, p1.w, p2.w, p3.w, __uint_as_float(r.z), xpbd_bend,
                                  o.c0, o.c1, o.c2, o.c3);
#else
    bool ok = true;
    const bool active = bend_eval_u(V3(p0), V3(p1), V3(p2), V3(p3), p0.w, p1.w, p2.w, p3.w, __uint_as_float(r.z), xpbd_bend,
                                    o.c0, o.c1, o.c2, o.c3, ok);
    if (!ok) {
        bend_slow_to_slots<LOG2T>(r, sp, slots, xpbd_bend, dumpRow);
        return;
    }
    if (!active) o.c0 = o.c1 = o.c2 = o.c3 = V3(0, 0, 0);
#endif
    o.flag = active ? 1.0f : 0.0f;
    bend_store<LOG2T>(r, o, slots, dumpRow);
}

// Sum of a particle's slots in ordinal (= constraint id) order.  Inactive slots hold +0 vectors with flag 0: adding +0 is
// bit-neutral because the accumulator starts at +0 and can never become -0.
template <int LOG2T>
__device__ __forceinline__ void sum_slots(const float4* __restrict__ slots, unsigned tid, unsigned n, vec3& delta, float& count)
{
    for (unsigned k = 0; k < n; k++) {
        const float4 v = slots[(k << LOG2T) + tid];
        delta.x += v.x;
        delta.y += v.y;
        delta.z += v.z;
        count += v.w;
    }
}

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- bulk copies (TMA, non-tensor form) with mbarrier completion: one thread moves a tile's whole record range
__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count)
{
    asm volatile("mbarrier.init.shared.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity)
{
    unsigned done;
    do {
        asm volatile(
            "{\n\t.reg .pred P_OUT;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P_OUT, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, P_OUT;\n\t}"
            : "=r"(done)
            : "r"(mbar), "r"(parity)
            : "memory");
    } while (!done);
}
// global -> shared, `bytes` a multiple of 16 (both addresses 16-byte aligned); completes one phase of `mbar`
__device__ __forceinline__ void bulk_load(void* smemDst, const void* gmemSrc, unsigned bytes, unsigned mbar)
{
    if (bytes) {
        const unsigned dst = (unsigned)__cvta_generic_to_shared(smemDst);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic-proxy reads of the buffer are ordered first
        asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                     "l"(gmemSrc), "r"(bytes), "r"(mbar)
                     : "memory");
    } else {
        asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
    }
}

// Persistent, software-pipelined form: CTA b walks the work items b, b + gridDim.x, ... (work item = tile x instance).
// While tile k is being solved, the inputs of tile k+1 are already on their way into shared memory:
//   top of k      : positions of k+1 (gathered by id with cp.async into the other half of the double-buffered sp)
//   after stretch : stretch records of k+1 (s_srec is free again)
//   after bending : bending records of k+1 (s_brec is free again)
// and the data those copies depend on is fetched one step earlier still (particle ids of k+2 during k, tile descriptor
// of k+2 at the top of k), so that no global-memory latency sits on the critical path of a tile.  Without this the chain
// descriptor -> ids -> positions (three dependent misses, ~2000 cycles) opened every tile: 30 % of all warp stalls (ncu).
#ifndef VT_IT_REGCAP_BLOCKS
#define VT_IT_REGCAP_BLOCKS 4  // resident 256-thread CTAs per SM the register budget is sized for
#endif
// LOG2T: width of a slot row (= tile capacity in particles); NT: threads per CTA.  NT > 2^LOG2T gives the tile extra
// constraint threads: a 16x16-particle tile of a grid cloth is touched by 17x17 = 289 bending constraints, so with 256
// threads 33 of them (two warps) did a second, latency-long bend while six warps idled at the barrier (ncu: 12 % of all
// stall samples); with 320 threads every bend phase is a single trip.
template <int LOG2T, int NT>
__global__ void __launch_bounds__(NT, (VT_IT_REGCAP_BLOCKS * 256) / NT)
iterate_tile_kernel(const float4* __restrict__ predInAll, float4* __restrict__ predOutAll, const TilePlanDev plan,
                    const float* __restrict__ attachSlotsAll, const FrameParams* __restrict__ fp, const Instancing inst,
                    const unsigned totalWork)
{
    vt_pdl_trigger();
    vt_pdl_wait();
    constexpr unsigned T = NT;  // stride of every cooperative loop
    constexpr unsigned TD_WORDS = sizeof(TileDesc) / 4;
    extern __shared__ float4 s_mem[];
    float4* const spBase = s_mem;  // two buffers of maxLocals
    float4* const slots = s_mem + 2 * plan.maxLocals;
    const unsigned slotRows = plan.maxKS > plan.maxKB ? plan.maxKS : plan.maxKB;
    uint4* const s_brec = reinterpret_cast<uint4*>(slots + (slotRows << LOG2T));
    uint2* const s_srec = reinterpret_cast<uint2*>(s_brec + plan.maxBendPerTile);
    unsigned* const s_tdw = reinterpret_cast<unsigned*>(s_srec + ((plan.maxStretchPerTile + 1u) & ~1u));  // ring of 3 descriptors
    const TileDesc* const s_td = reinterpret_cast<const TileDesc*>(s_tdw);
    const unsigned mbarS = (unsigned)__cvta_generic_to_shared(s_tdw + 3 * TD_WORDS);  // stretch / bend records of a tile landed
    const unsigned mbarB = mbarS + 8u;

    const unsigned tid = threadIdx.x;
    const unsigned stride = gridDim.x;
    unsigned w = blockIdx.x;
    if (w >= totalWork) return;
    const float xpbd_bend = fp->xpbdBend;

    // the batched-instances case: work item -> (instance, tile); the plan is shared by all instances
    auto tile_of = [&](unsigned item) { return inst.count > 1 ? item % plan.numTiles : item; };
    auto inst_of = [&](unsigned item) { return inst.count > 1 ? item / plan.numTiles : 0u; };
    auto issue_positions = [&](float4* sp, const TileDesc& d, const float4* predIn, unsigned gid, unsigned hid) {
        if (tid < d.nOwned) cp_async_16(sp + tid, predIn + gid);
        // halo locals start at tileSize in every tile (tile_plan.cpp), so these never touch an entry another thread still reads
        if (tid < d.nHalo) cp_async_16(sp + plan.tileSize + tid, predIn + hid);
        for (unsigned i = tid + T; i < d.nHalo; i += T)
            cp_async_16(sp + plan.tileSize + i, predIn + __ldg(plan.haloIds + d.haloOff + i));
    };
    // record ranges are contiguous per tile: ONE bulk copy each by thread 0 (the per-thread cp.async form spent 3.5 % of
    // the kernel's instructions on ~3.2 16-byte copies per thread and tile)
    auto issue_stretch = [&](const TileDesc& d) {
        if (tid == 0) bulk_load(s_srec, plan.stretchRec + d.stretchOff, ((d.nStretch + 1u) >> 1) * 16u, mbarS);  // even offset: aligned
    };
    auto issue_bend = [&](const TileDesc& d) {
        if (tid == 0) bulk_load(s_brec, plan.bendRec + d.bendOff, d.nBend * 16u, mbarB);
    };
    auto load_ids = [&](const TileDesc& d, unsigned& gid, unsigned& hid) {
        gid = tid < d.nOwned ? __ldg(plan.ownedIds + d.ownedOff + tid) : 0u;
        hid = tid < d.nHalo ? __ldg(plan.haloIds + d.haloOff + tid) : 0u;
    };
    // stretch | bend << 8 constraint counts of this thread's particle; fetched one tile ahead (the compiler sinks a plain
    // load to its first use after the stretch phase, where ncu showed it as the kernel's largest long-scoreboard stall)
    auto load_counts = [&](const TileDesc& d) -> unsigned {
        unsigned c = 0;
        if (tid < d.nOwned) asm volatile("ld.global.nc.u16 %0, [%1];" : "=r"(c) : "l"(plan.cnt16 + d.ownedOff + tid));
        return c;
    };

    // ---- pipeline prologue: descriptors of the first two items, everything of the first, ids of the second
    if (tid == 0) {
        mbar_init(mbarS, 1);
        mbar_init(mbarB, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < TD_WORDS) s_tdw[tid] = __ldg(reinterpret_cast<const unsigned*>(plan.tiles + tile_of(w)) + tid);
    if (tid >= 32 && tid < 32 + TD_WORDS && w + stride < totalWork)
        s_tdw[TD_WORDS + tid - 32] = __ldg(reinterpret_cast<const unsigned*>(plan.tiles + tile_of(w + stride)) + tid - 32);
    __syncthreads();
    unsigned gidCur, gidNext = 0, hidNext = 0;
    unsigned cntCur = load_counts(s_td[0]);
    {
        unsigned hid;
        load_ids(s_td[0], gidCur, hid);
        issue_positions(spBase, s_td[0], predInAll + (size_t)inst_of(w) * inst.particles, gidCur, hid);
        cp_async_commit();  // P(0)
        issue_stretch(s_td[0]);  // S(0)
        issue_bend(s_td[0]);     // B(0)
        if (w + stride < totalWork) load_ids(s_td[1], gidNext, hidNext);
    }

    for (unsigned k = 0; w < totalWork; k++, w += stride) {
        const unsigned slotCur = k % 3, slotNext = (k + 1) % 3, slotNN = (k + 2) % 3;
        const TileDesc& td = s_td[slotCur];
        float4* const sp = spBase + (k & 1u) * plan.maxLocals;
        const bool hasNext = w + stride < totalWork;
        const bool hasNN = w + 2 * (size_t)stride < totalWork;
        const size_t instOff = (size_t)inst_of(w) * inst.particles;
        const bool owner = tid < td.nOwned;

        // descriptor of k+2 (lands in the ring after the first barrier), positions of k+1
        unsigned tdWord = 0;
        if (hasNN && tid < TD_WORDS) tdWord = __ldg(reinterpret_cast<const unsigned*>(plan.tiles + tile_of(w + 2 * stride)) + tid);
        if (hasNext)
            issue_positions(spBase + ((k + 1) & 1u) * plan.maxLocals, s_td[slotNext],
                            predInAll + (size_t)inst_of(w + stride) * inst.particles, gidNext, hidNext);
        cp_async_commit();      // P(k+1)
        const unsigned cntNext = hasNext ? load_counts(s_td[slotNext]) : 0u;
        cp_async_wait_group<1>();  // P(k) has landed; P(k+1) may still be in flight
        mbar_wait(mbarS, k & 1u);  // S(k) has landed
        __syncthreads();
        if (hasNN && tid < TD_WORDS) s_tdw[slotNN * TD_WORDS + tid] = tdWord;
        const unsigned cntS = cntCur & 0xffu, cntB = cntCur >> 8;

        // ---- SolveStretch_Kernel, VtClothSolverGPU.cu L76-101: one evaluation per constraint
        // two constraints per trip (their dependent chains -- sqrt, reciprocals -- interleave), then at most one single
        {
            unsigned c = tid;
            for (; c + T < td.nStretch; c += 2 * T) {
                const uint2 r0 = s_srec[c], r1 = s_srec[c + T];
                bool ok0 = true, ok1 = true;
                const StretchWork w0 = stretch_begin(r0, sp, ok0), w1 = stretch_begin(r1, sp, ok1);
                StretchOut o0 = stretch_inactive(), o1 = o0;
                if (w0.active || w1.active) {
                    o0 = stretch_finish(r0, w0, ok0);
                    o1 = stretch_finish(r1, w1, ok1);
                    if (!w0.active) o0 = stretch_inactive();
                    if (!w1.active) o1 = stretch_inactive();
                }
                if (ok0) stretch_store<LOG2T>(r0, o0, slots, plan.maxKS);
                if (ok1) stretch_store<LOG2T>(r1, o1, slots, plan.maxKS);
#if !VT_FAST_MATH
                if (!ok0) stretch_slow_to_slots<LOG2T>(r0, sp, slots, plan.maxKS);
                if (!ok1) stretch_slow_to_slots<LOG2T>(r1, sp, slots, plan.maxKS);
#endif
            }
            if (c < td.nStretch) {
                const uint2 r0 = s_srec[c];
                bool ok0 = true;
                const StretchWork w0 = stretch_begin(r0, sp, ok0);
                StretchOut o0 = stretch_inactive();
                if (w0.active) o0 = stretch_finish(r0, w0, ok0);
                if (ok0) stretch_store<LOG2T>(r0, o0, slots, plan.maxKS);
#if !VT_FAST_MATH
                else stretch_slow_to_slots<LOG2T>(r0, sp, slots, plan.maxKS);
#endif
            }
        }
        __syncthreads();

        if (hasNext) issue_stretch(s_td[slotNext]);  // S(k+1): every thread is past its last read of s_srec
        unsigned gidNN = 0, hidNN = 0;
        if (hasNN) load_ids(s_td[slotNN], gidNN, hidNN);  // consumed at the top of k+1

        vec3 delta = V3(0, 0, 0);
        float count = 0;
        if (owner) {
            sum_slots<LOG2T>(slots, tid, cntS, delta, count);
            // SolveAttachment_Kernel, L218-234: per-particle, no slot needed
            if (plan.hasAttach) {
                const float4 mine = sp[tid];
                const float* attachSlotPositions = attachSlotsAll + (size_t)inst_of(w) * inst.slots * 3;
                const float lrs = fp->P.longRangeStretchiness;
                const unsigned a1 = __ldg(plan.attOff + td.baseOff + tid + 1);
                for (unsigned a = __ldg(plan.attOff + td.baseOff + tid); a < a1; a++) {
                    const uint2 r = __ldg(plan.attachRec + td.attachOff + a);
                    vec3 corr;
                    if (attach_eval(V3(mine), mine.w, load3(attachSlotPositions, r.x), __uint_as_float(r.y), lrs, corr)) {
                        delta += corr;
                        count += 1.0f;
                    }
                }
            }
        }
        mbar_wait(mbarB, k & 1u);  // B(k) has landed
        __syncthreads();           // slots are reused by the bending phase

        // ---- SolveBending_Kernel, L128-188
        for (unsigned c = tid; c < td.nBend; c += T) bend_to_slots<LOG2T>(s_brec[c], sp, slots, xpbd_bend, plan.maxKB);
        __syncthreads();

        if (hasNext) issue_bend(s_td[slotNext]);  // B(k+1)
        if (owner) {
            sum_slots<LOG2T>(slots, tid, cntB, delta, count);
            // ApplyDeltas_Kernel, L257-263
            const float4 mine = sp[tid];
            vec3 p = V3(mine);
            if (count > 0) p += delta / count * fp->P.relaxationFactor;
            predOutAll[instOff + gidCur] = F4(p, mine.w);
        }
        gidCur = gidNext;
        gidNext = gidNN;
        hidNext = hidNN;
        cntCur = cntNext;
    }
    cp_async_wait_all();
}

// ---------------------------------------------------------------- implicit-grid Jacobi iteration (round 2)
//
// ncu on the record-driven kernel above (round 1/2): only a third of its executed instructions are floating-point math; the
// rest decodes records, computes slot addresses, tests ordinals and keeps a three-stage tile pipeline alive.  For a grid
// cloth -- the only kind the reference creates, VtClothObjectGPU.hpp L75-132 -- all of that is implied by the vertex
// coordinates (grid_plan.hpp), so this kernel has no index records at all:
//   * a tile owns 15 x 15 particles and evaluates the 16 x 16 "bundles" of constraints generated at the vertices
//     (x0-1 .. x0+14) x (y0-1 .. y0+14): thread (bx, by) holds the four corners c00 c01 c10 c11 of its quad in registers and
//     evaluates the bundle's four stretch constraints (vertical c00-c01, horizontal c00-c10, diagonal c00-c11, anti-diagonal
//     c01-c10) and its bending constraint (p0 p1 p2 p3 = c00 c11 c01 c10) -- one trip, every thread the same work, and the
//     edge vectors are shared: the bend's e = p3 - p2 is minus the anti-diagonal difference (same length, same bits), its
//     normals are cross products of the stretch differences;
//   * corrections for the three far corners go to eight slot arrays indexed by THREAD id (static addresses), those for the
//     thread's own particle c00 stay in registers: particle (x, y) sums, in constraint-id order, the slots of bundles
//     (x-1,y-1), (x-1,y), (x,y-1) and then its own -- the order of the oracle, bit-identical to the record-driven kernel;
//   * two barriers per tile; the next tile's positions, rest lengths and rest angles are in flight (cp.async) while the
//     current one is solved; tile coordinates are arithmetic on the work index.
// Algorithmic bytes per particle: 16 (position in) + 16 (out) + 16 (rest lengths) + 4 (rest angle) = 52.
constexpr int GRID_B = GRID_TILE + 1;  // bundles per tile side
constexpr int GRID_V = GRID_TILE + 2;  // staged vertices per tile side
struct GridTileCoord {
    unsigned base, planBase, instance;  // first particle of the tile's cloth in the state arrays / in the plan arrays; instance
    int side, x0, y0;                   // vertices per cloth side; first owned vertex of the tile
    unsigned pad[2];
};
// 2 x 17 x 17 positions, 2 x 256 rest-length quadruples, 8 x 256 slots, 2 x 256 rest angles, three tile coordinates, the cloth table
constexpr size_t GRID_SMEM_BYTES = sizeof(float4) * (2 * GRID_V * GRID_V + 10 * 256 + 256 / 2) + 3 * sizeof(GridTileCoord) +
                                   sizeof(GridCloth) * GRID_MAX_CLOTHS;

struct GridStretchOut {
    vec3 c1, c2;
    float flag;
};
struct GridBendOut {
    vec3 c0, c1, c2, c3;
    float flag;
};
#if !VT_FAST_MATH
// rare paths (degenerate or non-finite geometry): the branchy evaluators, out of line
__device__ __noinline__ GridStretchOut grid_stretch_slow(float4 pa, float4 pb, float rest)
{
    GridStretchOut o;
    o.c1 = o.c2 = V3(0, 0, 0);
    o.flag = stretch_eval(V3(pa), V3(pb), pa.w, pb.w, rest, o.c1, o.c2) ? 1.0f : 0.0f;
    return o;
}
__device__ __noinline__ GridBendOut grid_bend_slow(float4 p0, float4 p1, float4 p2, float4 p3, float restAngle, float xpbd_bend)
{
    GridBendOut o;
    o.c0 = o.c1 = o.c2 = o.c3 = V3(0, 0, 0);
    o.flag = bend_eval(V3(p0), V3(p1), V3(p2), V3(p3), p0.w, p1.w, p2.w, p3.w, restAngle, xpbd_bend, o.c0, o.c1, o.c2, o.c3) ? 1.0f : 0.0f;
    return o;
}
#endif

#ifndef VT_GRID_BLOCKS
#define VT_GRID_BLOCKS 3  // resident CTAs per SM the register budget is sized for (80 registers; at 4 x 64 the kernel spills: 41.0 vs 39.2 us)
#endif
// BX x BY bundles per tile = (BX - 1) x (BY - 1) owned particles: 16 x 16, or 15 x 17 (grid_plan.hpp: the cloth side decides)
template <int BX, int BY>
__global__ void __launch_bounds__(256, VT_GRID_BLOCKS)
iterate_grid_kernel(float4* __restrict__ predA, float4* __restrict__ predB, const GridPlanDev plan,
                    const float* __restrict__ attachSlotsAll, const FrameParams* __restrict__ fp, const Instancing inst,
                    const unsigned totalWork, const ddpeer::StripArgs strip, const unsigned iterations, unsigned* __restrict__ gridBarrier)
{
    vt_pdl_trigger();  // the next kernel may set itself up while this one runs
    constexpr unsigned NT = 256;
    constexpr int TX = BX - 1, TY = BY - 1;  // owned particles per tile along x (slow index) and y
    constexpr int VY = BY + 1;               // staged vertices per tile row
    static_assert(BX * BY <= (int)NT && (BX + 1) * (BY + 1) <= GRID_V * GRID_V, "tile shape exceeds the CTA or its staging buffers");
    extern __shared__ float4 s_mem[];  // GRID_SMEM_BYTES, carved below (more than the 48 KB a static allocation may take)
    float4(*const s_sp)[GRID_V * GRID_V] = reinterpret_cast<float4(*)[GRID_V * GRID_V]>(s_mem);
    float4(*const s_rest)[NT] = reinterpret_cast<float4(*)[NT]>(s_mem + 2 * GRID_V * GRID_V);
    float4(*const s_slots)[NT] = reinterpret_cast<float4(*)[NT]>(s_mem + 2 * GRID_V * GRID_V + 2 * NT);
    float(*const s_angle)[NT] = reinterpret_cast<float(*)[NT]>(s_mem + 2 * GRID_V * GRID_V + 10 * NT);
    GridTileCoord* const s_tile = reinterpret_cast<GridTileCoord*>(s_mem + 2 * GRID_V * GRID_V + 10 * NT + NT / 2);  // ring of three
    GridCloth* const s_cloth = reinterpret_cast<GridCloth*>(s_tile + 3);

    const unsigned tid = threadIdx.x;
    const bool live = BX * BY == (int)NT || tid < (unsigned)(BX * BY);  // 15 x 17 leaves the last thread without a bundle
    const int by = live ? (int)(tid % (unsigned)BY) : 0, bx = live ? (int)(tid / (unsigned)BY) : 0;
    const unsigned stride = gridDim.x;
    if (blockIdx.x >= totalWork) return;  // (never with more than one iteration per launch: the grid is at most totalWork)
    if (tid < plan.numCloths) s_cloth[tid] = plan.cloths[tid];
    // vertices outside the cloth are never staged: whatever their entries hold must at least be finite
    for (unsigned i = tid; i < 2 * GRID_V * GRID_V + 2 * NT; i += NT) s_mem[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    const float xpbd_bend = fp->xpbdBend;
    const float relaxation = fp->P.relaxationFactor;
    const float lrs = fp->P.longRangeStretchiness;
    // everything above reads data that no kernel of the frame writes (cloth table, frame parameters); the predecessor's output
    // (predIn, and the exchange state of a decomposed cloth) is first touched below
    vt_pdl_wait();
    __syncthreads();

    // ---- one strip of a decomposed cloth (dd_peer.cuh): this rank owns the tile rows [tileRowBegin, tileRowEnd) of the single
    // cloth.  The tile rows next to another rank come FIRST in the work order: their outermost particle rows go straight into
    // the neighbours' output arrays and the launch publishes as soon as the last of them is done, so the transfer and the
    // neighbours' wait overlap the interior tiles.  The neighbours' rows this launch reads were published by their previous
    // launch, early in it, for the same reason.
    unsigned stripSeq = 0, stripBoundaryTiles = 0, stripBoundaryRows = 0;
    bool stripHasA = false;
    if (strip.enabled) {
        stripHasA = strip.up >= 0;
        const bool hasB = strip.down >= 0 && (!stripHasA || strip.tileRowEnd - 1 != strip.tileRowBegin);
        stripBoundaryRows = (stripHasA ? 1u : 0u) + (hasB ? 1u : 0u);
        stripBoundaryTiles = stripBoundaryRows * s_cloth[0].tilesY;
        if (tid == 0) {
            stripSeq = strip.ctl->seq;
            ddpeer::strip_wait_for(strip.localFlags, strip.up, stripSeq, strip.ctl, strip.timeoutNs);
            ddpeer::strip_wait_for(strip.localFlags, strip.down, stripSeq, strip.ctl, strip.timeoutNs);
            if (stripBoundaryTiles == 0 && blockIdx.x == 0) ddpeer::strip_publish(strip.T, stripSeq + 1);
        }
        __syncthreads();
    }

    const float4* predInAll = predA;  // input / output of the iteration in progress
    float4* predOutAll = predB;
    // work item -> first particle of its cloth (instance included), grid side, tile origin.  One thread does this (an integer
    // division and a table walk) two tiles ahead and leaves the result in shared memory for the others.
    auto locate = [&](unsigned item) {
        if (strip.enabled) {  // boundary tile rows first, then the interior rows in ascending order
            const GridCloth g = s_cloth[0];
            const unsigned r = item / g.tilesY, c = item - r * g.tilesY;
            unsigned row;
            if (r < stripBoundaryRows)
                row = (r == 0 && stripHasA) ? strip.tileRowBegin : strip.tileRowEnd - 1;
            else
                row = strip.tileRowBegin + (r - stripBoundaryRows) + (stripHasA ? 1u : 0u);
            GridTileCoord t;
            t.base = g.base;
            t.planBase = g.base;
            t.instance = 0;
            t.side = (int)g.side;
            t.x0 = (int)row * TX;  // (strips are cut in rows of 15: launched with the square shape only)
            t.y0 = (int)c * TY;
            return t;
        }
        unsigned tile = item, in = 0;
        if (inst.count > 1) {
            in = item / plan.numTiles;
            tile = item - in * plan.numTiles;
        }
        unsigned c = 0;
        while (c + 1 < plan.numCloths && tile >= s_cloth[c + 1].firstTile) c++;
        const GridCloth g = s_cloth[c];
        const unsigned lt = tile - g.firstTile, tx = lt / g.tilesY;
        GridTileCoord t;
        t.base = in * inst.particles + g.base;
        t.planBase = g.base;  // attach CSR, rest lengths and angles cover one instance
        t.instance = in;
        t.side = (int)g.side;
        t.x0 = (int)tx * TX;
        t.y0 = (int)(lt - tx * g.tilesY) * TY;
        return t;
    };
    // stage the 17 x 17 vertices around the tile, the bundle's rest lengths and rest angle (vertices outside the cloth are skipped)
    const unsigned spAddr = (unsigned)__cvta_generic_to_shared(&s_sp[0][bx * VY + by]);
    const unsigned restAddr = (unsigned)__cvta_generic_to_shared(&s_rest[0][tid]);
    const unsigned angleAddr = (unsigned)__cvta_generic_to_shared(&s_angle[0][tid]);
    constexpr unsigned SP_BYTES = GRID_V * GRID_V * 16, V16 = VY * 16;
    auto cp16 = [](unsigned dst, const void* src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src)); };
    auto cp4 = [](unsigned dst, const void* src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src)); };
    auto issue_tile = [&](unsigned buf, const GridTileCoord& t) {
        const int gx = t.x0 - 1 + bx, gy = t.y0 - 1 + by;
        const bool inX = live && (unsigned)gx < (unsigned)t.side, inY = (unsigned)gy < (unsigned)t.side;
        const bool inX1 = (unsigned)(gx + 1) < (unsigned)t.side, inY1 = (unsigned)(gy + 1) < (unsigned)t.side;
        const int idx = gx * t.side + gy;
        const float4* src = predInAll + t.base + idx;  // (predInAll: this iteration's input, set by the loop below)
        const unsigned dst = spAddr + buf * SP_BYTES;
        if (inX && inY) {
            cp16(dst, src);
            cp16(restAddr + buf * (NT * 16), plan.rest4 + t.planBase + idx);
            if (plan.restAngle) cp4(angleAddr + buf * (NT * 4), plan.restAngle + t.planBase + idx);
        }
        if (by == BY - 1 && inX && inY1) cp16(dst + 16, src + 1);
        if (bx == BX - 1 && live) {
            if (inX1 && inY) cp16(dst + V16, src + t.side);
            if (by == BY - 1 && inX1 && inY1) cp16(dst + V16 + 16, src + t.side + 1);
        }
    };

    // Several Jacobi iterations per launch (`iterations` > 1; a single cloth or batch without a strip exchange): the CTAs of the
    // launch are all resident (the grid is one wave), so an iteration ends at a grid-wide barrier instead of a kernel
    // boundary -- a launch and its drain cost ~5 us of the 36 us an iteration takes at 1M particles, the barrier ~2.
    // Iteration `it` reads predA and writes predB when it is even, the other way round when it is odd.
    for (unsigned it = 0; it < iterations; it++) {
    predInAll = (it & 1u) ? predB : predA;
    predOutAll = (it & 1u) ? predA : predB;
    unsigned w = blockIdx.x;
    // ring of three tile coordinates: tile k + 2 is written while k and k + 1 are still being read
    if (tid == 0) {
        s_tile[0] = locate(w);
        if (w + stride < totalWork) s_tile[1] = locate(w + stride);
    }
    __syncthreads();
    issue_tile(0, s_tile[0]);
    cp_async_commit();

    unsigned slotCur = 0;  // k % 3
    for (unsigned k = 0; w < totalWork; k++, w += stride) {
        const unsigned buf = k & 1u;
        const unsigned slotNext = slotCur == 2 ? 0 : slotCur + 1, slotNN = slotNext == 2 ? 0 : slotNext + 1;
        if (w + stride < totalWork) issue_tile(buf ^ 1u, s_tile[slotNext]);
        cp_async_commit();
        cp_async_wait_group<1>();  // this tile has landed; the next one may still be in flight
        __syncthreads();           // ... for every thread; and every thread is past the sums of the previous tile
        // (whose coordinates sat in the ring slot that now takes tile k + 2)
        if (tid == 0 && w + 2 * (size_t)stride < totalWork) s_tile[slotNN] = locate(w + 2 * stride);
        const int sideCur = s_tile[slotCur].side, x0Cur = s_tile[slotCur].x0, y0Cur = s_tile[slotCur].y0;

        // ---- which constraints of this bundle exist (cloth border, partially filled tiles)
        const int gx = x0Cur - 1 + bx, gy = y0Cur - 1 + by;
        const bool inGrid = live && (unsigned)gx < (unsigned)sideCur && (unsigned)gy < (unsigned)sideCur;
        const bool vV = inGrid && gy + 1 < sideCur;  // (x,y)-(x,y+1)
        const bool vH = inGrid && gx + 1 < sideCur;  // (x,y)-(x+1,y)
        const bool vQ = vV && vH;                    // both diagonals and the bending constraint of the quad

        const float4* sp = &s_sp[buf][bx * VY + by];
        const float4 c00 = sp[0], c01 = sp[1], c10 = sp[VY], c11 = sp[VY + 1];
        const float4 rest = s_rest[buf][tid];
        const float restAngle = plan.restAngle ? s_angle[buf][tid] : plan.uniformAngle;

        // SolveStretch_Kernel, VtClothSolverGPU.cu L76-101 (four constraints) and SolveBending_Kernel, L128-188 (one)
        const vec3 dV = V3(c00) - V3(c01), dH = V3(c00) - V3(c10), dD = V3(c00) - V3(c11), dA = V3(c01) - V3(c10);
        GridStretchOut oV, oH, oD, oA;
        GridBendOut oB;
#if VT_FAST_MATH
        {
            const bool aV = stretch_eval_flagged(V3(c00), V3(c01), c00.w, c01.w, rest.x, oV.c1, oV.c2) && vV;
            const bool aH = stretch_eval_flagged(V3(c00), V3(c10), c00.w, c10.w, rest.y, oH.c1, oH.c2) && vH;
            const bool aD = stretch_eval_flagged(V3(c00), V3(c11), c00.w, c11.w, rest.z, oD.c1, oD.c2) && vQ;
            const bool aA = stretch_eval_flagged(V3(c01), V3(c10), c01.w, c10.w, rest.w, oA.c1, oA.c2) && vQ;
            if (!aV) oV.c1 = oV.c2 = V3(0, 0, 0);
            if (!aH) oH.c1 = oH.c2 = V3(0, 0, 0);
            if (!aD) oD.c1 = oD.c2 = V3(0, 0, 0);
            if (!aA) oA.c1 = oA.c2 = V3(0, 0, 0);
            oV.flag = aV ? 1.0f : 0.0f;
            oH.flag = aH ? 1.0f : 0.0f;
            oD.flag = aD ? 1.0f : 0.0f;
            oA.flag = aA ? 1.0f : 0.0f;
            oB.c0 = oB.c1 = oB.c2 = oB.c3 = V3(0, 0, 0);
            const bool aB = vQ && bend_eval_dv(-dA, length(dA), -dV, -dH, V3(c10) - V3(c11), V3(c01) - V3(c11), c00.w, c11.w, c01.w,
                                               c10.w, restAngle, xpbd_bend, oB.c0, oB.c1, oB.c2, oB.c3);
            if (!aB) oB.c0 = oB.c1 = oB.c2 = oB.c3 = V3(0, 0, 0);
            oB.flag = aB ? 1.0f : 0.0f;
        }
#else
        {
            // the bend first (it needs the most registers; its three far-corner corrections leave for their slots at once),
            // then the four stretches.  p0 p1 p2 p3 = c00 c11 c01 c10: e = p3 - p2 = -dA, p2 - p0 = -dV, p3 - p0 = -dH.
            // ONE rarely taken branch per bundle redoes whatever left the fast-path windows (degenerate or non-finite geometry).
            bool okLenA = true;
            const float lenA = vt_sqrt_dist_u(dot(dA, dA), okLenA);
            bool okB = okLenA;
            const bool aB = bend_eval_dv_u(-dA, lenA, -dV, -dH, V3(c10) - V3(c11), V3(c01) - V3(c11), c00.w, c11.w, c01.w, c10.w,
                                           restAngle, xpbd_bend, oB.c0, oB.c1, oB.c2, oB.c3, okB) && vQ;
            if (!aB) oB.c0 = oB.c1 = oB.c2 = oB.c3 = V3(0, 0, 0);
            oB.flag = aB ? 1.0f : 0.0f;
            s_slots[5][tid] = F4(oB.c1, oB.flag);  // bending p1 -> c11
            s_slots[6][tid] = F4(oB.c3, oB.flag);  // bending p3 -> c10
            s_slots[7][tid] = F4(oB.c2, oB.flag);  // bending p2 -> c01

            bool okV = true, okH = true, okD = true, okA = okLenA;
            StretchHalf hV = stretch_begin_len(dV, vt_sqrt_dist_u(dot(dV, dV), okV), c00.w, c01.w, rest.x);
            StretchHalf hH = stretch_begin_len(dH, vt_sqrt_dist_u(dot(dH, dH), okH), c00.w, c10.w, rest.y);
            StretchHalf hD = stretch_begin_len(dD, vt_sqrt_dist_u(dot(dD, dD), okD), c00.w, c11.w, rest.z);
            StretchHalf hA = stretch_begin_len(dA, lenA, c01.w, c10.w, rest.w);
            hV.active = hV.active && vV;
            hH.active = hH.active && vH;
            hD.active = hD.active && vQ;
            hA.active = hA.active && vQ;
            oV.c1 = oV.c2 = oH.c1 = oH.c2 = oD.c1 = oD.c2 = oA.c1 = oA.c2 = V3(0, 0, 0);
            if (hV.active || hH.active || hD.active || hA.active) {  // a freely falling, undeformed cloth keeps every distance at rest: no divisions then
                // (distance - rest) / (w1 + w2): the four denominators are powers of two for unit inverse masses, a division by
                // which is an exact multiplication -- tested once for the bundle
                float sV, sH, sD, sA;
                const bool pV = vt_pow2_rcp(hV.denom, sV), pH = vt_pow2_rcp(hH.denom, sH), pD = vt_pow2_rcp(hD.denom, sD),
                           pA = vt_pow2_rcp(hA.denom, sA);
                const bool pow2 = pV && pH && pD && pA;
                stretch_finish_grid_u(hV, c00.w, c01.w, rest.x, pow2, sV, oV.c1, oV.c2, okV);
                stretch_finish_grid_u(hH, c00.w, c10.w, rest.y, pow2, sH, oH.c1, oH.c2, okH);
                stretch_finish_grid_u(hD, c00.w, c11.w, rest.z, pow2, sD, oD.c1, oD.c2, okD);
                stretch_finish_grid_u(hA, c01.w, c10.w, rest.w, pow2, sA, oA.c1, oA.c2, okA);
            }
            oV.flag = hV.active ? 1.0f : 0.0f;
            oH.flag = hH.active ? 1.0f : 0.0f;
            oD.flag = hD.active ? 1.0f : 0.0f;
            oA.flag = hA.active ? 1.0f : 0.0f;
            if ((vV && !okV) || (vH && !okH) || (vQ && !(okD && okA && okB))) {
                if (vV && !okV) oV = grid_stretch_slow(c00, c01, rest.x);
                if (vH && !okH) oH = grid_stretch_slow(c00, c10, rest.y);
                if (vQ && !okD) oD = grid_stretch_slow(c00, c11, rest.z);
                if (vQ && !okA) oA = grid_stretch_slow(c01, c10, rest.w);
                if (vQ && !okB) {
                    oB = grid_bend_slow(c00, c11, c01, c10, restAngle, xpbd_bend);
                    s_slots[5][tid] = F4(oB.c1, oB.flag);
                    s_slots[6][tid] = F4(oB.c3, oB.flag);
                    s_slots[7][tid] = F4(oB.c2, oB.flag);
                }
            }
        }
#endif
#if VT_FAST_MATH
        s_slots[5][tid] = F4(oB.c1, oB.flag);  // bending p1    -> c11
        s_slots[6][tid] = F4(oB.c3, oB.flag);  // bending p3    -> c10
        s_slots[7][tid] = F4(oB.c2, oB.flag);  // bending p2    -> c01
#endif
        // far corners: slots by thread id; own particle: registers
        s_slots[0][tid] = F4(oD.c2, oD.flag);  // diagonal      -> c11
        s_slots[1][tid] = F4(oH.c2, oH.flag);  // horizontal    -> c10
        s_slots[2][tid] = F4(oA.c2, oA.flag);  // anti-diagonal -> c10
        s_slots[3][tid] = F4(oV.c2, oV.flag);  // vertical      -> c01
        s_slots[4][tid] = F4(oA.c1, oA.flag);  // anti-diagonal -> c01
        __syncthreads();

        // ---- particle (gx, gy): stretch constraints generated at (gx-1,gy-1), (gx-1,gy), (gx,gy-1), (gx,gy) in that (= id)
        // order, attachments, bending constraints of the same four quads; ApplyDeltas_Kernel, L257-263
        if (bx > 0 && by > 0 && inGrid) {
            vec3 delta = V3(0, 0, 0);
            float count = 0;
            auto add = [&](const float4 v) {
                delta.x += v.x;
                delta.y += v.y;
                delta.z += v.z;
                count += v.w;
            };
            add(s_slots[0][tid - BY - 1]);
            add(s_slots[1][tid - BY]);
            add(s_slots[2][tid - BY]);
            add(s_slots[3][tid - 1]);
            add(s_slots[4][tid - 1]);
            add(F4(oV.c1, oV.flag));
            add(F4(oH.c1, oH.flag));
            add(F4(oD.c1, oD.flag));
            const unsigned local = (unsigned)(gx * sideCur + gy);
            if (plan.hasAttach) {  // SolveAttachment_Kernel, L218-234
                const float* attachSlotPositions = attachSlotsAll + (size_t)s_tile[slotCur].instance * inst.slots * 3;
                const unsigned attCur = s_tile[slotCur].planBase;
                const unsigned a1 = __ldg(plan.attOff + attCur + local + 1);
                for (unsigned a = __ldg(plan.attOff + attCur + local); a < a1; a++) {
                    const uint2 r = __ldg(plan.attachRec + a);
                    vec3 corr;
                    if (attach_eval(V3(c00), c00.w, load3(attachSlotPositions, r.x), __uint_as_float(r.y), lrs, corr)) {
                        delta += corr;
                        count += 1.0f;
                    }
                }
            }
            add(s_slots[5][tid - BY - 1]);
            add(s_slots[6][tid - BY]);
            add(s_slots[7][tid - 1]);
            add(F4(oB.c0, oB.flag));
            vec3 p = V3(c00);
            if (count > 0) p += delta / count * relaxation;
            const float4 result = F4(p, c00.w);
            const unsigned idx = s_tile[slotCur].base + local;
            predOutAll[idx] = result;
            if (strip.enabled) {
                if (strip.gatherAll) {  // last iteration of a substep: the all-gather of the results rides on the epilogue
                    for (int q = 0; q < strip.T.world; q++)
                        if (q != strip.T.rank) strip.T.pred[strip.which][q][idx] = result;
                } else {  // outermost owned rows: also into the neighbour's array, at the same index
                    if ((unsigned)gx == strip.rowFirst && strip.up >= 0) strip.T.pred[strip.which][strip.up][idx] = result;
                    if ((unsigned)gx == strip.rowLast && strip.down >= 0) strip.T.pred[strip.which][strip.down][idx] = result;
                }
            }
        }
        if (strip.enabled && w < stripBoundaryTiles) {  // the boundary tile that finishes last publishes this launch
            __threadfence_system();
            __syncthreads();
            if (tid == 0 && atomicAdd(&strip.ctl->sendsDone, 1u) == stripBoundaryTiles - 1) {
                __threadfence_system();
                strip.ctl->sendsDone = 0;
                ddpeer::strip_publish(strip.T, stripSeq + 1);
            }
        }
        slotCur = slotNext;
    }
    cp_async_wait_all();
    if (it + 1 < iterations) {  // grid-wide barrier: every CTA's results of this iteration are in memory before anyone reads them
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            atomicAdd(gridBarrier, 1u);
            const unsigned target = (it + 1u) * gridDim.x;
            const long long t0 = clock64();
            unsigned seen;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(gridBarrier) : "memory");
            } while (seen < target && clock64() - t0 < (1ll << 32));  // (~2 s: a launch that is not fully resident must not hang the device)
        }
        __syncthreads();
    }
    }  // iterations
    if (strip.enabled && tid == 0 && atomicAdd(&strip.ctl->exits, 1u) == gridDim.x - 1) {  // last block out completes the launch
        strip.ctl->exits = 0;
        strip.ctl->seq = stripSeq + 1;
    }
}

__global__ void __launch_bounds__(PB) end_substep_kernel(const float4* __restrict__ predIn, float4* __restrict__ pos4,
                                                         float4* __restrict__ predNext,
                                                         int last, float* __restrict__ positionsOut,
                                                         float* __restrict__ velocitiesOut,
                                                         float* __restrict__ predictedOut,
                                                         const FrameParams* __restrict__ fp, unsigned n)
{
    vt_pdl_trigger();
    vt_pdl_wait();
    const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n) return;
    const VtSimParams& P = fp->P;
    const float dt = fp->substepTime;
    const float4 pr = predIn[id];
    const float4 po = pos4[id];
    vec3 newPos, vel;
    finalize_point(V3(pr), V3(po), dt, P.maxSpeed, P.damping, newPos, vel);  // Finalize_Kernel, .cu L396-406
    pos4[id] = F4(newPos, po.w);
    if (last) {
        store3(positionsOut, id, newPos);
        store3(velocitiesOut, id, vel);
        store3(predictedOut, id, V3(pr));
        predNext[id] = F4(newPos, po.w);  // the frame's positions as float4 (a decomposed cloth gathers them from here)
    } else {
        // PredictPositions of the next substep, .cu L51-52
        vel = vel + V3(P.gravity[0], P.gravity[1], P.gravity[2]) * dt;
        predNext[id] = F4(newPos + vel * dt, po.w);
    }
}

// ComputeTriangleNormals + ComputeVertexNormals (.cu L419-450) as a gather: vertex v sums the face normals of
// its incident triangles in ascending triangle id (the oracle's order), then normalises.
__global__ void __launch_bounds__(PB) normals_kernel(const float4* __restrict__ pos4, const unsigned* __restrict__ indices,
                                                     const unsigned* __restrict__ vtxTriOff,
                                                     const unsigned* __restrict__ vtxTris, float* __restrict__ normalsOut,
                                                     unsigned n)
{
    vt_pdl_trigger();
    vt_pdl_wait();
    const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n) return;
    pos4 += (size_t)blockIdx.y * n;  // instance (indices / CSR are per-instance local)
    normalsOut += (size_t)blockIdx.y * n * 3;
    vec3 sum = V3(0, 0, 0);
    const unsigned t1 = __ldg(vtxTriOff + id + 1);
    for (unsigned t = __ldg(vtxTriOff + id); t < t1; t++) {
        const unsigned tri = __ldg(vtxTris + t);
        const vec3 p1 = V3(__ldg(pos4 + __ldg(indices + 3 * (size_t)tri)));
        const vec3 p2 = V3(__ldg(pos4 + __ldg(indices + 3 * (size_t)tri + 1)));
        const vec3 p3 = V3(__ldg(pos4 + __ldg(indices + 3 * (size_t)tri + 2)));
        sum += cross(p2 - p1, p3 - p1);
    }
    store3(normalsOut, id, normalize(sum));
}

#if !VT_FAST_MATH  // plumbing kernels exist in the exact build only
__global__ void __launch_bounds__(PB) gather_by_id_kernel(const float4* __restrict__ src, const unsigned* __restrict__ ids, unsigned n,
                                                          float4* __restrict__ out)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = src[ids[i]];
}

__global__ void __launch_bounds__(PB) scatter_by_id_kernel(const float4* __restrict__ in, const unsigned* __restrict__ ids, unsigned n,
                                                           float4* __restrict__ dst)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[ids[i]] = in[i];
}

__global__ void __launch_bounds__(PB) unpack_float4_kernel(const float4* __restrict__ in, float* __restrict__ packed3, unsigned n)
{
    const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id < n) store3(packed3, id, V3(in[id]));
}

__global__ void __launch_bounds__(PB) pack_float4_kernel(const float* __restrict__ packed3, float4* __restrict__ out, unsigned n)
{
    const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n) return;
    out[id] = F4(load3(packed3, id), 0.0f);
}

#endif

}  // namespace

void launch_prepare_inputs(const FusedLaunch& L, const VtSDFCollider* colliders, PreparedCollider* prepared,
                           const float* slotPositions, float* slotPositionsOut, unsigned numSlotFloats, const FrameParams* fp)
{
    launch_pdl(prepare_inputs_kernel, dim3(1), dim3(256), 0, L.stream, colliders, prepared, slotPositions, slotPositionsOut, numSlotFloats, fp);
}

void launch_begin_frame(const FusedLaunch& L, const float* positions, const float* velocities, const float* invMasses,
                        float4* pos4, float4* pred, const PreparedCollider* colliders, const FrameParams* fp)
{
    launch_pdl(begin_frame_kernel, dim3(pgrid(L.numParticles)), dim3(PB), 0, L.stream, positions, velocities, invMasses, pos4, pred,
               colliders, fp, L.numParticles);
}

void launch_collide(const FusedLaunch& L, const float4* predIn, float4* predOut, const float4* pos4,
                    const unsigned* neighbors, const PreparedCollider* colliders, const FrameParams* fp, bool selfCollision,
                    const unsigned* subset, unsigned subsetCount)
{
    const unsigned n = subset ? subsetCount : L.numParticles;
    if (!n) return;
    launch_pdl(collide_kernel, dim3(pgrid(n)), dim3(PB), 0, L.stream, predIn, predOut, pos4, neighbors, colliders, fp, L.numParticles,
               selfCollision ? 1 : 0, subset, subsetCount, 0u);
}

void launch_collide_range(const FusedLaunch& L, const float4* predIn, float4* predOut, const float4* pos4, const unsigned* neighbors,
                          const PreparedCollider* colliders, const FrameParams* fp, bool selfCollision, unsigned begin, unsigned count)
{
    if (!count) return;
    launch_pdl(collide_kernel, dim3(pgrid(count)), dim3(PB), 0, L.stream, predIn, predOut, pos4, neighbors, colliders, fp, L.numParticles,
               selfCollision ? 1 : 0, (const unsigned*)nullptr, count, begin);
}

size_t iterate_smem_bytes(const TilePlanDev& plan)
{
    // 2 x sp[maxLocals] + slot rows (halo endpoints have no slot: they are not stored) + bend records + stretch records
    // + a ring of three tile descriptors
    const size_t rows = (size_t)(plan.maxKS > plan.maxKB ? plan.maxKS : plan.maxKB);
    return sizeof(float4) * (2 * (size_t)plan.maxLocals + rows * plan.threads + plan.maxBendPerTile + ((size_t)plan.maxStretchPerTile + 1) / 2) +
           3 * sizeof(TileDesc) + 16;  // + two mbarriers
}

// kernel variants: (slot-row width, threads)
#define VT_ITERATE_VARIANTS(X) X(7, 128) X(7, 160) X(8, 256) X(8, 320) X(9, 512) X(9, 640)

void launch_iterate(const FusedLaunch& L, const float4* predIn, float4* predOut, const TilePlanDev& plan,
                    const float* attachSlotPositions, const FrameParams* fp, Instancing inst)
{
    const unsigned total = plan.numTiles * inst.count;
    if (!total) return;
    const unsigned grid = total < plan.residentCtas ? total : plan.residentCtas;  // persistent: one wave
    const size_t smem = iterate_smem_bytes(plan);
#define VT_LAUNCH(LOG2T, NT)                                                                                               \
    if (plan.threads == (1u << LOG2T) && plan.ctaThreads == NT) {                                                          \
        launch_pdl(iterate_tile_kernel<LOG2T, NT>, dim3(grid), dim3(NT), smem, L.stream, predIn, predOut, plan, attachSlotPositions, fp, inst, total); \
        return;                                                                                                            \
    }
    VT_ITERATE_VARIANTS(VT_LAUNCH)
#undef VT_LAUNCH
    throw Error(VELVET_ERR_INVALID_ARGUMENT, "unsupported Jacobi tile size");
}

// Opts in to > 48 KB dynamic shared memory and returns how many CTAs of this plan the current device keeps resident.
unsigned configure_iterate_kernel(size_t smemBytes, unsigned threads, unsigned ctaThreads)
{
    int dev = 0, sms = 0, perSm = 0;
    VT_CUDA(cudaGetDevice(&dev));
    VT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
#define VT_CONFIG(LOG2T, NT)                                                                                               \
    if (threads == (1u << LOG2T) && ctaThreads == NT) {                                                                    \
        VT_CUDA(cudaFuncSetAttribute(iterate_tile_kernel<LOG2T, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes)); \
        VT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, iterate_tile_kernel<LOG2T, NT>, NT, smemBytes));     \
    }
    VT_ITERATE_VARIANTS(VT_CONFIG)
#undef VT_CONFIG
    if (perSm < 1) throw Error(VELVET_ERR_UNSUPPORTED, "the Jacobi tile kernel does not fit on an SM");
    return (unsigned)(sms * perSm);
}

void launch_iterate_grid(const FusedLaunch& L, float4* predIn, float4* predOut, const GridPlanDev& plan,
                         const float* attachSlotPositions, const FrameParams* fp, Instancing inst, const ddpeer::StripArgs* strip,
                         unsigned iterations, unsigned* gridBarrier)
{
    ddpeer::StripArgs a{};
    unsigned total = plan.numTiles * inst.count;
    if (strip) {  // one strip of a decomposed cloth: the owned tile rows of the (single) cloth
        a = *strip;
        a.enabled = 1;
        total = (a.tileRowEnd - a.tileRowBegin) * plan.tilesY0;
    }
    if (!total || !iterations) return;
    if (iterations > 1 && (strip || !gridBarrier)) throw Error(VELVET_ERR_STATE, "iterate_grid: several iterations per launch need a barrier word and no strip");
    const unsigned grid = total < plan.residentCtas ? total : plan.residentCtas;  // persistent: one wave, all CTAs resident
    if (iterations > 1) VT_CUDA(cudaMemsetAsync(gridBarrier, 0, sizeof(unsigned), L.stream));
    const bool coop = iterations > 1;  // grid-wide barriers inside: every CTA must be resident
    if (plan.tileX == (unsigned)GRID_TILE_RX && plan.tileY == (unsigned)GRID_TILE_RY && !strip)
        launch_kernel_ex(coop, iterate_grid_kernel<GRID_TILE_RX + 1, GRID_TILE_RY + 1>, dim3(grid), dim3(256), GRID_SMEM_BYTES, L.stream,
                         predIn, predOut, plan, attachSlotPositions, fp, inst, total, a, iterations, gridBarrier);
    else if (plan.tileX == (unsigned)GRID_TILE && plan.tileY == (unsigned)GRID_TILE)
        launch_kernel_ex(coop, iterate_grid_kernel<GRID_B, GRID_B>, dim3(grid), dim3(256), GRID_SMEM_BYTES, L.stream, predIn, predOut, plan,
                         attachSlotPositions, fp, inst, total, a, iterations, gridBarrier);
    else
        throw Error(VELVET_ERR_STATE, "iterate_grid: no kernel for this tile shape");
}

unsigned configure_iterate_grid_kernel()
{
    int dev = 0, sms = 0, perSm = 0;
    VT_CUDA(cudaGetDevice(&dev));
    VT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int perSmRect = 0;
    VT_CUDA(cudaFuncSetAttribute(iterate_grid_kernel<GRID_B, GRID_B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GRID_SMEM_BYTES));
    VT_CUDA(cudaFuncSetAttribute(iterate_grid_kernel<GRID_TILE_RX + 1, GRID_TILE_RY + 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)GRID_SMEM_BYTES));
    VT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, iterate_grid_kernel<GRID_B, GRID_B>, 256, GRID_SMEM_BYTES));
    VT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSmRect, iterate_grid_kernel<GRID_TILE_RX + 1, GRID_TILE_RY + 1>, 256,
                                                          GRID_SMEM_BYTES));
    if (perSmRect < perSm) perSm = perSmRect;
    if (perSm < 1) throw Error(VELVET_ERR_UNSUPPORTED, "the grid Jacobi kernel does not fit on an SM");
    return (unsigned)(sms * perSm);
}

void launch_end_substep(const FusedLaunch& L, const float4* predIn, float4* pos4, float4* predNext, bool last,
                        float* positionsOut, float* velocitiesOut, float* predictedOut, const FrameParams* fp)
{
    launch_pdl(end_substep_kernel, dim3(pgrid(L.numParticles)), dim3(PB), 0, L.stream, predIn, pos4, predNext, last ? 1 : 0, positionsOut,
               velocitiesOut, predictedOut, fp, L.numParticles);
}

void launch_normals(const FusedLaunch& L, const float4* pos4, const unsigned* indices, const unsigned* vtxTriOff,
                    const unsigned* vtxTris, float* normalsOut, Instancing inst)
{
    launch_pdl(normals_kernel, dim3(pgrid(inst.particles), inst.count), dim3(PB), 0, L.stream, pos4, indices, vtxTriOff, vtxTris, normalsOut,
               inst.particles);
}

#if !VT_FAST_MATH  // the spatial hash is integer work: one (exact) build only
void launch_hash_particles(const FusedLaunch& L, unsigned* keys, unsigned* vals, const float4* pred, float cellSpacing,
                           int tableSizePerInstance, Instancing inst, unsigned* cellStart, int tableSize)
{
    hash_particles_kernel<PosFloat4><<<pgrid(L.numParticles), PB, 0, L.stream>>>(keys, vals, PosFloat4{pred}, L.numParticles,
                                                                                 cellSpacing, tableSizePerInstance, inst.particles,
                                                                                 cellStart, (unsigned)tableSize);
}

// cellStart was filled with 0xffffffff by launch_hash_particles of the same rebuild
void launch_find_cell_start(const FusedLaunch& L, unsigned* cellStart, unsigned* cellEnd, const unsigned* particleHash)
{
    find_cell_start_kernel<<<pgrid(L.numParticles), PB, 0, L.stream>>>(cellStart, cellEnd, particleHash, L.numParticles);
}

void launch_cache_neighbors(const FusedLaunch& L, unsigned* neighbors, const unsigned* particleIndex,
                            const unsigned* cellStart, const unsigned* cellEnd, const float4* pred, const float4* init4,
                            VtHashParams hp)
{
    cache_neighbors_kernel<PosFloat4, PosFloat4><<<pgrid(L.numParticles), PB, 0, L.stream>>>(
        neighbors, particleIndex, cellStart, cellEnd, PosFloat4{pred}, PosFloat4{init4}, hp);
}

int launch_cache_neighbors_sorted(const FusedLaunch& L, unsigned* neighbors, const unsigned* particleIndex,
                                  const unsigned* cellStart, const unsigned* cellEnd, const float4* pred,
                                  const float4* init4, float4* sortedScratch, VtHashParams hp, Instancing inst,
                                  const unsigned char* ownedMask, unsigned numOwned, const unsigned* sortedHashForCells,
                                  unsigned ownedBegin, unsigned bandParticles)
{
    if (hp.tableSize <= 0) return 0;
    const unsigned n = L.numParticles;
    SortedParticle* sorted = reinterpret_cast<SortedParticle*>(sortedScratch);
    // (with sortedHashForCells the cell table is built in the same launch: cellStart was cleared by launch_hash_particles)
    reorder_sorted_kernel<<<pgrid(n), PB, 0, L.stream>>>(sorted, particleIndex, pred, init4, n, hp.cellSpacing,
                                                         const_cast<unsigned*>(cellStart), const_cast<unsigned*>(cellEnd), sortedHashForCells);
    // bucket keys through shared memory up to VT_WALK_SMEM_KEYS_MAX particles (VELVET_WALK_KEYS=smem|regs overrides: A/B runs)
    static const int keysMode = [] {
        const char* e = getenv("VELVET_WALK_KEYS");
        return e && !strcmp(e, "smem") ? 1 : e && !strcmp(e, "regs") ? 2 : 0;
    }();
    // (batched instances sort instance by instance: what counts for the locality of the walk is the size of one instance)
    const bool smemKeys = keysMode == 1 || (keysMode == 0 && (inst.count > 1 ? inst.particles : n) <= VT_WALK_SMEM_KEYS_MAX);
    const FastMod fm = make_fastmod((unsigned)hp.tableSize);
    auto walk = [&](unsigned threads, const unsigned* slots) {
        const unsigned grid = (threads + CN_THREADS - 1) / CN_THREADS;
        if (smemKeys)
            cache_neighbors_sorted_kernel<true><<<grid, CN_THREADS, 0, L.stream>>>(neighbors, cellStart, cellEnd, sorted, hp, fm, inst.particles,
                                                                                   slots, threads);
        else
            cache_neighbors_sorted_kernel<false><<<grid, CN_THREADS, 0, L.stream>>>(neighbors, cellStart, cellEnd, sorted, hp, fm,
                                                                                    inst.particles, slots, threads);
    };
    // scratch behind the sorted records: CN_MAX_BANDS counters + up to n slot indices
    unsigned* counter = reinterpret_cast<unsigned*>(sorted + n);
    unsigned* slots = counter + CN_MAX_BANDS;
    // a contiguous index range of a large cloth (all of it, or the strip a rank owns): walked band by band (hash_kernels.cuh)
    if (const char* e = getenv("VELVET_WALK_BAND")) bandParticles = (unsigned)atoi(e);  // particles per band, 0 = never (tests, A/B runs)
    const bool contiguous = !ownedMask || ownedBegin != 0xffffffffu;
    if (contiguous && bandParticles) {
        const unsigned begin = ownedMask ? ownedBegin : 0u, count = ownedMask ? numOwned : n;
        const unsigned numBands = (count + bandParticles - 1) / bandParticles;
        if (count && numBands <= CN_MAX_BANDS) {
            VT_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned) * CN_MAX_BANDS, L.stream));
            band_slots_kernel<<<pgrid(n), PB, 0, L.stream>>>(sorted, n, begin, count, bandParticles, numBands, slots, counter);
            walk(count, slots);
            return 4;
        }
    }
    if (!ownedMask) {
        walk(n, nullptr);
        return 2;
    }
    // decomposed mode: compact the owned slots
    VT_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned), L.stream));
    compact_owned_slots_kernel<<<pgrid(n), PB, 0, L.stream>>>(sorted, ownedMask, n, slots, counter);
    if (numOwned) walk(numOwned, slots);
    return 4;
}

size_t cache_neighbors_scratch_float4(size_t n) { return 2 * n + (n + CN_MAX_BANDS + 3) / 4 + 1; }  // sorted records + counters + slots

// plain device-side copy (read-back staging): a kernel rather than cudaMemcpyAsync because the source is managed memory,
// for which the driver's copy path is not stream-fast
static __global__ void __launch_bounds__(256) copy_words_kernel(const unsigned* __restrict__ src, unsigned* __restrict__ dst, size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
}
void launch_copy_words(cudaStream_t stream, const void* src, void* dst, size_t words)
{
    if (!words) return;
    const unsigned grid = (unsigned)std::min<size_t>((words + 255) / 256, 148 * 16);
    copy_words_kernel<<<grid, 256, 0, stream>>>(static_cast<const unsigned*>(src), static_cast<unsigned*>(dst), words);
}

void launch_pack_float4(const FusedLaunch& L, const float* packed3, float4* out, unsigned n)
{
    if (n) pack_float4_kernel<<<pgrid(n), PB, 0, L.stream>>>(packed3, out, n);
}

void launch_unpack_float4(const FusedLaunch& L, const float4* in, float* packed3, unsigned n)
{
    if (n) unpack_float4_kernel<<<pgrid(n), PB, 0, L.stream>>>(in, packed3, n);
}

void launch_gather_by_id(const FusedLaunch& L, const float4* src, const unsigned* ids, unsigned n, float4* out)
{
    if (n) gather_by_id_kernel<<<pgrid(n), PB, 0, L.stream>>>(src, ids, n, out);
}

void launch_scatter_by_id(const FusedLaunch& L, const float4* in, const unsigned* ids, unsigned n, float4* dst)
{
    if (n) scatter_by_id_kernel<<<pgrid(n), PB, 0, L.stream>>>(in, ids, n, dst);
}
#endif

}  // namespace VT_MATH_NS
}  // namespace velvet
