// seam_kernels.cu -- the kernel seam: one sm_100a kernel + C entry point per free function the reference's
// host code calls (VtClothSolverGPU.cuh L99-168, SpatialHashGPU.cuh L17-25).  Same buffers, same layouts
// (packed float3), same per-stage semantics, so a maintainer can swap the reference's .cu files for this
// library without touching VtClothSolverGPU.hpp.  These are the compatibility kernels; the fused pipeline
// behind velvet_solver_simulate lives in fused_kernels.cu.
//
// Differences from the reference kernels: parameters travel as a __grid_constant__ kernel argument instead
// of a cudaMemcpyToSymbol per frame; colliders are staged once per block in shared memory; no per-call
// cudaEvent timers; launches go to a selectable stream; errors are returned, never exit().
#include <mutex>

#include "capi_util.hpp"
#include "hash_kernels.cuh"
#include "radix_sort.cuh"
#include <map>

#include "seam.hpp"
#include "vt_math.cuh"

namespace velvet {

namespace {

VtSimParams g_params;  // host copy, like h_params in VtClothSolverGPU.cu L11
bool g_paramsSet = false;
cudaStream_t g_stream = 0;
std::mutex g_mutex;

// Scratch for HashObjects / SortPairs (the reference keeps a function-static VtBuffer, SpatialHashGPU.cu L140).  One set per
// DEVICE (a pointer allocated on device 0 must never be handed to a kernel on device 1), grow-only, and handed from one
// user to the next through an event: callers run on their own non-blocking streams, so the next sort waits (on the device)
// for the previous one instead of overwriting scratch that is still being read.  g_mutex covers the bookkeeping only.
struct SortScratch {
    RadixSorter sorter;
    DeviceBuffer<unsigned> keysAlt, valsAlt;
    cudaEvent_t lastUse = nullptr;
};
std::map<int, SortScratch> g_scratch;

// g_mutex must be held.  Orders `st` after the previous user of this device's scratch and makes room for n items.
SortScratch& acquire_scratch(unsigned n, cudaStream_t st)
{
    int dev = 0;
    VT_CUDA(cudaGetDevice(&dev));
    SortScratch& s = g_scratch[dev];
    if (s.lastUse && (n > s.keysAlt.size() || n > s.valsAlt.size())) VT_CUDA(cudaEventSynchronize(s.lastUse));  // about to reallocate
    s.keysAlt.reserve(n);
    s.valsAlt.reserve(n);
    s.sorter.reserve(n);
    if (!s.lastUse) VT_CUDA(cudaEventCreateWithFlags(&s.lastUse, cudaEventDisableTiming));
    else VT_CUDA(cudaStreamWaitEvent(st, s.lastUse, 0));
    return s;
}
void release_scratch(SortScratch& s, cudaStream_t st) { VT_CUDA(cudaEventRecord(s.lastUse, st)); }

constexpr int BLOCK = 256;  // Common.cuh L46
inline unsigned grid_for(unsigned n) { return (n + BLOCK - 1) / BLOCK; }

__device__ __forceinline__ void atomic_add3(float* base, unsigned idx, vec3 v, int reorder)
{
    // AtomicAdd, VtClothSolverGPU.cu L13-21: component order rotated by `reorder` to spread contention.
    float* p = base + 3 * (size_t)idx;
    const float c[3] = {v.x, v.y, v.z};
    const int r1 = reorder % 3, r2 = (reorder + 1) % 3, r3 = (reorder + 2) % 3;
    atomicAdd(p + r1, c[r1]);
    atomicAdd(p + r2, c[r2]);
    atomicAdd(p + r3, c[r3]);
}

__global__ void __launch_bounds__(BLOCK) initialize_positions_kernel(float* positions, int start, int count, mat4 model)
{
    const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (unsigned)count) return;
    store3(positions, (size_t)start + id, mul_point(model.m, load3(positions, (size_t)start + id), 1.0f));
}

__global__ void __launch_bounds__(BLOCK) predict_positions_kernel(float* predicted, float* velocities,
                                                                  const float* positions, float dt,
                                                                  const __grid_constant__ VtSimParams P)
{
    const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= P.numParticles) return;
    vec3 v = load3(velocities, id) + V3(P.gravity[0], P.gravity[1], P.gravity[2]) * dt;
    store3(velocities, id, v);
    store3(predicted, id, load3(positions, id) + v * dt);
}

__global__ void __launch_bounds__(BLOCK) solve_stretch_kernel(const float* predicted, float* deltas, int* deltaCounts,
                                                              const int* __restrict__ stretchIndices,
                                                              const float* __restrict__ stretchLengths,
                                                              const float* __restrict__ invMasses, unsigned n)
{
    const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n) return;
    const int idx1 = stretchIndices[2 * id], idx2 = stretchIndices[2 * id + 1];
    vec3 c1, c2;
    if (stretch_eval(load3(predicted, idx1), load3(predicted, idx2), invMasses[idx1], invMasses[idx2],
                     stretchLengths[id], c1, c2)) {
        const int reorder = idx1 + idx2;
        atomic_add3(deltas, idx1, c1, reorder);
        atomic_add3(deltas, idx2, c2, reorder);
        atomicAdd(&deltaCounts[idx1], 1);
        atomicAdd(&deltaCounts[idx2], 1);
    }
}

__global__ void __launch_bounds__(BLOCK) solve_bending_kernel(const float* predicted, float* deltas, int* deltaCounts,
                                                              const unsigned* __restrict__ bendIndices,
                                                              const float* __restrict__ restAngles,
                                                              const float* __restrict__ invMass, unsigned n, float dt,
                                                              float bendCompliance)
{
    const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n) return;
    const unsigned i0 = bendIndices[4 * id], i1 = bendIndices[4 * id + 1], i2 = bendIndices[4 * id + 2],
                   i3 = bendIndices[4 * id + 3];
    vec3 c0, c1, c2, c3;
    const float xpbd_bend = bendCompliance / dt / dt;
    if (!bend_eval(load3(predicted, i0), load3(predicted, i1), load3(predicted, i2), load3(predicted, i3), invMass[i0],
                   invMass[i1], invMass[i2], invMass[i3], restAngles[id], xpbd_bend, c0, c1, c2, c3))
        return;
    const int reorder = (int)(i0 + i1 + i2 + i3);
    atomic_add3(deltas, i0, c0, reorder);
    atomic_add3(deltas, i1, c1, reorder);
    atomic_add3(deltas, i2, c2, reorder);
    atomic_add3(deltas, i3, c3, reorder);
    atomicAdd(&deltaCounts[i0], 1);
    atomicAdd(&deltaCounts[i1], 1);
    atomicAdd(&deltaCounts[i2], 1);
    atomicAdd(&deltaCounts[i3], 1);
}

__global__ void __launch_bounds__(BLOCK) solve_attachment_kernel(const float* predicted, float* deltas, int* deltaCounts,
                                                                 const float* __restrict__ invMass,
                                                                 const int* __restrict__ attachParticleIDs,
                                                                 const int* __restrict__ attachSlotIDs,
                                                                 const float* __restrict__ attachSlotPositions,
                                                                 const float* __restrict__ attachDistances, int n,
                                                                 float longRangeStretchiness)
{
    const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (unsigned)n) return;
    const unsigned pid = (unsigned)attachParticleIDs[id];
    vec3 corr;
    if (attach_eval(load3(predicted, pid), invMass[pid], load3(attachSlotPositions, (size_t)attachSlotIDs[id]),
                    attachDistances[id], longRangeStretchiness, corr)) {
        atomic_add3(deltas, pid, corr, (int)id);
        atomicAdd(&deltaCounts[pid], 1);
    }
}

__global__ void __launch_bounds__(BLOCK) apply_deltas_kernel(float* predicted, float* deltas, int* deltaCounts,
                                                             unsigned numParticles, float relaxationFactor)
{
    const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= numParticles) return;
    const float count = (float)deltaCounts[id];
    if (count > 0) {
        store3(predicted, id, load3(predicted, id) + load3(deltas, id) / count * relaxationFactor);
        store3(deltas, id, V3(0, 0, 0));
        deltaCounts[id] = 0;
    }
}

__global__ void __launch_bounds__(BLOCK) collide_sdf_kernel(float* predicted, const VtSDFCollider* __restrict__ colliders,
                                                            const float* positions, unsigned numColliders, float dt,
                                                            const __grid_constant__ VtSimParams P)
{
    extern __shared__ unsigned char s_raw[];
    PreparedCollider* s_col = reinterpret_cast<PreparedCollider*>(s_raw);
    for (unsigned i = threadIdx.x; i < numColliders; i += blockDim.x) prepare_collider(colliders[i], s_col[i]);
    __syncthreads();
    const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= P.numParticles) return;
    const vec3 pos = load3(positions, id);  // read before the store: predicted may alias positions
    const vec3 pred = load3(predicted, id);
    store3(predicted, id, collide_sdf_point(s_col, numColliders, pred, pos, P.collisionMargin, P.friction, dt));
}

__global__ void __launch_bounds__(BLOCK) collide_particles_kernel(float* deltas, int* deltaCounts, const float* predicted,
                                                                  const float* __restrict__ invMasses,
                                                                  const unsigned* __restrict__ neighbors,
                                                                  const float* __restrict__ positions,
                                                                  const __grid_constant__ VtSimParams P)
{
    const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned N = P.numParticles;
    if (id >= N) return;
    vec3 positionDelta = V3(0, 0, 0);
    int deltaCount = 0;
    const vec3 pred_i = load3(predicted, id);
    const vec3 vel_i = pred_i - load3(positions, id);
    const float w_i = invMasses[id];
    const unsigned long long limit = (unsigned long long)N * (unsigned)P.maxNumNeighbors;
    for (unsigned long long nb = id; nb < limit; nb += N) {
        const unsigned j = neighbors[nb];
        if (j > N) break;
        const float w_j = invMasses[j];
        const float denom = w_i + w_j;
        if (denom <= 0) continue;
        const vec3 pred_j = load3(predicted, j);
        const vec3 diff = pred_i - pred_j;
        const float distance = length(diff);
        if (distance >= P.particleDiameter) continue;
        const vec3 gradient = diff / (distance + VT_EPSILON);
        const float lambda = vt_div(distance - P.particleDiameter, denom);
        const vec3 common = lambda * gradient;
        deltaCount++;
        positionDelta -= w_i * common;
        const vec3 relativeVelocity = vel_i - (pred_j - load3(positions, j));
        positionDelta += w_i * compute_friction(P.friction, common, relativeVelocity);
    }
    store3(deltas, id, positionDelta);
    deltaCounts[id] = deltaCount;
}

__global__ void __launch_bounds__(BLOCK) finalize_kernel(float* velocities, float* positions, const float* predicted,
                                                         float dt, const __grid_constant__ VtSimParams P)
{
    const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= P.numParticles) return;
    vec3 newPos, vel;
    finalize_point(load3(predicted, id), load3(positions, id), dt, P.maxSpeed, P.damping, newPos, vel);
    store3(velocities, id, vel);
    store3(positions, id, newPos);
}

__global__ void __launch_bounds__(BLOCK) triangle_normals_kernel(float* normals, const float* __restrict__ positions,
                                                                 const unsigned* __restrict__ indices, unsigned numTriangles)
{
    const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= numTriangles) return;
    const unsigned a = indices[3 * id], b = indices[3 * id + 1], c = indices[3 * id + 2];
    const vec3 p1 = load3(positions, a), p2 = load3(positions, b), p3 = load3(positions, c);
    const vec3 normal = cross(p2 - p1, p3 - p1);
    const int reorder = (int)(a + b + c);
    atomic_add3(normals, a, normal, reorder);
    atomic_add3(normals, b, normal, reorder);
    atomic_add3(normals, c, normal, reorder);
}

__global__ void __launch_bounds__(BLOCK) vertex_normals_kernel(float* normals, unsigned numParticles)
{
    const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= numParticles) return;
    store3(normals, id, normalize(load3(normals, id)));
}

thread_local int t_lastLaunches = 0;

int need_params()
{
    if (!g_paramsSet) return set_error(VELVET_ERR_STATE, "velvet_SetSimulationParams must be called first");
    return VELVET_OK;
}

}  // namespace

// ------------------------------------------------------------------ C++ launchers (seam.hpp)
namespace seam {

int LastLaunchCount() { return t_lastLaunches; }

void InitializePositions(float* positions, int start, int count, const float* modelMatrix16, cudaStream_t st)
{
    t_lastLaunches = 0;
    if (count <= 0) return;  // CUDA_CALL early-out, Common.cuh L24-25
    mat4 m;
    for (int i = 0; i < 16; i++) m.m[i] = modelMatrix16[i];
    initialize_positions_kernel<<<grid_for((unsigned)count), BLOCK, 0, st>>>(positions, start, count, m);
    t_lastLaunches = 1;
    VT_CUDA(cudaGetLastError());
}

void PredictPositions(const VtSimParams& P, float* predicted, float* velocities, const float* positions, float dt,
                      cudaStream_t st)
{
    t_lastLaunches = 0;
    if (P.numParticles == 0) return;
    predict_positions_kernel<<<grid_for(P.numParticles), BLOCK, 0, st>>>(predicted, velocities, positions, dt, P);
    t_lastLaunches = 1;
    VT_CUDA(cudaGetLastError());
}

void SolveStretch(float* predicted, float* deltas, int* deltaCounts, const int* stretchIndices,
                  const float* stretchLengths, const float* invMasses, unsigned n, cudaStream_t st)
{
    t_lastLaunches = 0;
    if (n == 0) return;
    solve_stretch_kernel<<<grid_for(n), BLOCK, 0, st>>>(predicted, deltas, deltaCounts, stretchIndices, stretchLengths,
                                                        invMasses, n);
    t_lastLaunches = 1;
    VT_CUDA(cudaGetLastError());
}

void SolveBending(const VtSimParams& P, float* predicted, float* deltas, int* deltaCounts, const unsigned* bendIndices,
                  const float* bendAngles, const float* invMass, unsigned n, float dt, cudaStream_t st)
{
    t_lastLaunches = 0;
    if (n == 0) return;
    solve_bending_kernel<<<grid_for(n), BLOCK, 0, st>>>(predicted, deltas, deltaCounts, bendIndices, bendAngles, invMass, n,
                                                        dt, P.bendCompliance);
    t_lastLaunches = 1;
    VT_CUDA(cudaGetLastError());
}

void SolveAttachment(const VtSimParams& P, float* predicted, float* deltas, int* deltaCounts, const float* invMass,
                     const int* attachParticleIDs, const int* attachSlotIDs, const float* attachSlotPositions,
                     const float* attachDistances, int n, cudaStream_t st)
{
    t_lastLaunches = 0;
    if (n <= 0) return;
    solve_attachment_kernel<<<grid_for((unsigned)n), BLOCK, 0, st>>>(predicted, deltas, deltaCounts, invMass,
                                                                     attachParticleIDs, attachSlotIDs, attachSlotPositions,
                                                                     attachDistances, n, P.longRangeStretchiness);
    t_lastLaunches = 1;
    VT_CUDA(cudaGetLastError());
}

void ApplyDeltas(const VtSimParams& P, float* predicted, float* deltas, int* deltaCounts, cudaStream_t st)
{
    t_lastLaunches = 0;
    if (P.numParticles == 0) return;
    apply_deltas_kernel<<<grid_for(P.numParticles), BLOCK, 0, st>>>(predicted, deltas, deltaCounts, P.numParticles,
                                                                    P.relaxationFactor);
    t_lastLaunches = 1;
    VT_CUDA(cudaGetLastError());
}

void CollideSDF(const VtSimParams& P, float* predicted, const VtSDFCollider* colliders, const float* positions,
                unsigned numColliders, float dt, cudaStream_t st)
{
    t_lastLaunches = 0;
    if (numColliders == 0 || P.numParticles == 0) return;  // VtClothSolverGPU.cu L324
    const size_t smem = sizeof(PreparedCollider) * numColliders;
    if (smem > 48 * 1024) throw Error(VELVET_ERR_UNSUPPORTED, "CollideSDF: more than 250 colliders");
    collide_sdf_kernel<<<grid_for(P.numParticles), BLOCK, smem, st>>>(predicted, colliders, positions, numColliders, dt, P);
    t_lastLaunches = 1;
    VT_CUDA(cudaGetLastError());
}

void CollideParticles(const VtSimParams& P, float* deltas, int* deltaCounts, float* predicted, const float* invMasses,
                      const unsigned* neighbors, const float* positions, cudaStream_t st)
{
    t_lastLaunches = 0;
    const unsigned n = P.numParticles;
    if (n == 0) return;
    collide_particles_kernel<<<grid_for(n), BLOCK, 0, st>>>(deltas, deltaCounts, predicted, invMasses, neighbors, positions, P);
    apply_deltas_kernel<<<grid_for(n), BLOCK, 0, st>>>(predicted, deltas, deltaCounts, n, P.relaxationFactor);
    t_lastLaunches = 2;
    VT_CUDA(cudaGetLastError());
}

void Finalize(const VtSimParams& P, float* velocities, float* positions, const float* predicted, float dt, cudaStream_t st)
{
    t_lastLaunches = 0;
    if (P.numParticles == 0) return;
    finalize_kernel<<<grid_for(P.numParticles), BLOCK, 0, st>>>(velocities, positions, predicted, dt, P);
    t_lastLaunches = 1;
    VT_CUDA(cudaGetLastError());
}

void ComputeNormal(const VtSimParams& P, float* normals, const float* positions, const unsigned* indices,
                   unsigned numTriangles, cudaStream_t st)
{
    t_lastLaunches = 0;
    const unsigned n = P.numParticles;
    if (n == 0) return;  // VtClothSolverGPU.cu L459
    VT_CUDA(cudaMemsetAsync(normals, 0, sizeof(float) * 3 * (size_t)n, st));
    if (numTriangles) {
        triangle_normals_kernel<<<grid_for(numTriangles), BLOCK, 0, st>>>(normals, positions, indices, numTriangles);
        t_lastLaunches++;
    }
    vertex_normals_kernel<<<grid_for(n), BLOCK, 0, st>>>(normals, n);
    t_lastLaunches++;
    VT_CUDA(cudaGetLastError());
}

void HashObjects(unsigned* particleHash, unsigned* particleIndex, unsigned* cellStart, unsigned* cellEnd,
                 unsigned* neighbors, const float* positions, const float* originalPositions, VtHashParams hp,
                 cudaStream_t st)
{
    t_lastLaunches = 0;
    const unsigned n = hp.numObjects;
    if (n == 0) return;
    if (hp.tableSize <= 0) throw Error(VELVET_ERR_INVALID_ARGUMENT, "HashObjects: tableSize <= 0");
    std::lock_guard<std::mutex> lk(g_mutex);  // shared sort scratch (function-static in the reference, .cu L140)
    const int maxBit = (int)ceil(log2((double)hp.tableSize));  // SpatialHashGPU.cu L179
    SortScratch& S = acquire_scratch(n, st);
    // With an odd number of digit passes the keys are produced in the alternate buffer so that the sorted
    // result lands in the caller's particleHash / particleIndex (in-place semantics of the reference).
    const bool odd = RadixSorter::numPasses(maxBit) & 1;
    unsigned* k0 = odd ? S.keysAlt.data() : particleHash;
    unsigned* v0 = odd ? S.valsAlt.data() : particleIndex;
    unsigned* k1 = odd ? particleHash : S.keysAlt.data();
    unsigned* v1 = odd ? particleIndex : S.valsAlt.data();
    hash_particles_kernel<PosPacked3><<<grid_for(n), BLOCK, 0, st>>>(k0, v0, PosPacked3{positions}, n, hp.cellSpacing, hp.tableSize, n);
    S.sorter.sort(k0, v0, k1, v1, n, maxBit, st);
    release_scratch(S, st);
    VT_CUDA(cudaMemsetAsync(cellStart, 0xff, sizeof(unsigned) * (size_t)hp.tableSize, st));
    find_cell_start_kernel<<<grid_for(n), BLOCK, 0, st>>>(cellStart, cellEnd, particleHash, n);
    cache_neighbors_kernel<PosPacked3, PosPacked3><<<grid_for(n), BLOCK, 0, st>>>(
        neighbors, particleIndex, cellStart, cellEnd, PosPacked3{positions}, PosPacked3{originalPositions}, hp);
    t_lastLaunches = 3 + S.sorter.lastLaunchCount();
    VT_CUDA(cudaGetLastError());
}

}  // namespace seam
}  // namespace velvet

using namespace velvet;

// ------------------------------------------------------------------ C ABI
#define VT_REQUIRE(cond, msg) \
    if (!(cond)) return set_error(VELVET_ERR_INVALID_ARGUMENT, msg)


// ---- self-test of vt_div / vec3 operator/ (vt_math.cuh): quotients as raw bits, next to the compiler's own division
namespace velvet { namespace {
__global__ void selftest_division_kernel(const float* __restrict__ x, const float* __restrict__ y, unsigned n, unsigned* __restrict__ outDiv,
                                         unsigned* __restrict__ outVec, unsigned* __restrict__ outPlain)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float a = x[i], b = x[(i + 1) % n], c = x[(i + 2) % n], d = y[i];
    outDiv[i] = __float_as_uint(vt_div(a, d));
    const vec3 v = V3(a, b, c) / d;
    outVec[3 * (size_t)i + 0] = __float_as_uint(v.x);
    outVec[3 * (size_t)i + 1] = __float_as_uint(v.y);
    outVec[3 * (size_t)i + 2] = __float_as_uint(v.z);
    outPlain[i] = __float_as_uint(__fdiv_rn(a, d));
}

// ---- self-test of the checked-fast constraint evaluators (vt_math.cuh): whenever their validity predicate holds they must
// return exactly what the branchy evaluators return; whenever it does not, the caller falls back to those, so a mismatch is
// only counted while `ok` is true.  mismatches[0] stretch, [1] bend, [2] sqrt; fast[0..2] = how often the fast path was valid.
__device__ __forceinline__ bool same_bits(float a, float b) { return __float_as_uint(a) == __float_as_uint(b) || (a != a && b != b); }
__device__ __forceinline__ bool same_vec(vec3 a, vec3 b) { return same_bits(a.x, b.x) && same_bits(a.y, b.y) && same_bits(a.z, b.z); }

__global__ void selftest_constraints_kernel(const float* __restrict__ in, unsigned n, unsigned long long* __restrict__ mismatches,
                                            unsigned long long* __restrict__ fast)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* f = in + 18 * (size_t)i;
    const vec3 p0 = V3(f[0], f[1], f[2]), p1 = V3(f[3], f[4], f[5]), p2 = V3(f[6], f[7], f[8]), p3 = V3(f[9], f[10], f[11]);
    const float w0 = f[12], w1 = f[13], w2 = f[14], w3 = f[15], rest = f[16], xpbd = f[17];
    {
        vec3 a1 = V3(0, 0, 0), a2 = a1, b1, b2;
        const bool actA = stretch_eval(p0, p1, w0, w1, rest, a1, a2);
        bool ok = true;
        const bool actB = stretch_eval_u(p0, p1, w0, w1, rest, b1, b2, ok);
        if (ok) {
            atomicAdd(fast + 0, 1ull);
            if (actA != actB || (actA && !(same_vec(a1, b1) && same_vec(a2, b2)))) atomicAdd(mismatches + 0, 1ull);
        }
    }
    {
        vec3 a0 = V3(0, 0, 0), a1 = a0, a2 = a0, a3 = a0, b0, b1, b2, b3;
        const bool actA = bend_eval(p0, p1, p2, p3, w0, w1, w2, w3, rest, xpbd, a0, a1, a2, a3);
        bool ok = true;
        const bool actB = bend_eval_u(p0, p1, p2, p3, w0, w1, w2, w3, rest, xpbd, b0, b1, b2, b3, ok);
        if (ok) {
            atomicAdd(fast + 1, 1ull);
            if (actA != actB || (actA && !(same_vec(a0, b0) && same_vec(a1, b1) && same_vec(a2, b2) && same_vec(a3, b3))))
                atomicAdd(mismatches + 1, 1ull);
        }
    }
}

// every one of the 2^32 operands of sqrt
__global__ void selftest_sqrt_kernel(unsigned long long* __restrict__ mismatches, unsigned long long* __restrict__ fast)
{
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long bad = 0, valid = 0;
    for (unsigned long long u = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; u < (1ull << 32); u += stride) {
        const float x = __uint_as_float((unsigned)u);
        bool ok = true;
        const float s = vt_sqrt_u(x, ok);
        if (ok) {
            valid++;
            if (!same_bits(s, sqrtf(x))) bad++;
        }
    }
    if (bad) atomicAdd(mismatches + 2, bad);
    atomicAdd(fast + 2, valid);
}
} }  // namespace

extern "C" {

int velvet_SetSimulationParams(const VtSimParams* hostParams)
{
    VT_REQUIRE(hostParams, "SetSimulationParams: hostParams is NULL");
    std::lock_guard<std::mutex> lk(g_mutex);
    g_params = *hostParams;
    g_paramsSet = true;
    return VELVET_OK;
}

int velvet_seam_set_stream(void* cudaStream)
{
    g_stream = (cudaStream_t)cudaStream;
    return VELVET_OK;
}

int velvet_InitializePositions(float* positions, int start, int count, const float* modelMatrix16)
{
    VT_API_BEGIN
    if (count == 0) return VELVET_OK;
    VT_REQUIRE(positions && modelMatrix16 && count > 0 && start >= 0, "InitializePositions: bad argument");
    seam::InitializePositions(positions, start, count, modelMatrix16, g_stream);
    VT_API_END
}

int velvet_PredictPositions(float* predicted, float* velocities, const float* positions, float deltaTime)
{
    VT_API_BEGIN
    if (int e = need_params()) return e;
    if (g_params.numParticles == 0) return VELVET_OK;
    VT_REQUIRE(predicted && velocities && positions, "PredictPositions: NULL buffer");
    seam::PredictPositions(g_params, predicted, velocities, positions, deltaTime, g_stream);
    VT_API_END
}

int velvet_SolveStretch(float* predicted, float* deltas, int* deltaCounts, const int* stretchIndices,
                        const float* stretchLengths, const float* invMasses, unsigned numConstraints)
{
    VT_API_BEGIN
    if (numConstraints == 0) return VELVET_OK;
    VT_REQUIRE(predicted && deltas && deltaCounts && stretchIndices && stretchLengths && invMasses, "SolveStretch: NULL buffer");
    seam::SolveStretch(predicted, deltas, deltaCounts, stretchIndices, stretchLengths, invMasses, numConstraints, g_stream);
    VT_API_END
}

int velvet_SolveBending(float* predicted, float* deltas, int* deltaCounts, const unsigned* bendingIndices,
                        const float* bendingAngles, const float* invMass, unsigned numConstraints, float deltaTime)
{
    VT_API_BEGIN
    if (int e = need_params()) return e;
    if (numConstraints == 0) return VELVET_OK;
    VT_REQUIRE(predicted && deltas && deltaCounts && bendingIndices && bendingAngles && invMass, "SolveBending: NULL buffer");
    seam::SolveBending(g_params, predicted, deltas, deltaCounts, bendingIndices, bendingAngles, invMass, numConstraints,
                       deltaTime, g_stream);
    VT_API_END
}

int velvet_SolveAttachment(float* predicted, float* deltas, int* deltaCounts, const float* invMass,
                           const int* attachParticleIDs, const int* attachSlotIDs, const float* attachSlotPositions,
                           const float* attachDistances, int numConstraints)
{
    VT_API_BEGIN
    if (int e = need_params()) return e;
    if (numConstraints == 0) return VELVET_OK;
    VT_REQUIRE(numConstraints > 0 && predicted && deltas && deltaCounts && invMass && attachParticleIDs && attachSlotIDs &&
                   attachSlotPositions && attachDistances, "SolveAttachment: bad argument");
    seam::SolveAttachment(g_params, predicted, deltas, deltaCounts, invMass, attachParticleIDs, attachSlotIDs,
                          attachSlotPositions, attachDistances, numConstraints, g_stream);
    VT_API_END
}

int velvet_ApplyDeltas(float* predicted, float* deltas, int* deltaCounts)
{
    VT_API_BEGIN
    if (int e = need_params()) return e;
    if (g_params.numParticles == 0) return VELVET_OK;
    VT_REQUIRE(predicted && deltas && deltaCounts, "ApplyDeltas: NULL buffer");
    seam::ApplyDeltas(g_params, predicted, deltas, deltaCounts, g_stream);
    VT_API_END
}

int velvet_CollideSDF(float* predicted, const VtSDFCollider* colliders, const float* positions, unsigned numColliders,
                      float deltaTime)
{
    VT_API_BEGIN
    if (int e = need_params()) return e;
    if (numColliders == 0 || g_params.numParticles == 0) return VELVET_OK;
    VT_REQUIRE(predicted && colliders && positions, "CollideSDF: NULL buffer");
    seam::CollideSDF(g_params, predicted, colliders, positions, numColliders, deltaTime, g_stream);
    VT_API_END
}

int velvet_CollideParticles(float* deltas, int* deltaCounts, float* predicted, const float* invMasses,
                            const unsigned* neighbors, const float* positions)
{
    VT_API_BEGIN
    if (int e = need_params()) return e;
    if (g_params.numParticles == 0) return VELVET_OK;
    VT_REQUIRE(deltas && deltaCounts && predicted && invMasses && neighbors && positions, "CollideParticles: NULL buffer");
    seam::CollideParticles(g_params, deltas, deltaCounts, predicted, invMasses, neighbors, positions, g_stream);
    VT_API_END
}

int velvet_Finalize(float* velocities, float* positions, const float* predicted, float deltaTime)
{
    VT_API_BEGIN
    if (int e = need_params()) return e;
    if (g_params.numParticles == 0) return VELVET_OK;
    VT_REQUIRE(velocities && positions && predicted, "Finalize: NULL buffer");
    seam::Finalize(g_params, velocities, positions, predicted, deltaTime, g_stream);
    VT_API_END
}

int velvet_ComputeNormal(float* normals, const float* positions, const unsigned* indices, unsigned numTriangles)
{
    VT_API_BEGIN
    if (int e = need_params()) return e;
    if (g_params.numParticles == 0) return VELVET_OK;
    VT_REQUIRE(normals && positions && (numTriangles == 0 || indices), "ComputeNormal: NULL buffer");
    seam::ComputeNormal(g_params, normals, positions, indices, numTriangles, g_stream);
    VT_API_END
}

int velvet_HashObjects(unsigned* particleHash, unsigned* particleIndex, unsigned* cellStart, unsigned* cellEnd,
                       unsigned* neighbors, const float* positions, const float* originalPositions, VtHashParams params)
{
    VT_API_BEGIN
    if (params.numObjects == 0) return VELVET_OK;
    VT_REQUIRE(particleHash && particleIndex && cellStart && cellEnd && neighbors && positions && originalPositions,
               "HashObjects: NULL buffer");
    VT_REQUIRE(params.tableSize > 0, "HashObjects: tableSize <= 0");
    seam::HashObjects(particleHash, particleIndex, cellStart, cellEnd, neighbors, positions, originalPositions, params, g_stream);
    VT_API_END
}

int velvet_SortPairs(unsigned* keys, unsigned* values, unsigned numItems, int endBit)
{
    VT_API_BEGIN
    if (numItems == 0 || endBit <= 0) return VELVET_OK;
    VT_REQUIRE(keys && values && endBit <= 32, "SortPairs: bad argument");
    std::lock_guard<std::mutex> lk(g_mutex);
    SortScratch& S = acquire_scratch(numItems, g_stream);
    const int where = S.sorter.sort(keys, values, S.keysAlt.data(), S.valsAlt.data(), numItems, endBit, g_stream);
    if (where == 1) {
        VT_CUDA(cudaMemcpyAsync(keys, S.keysAlt.data(), sizeof(unsigned) * numItems, cudaMemcpyDeviceToDevice, g_stream));
        VT_CUDA(cudaMemcpyAsync(values, S.valsAlt.data(), sizeof(unsigned) * numItems, cudaMemcpyDeviceToDevice, g_stream));
    }
    release_scratch(S, g_stream);
    VT_API_END
}

int velvet_selftest_division(const float* x, const float* y, unsigned n, unsigned* outDiv, unsigned* outVec3, unsigned* outPlain)
{
    VT_API_BEGIN
    VT_REQUIRE(x && y && outDiv && outVec3 && outPlain && n > 0, "selftest_division: bad argument");
    velvet::selftest_division_kernel<<<(n + 255) / 256, 256>>>(x, y, n, outDiv, outVec3, outPlain);
    VT_CUDA(cudaGetLastError());
    VT_CUDA(cudaDeviceSynchronize());
    VT_API_END
}

int velvet_selftest_constraints(const float* operands, unsigned n, unsigned long long* mismatches3, unsigned long long* fast3)
{
    VT_API_BEGIN
    VT_REQUIRE(operands && mismatches3 && fast3 && n > 0, "selftest_constraints: bad argument");
    VT_CUDA(cudaMemset(mismatches3, 0, 3 * sizeof(unsigned long long)));
    VT_CUDA(cudaMemset(fast3, 0, 3 * sizeof(unsigned long long)));
    velvet::selftest_constraints_kernel<<<(n + 255) / 256, 256>>>(operands, n, mismatches3, fast3);
    velvet::selftest_sqrt_kernel<<<148 * 8, 256>>>(mismatches3, fast3);
    VT_CUDA(cudaGetLastError());
    VT_CUDA(cudaDeviceSynchronize());
    VT_API_END
}

int velvet_device_synchronize(void)
{
    VT_API_BEGIN
    VT_CUDA(cudaDeviceSynchronize());
    VT_API_END
}

int velvet_alloc(void** devPtr, size_t bytes)
{
    VT_API_BEGIN
    VT_REQUIRE(devPtr, "alloc: devPtr is NULL");
    *devPtr = nullptr;
    if (bytes) VT_CUDA(cudaMallocManaged(devPtr, bytes));
    VT_API_END
}

int velvet_free(void* devPtr)
{
    VT_API_BEGIN
    if (devPtr) VT_CUDA(cudaFree(devPtr));
    VT_API_END
}

int velvet_copy(void* dst, const void* src, size_t bytes)
{
    VT_API_BEGIN
    if (bytes) VT_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDefault));
    VT_API_END
}

}  // extern "C"
