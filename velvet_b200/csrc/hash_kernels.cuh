// hash_kernels.cuh -- spatial-hash kernels shared by the seam (packed float3 input) and the fused
// pipeline (float4 input).  Reference: SpatialHashGPU.cu L13-130.
//
// Integer path (must be bit-exact): ix = (int)floor(p.x / cell) with a true fp32 division,
// h = (ix*92837111) ^ (iy*689287499) ^ (iz*283923481) in wrapping int32, key = abs(h % tableSize).
#pragma once

#include "vt_math.cuh"

namespace velvet {

// position accessors
struct PosPacked3 {
    const float* p;
    __device__ __forceinline__ vec3 operator()(unsigned i) const { return load3(p, i); }
};
struct PosFloat4 {
    const float4* p;
    __device__ __forceinline__ vec3 operator()(unsigned i) const { return V3(__ldg(p + i)); }
};

// H1: ComputeParticleHash_Kernel (SpatialHashGPU.cu L34-42)
template <class Pos>
__global__ void __launch_bounds__(256) hash_particles_kernel(unsigned* __restrict__ particleHash,
                                                             unsigned* __restrict__ particleIndex, Pos positions,
                                                             unsigned n, float cellSpacing, int tableSize,
                                                             unsigned instanceParticles, unsigned* __restrict__ emptyTable = nullptr,
                                                             unsigned emptyCount = 0)
{
    // Batched independent cloths: instance i = id / instanceParticles owns table rows [i*tableSize, (i+1)*tableSize), so
    // one global sort orders every instance separately and instances never see each other's particles.  A single cloth
    // (or several interacting cloths, the reference's case) is instance 0 with instanceParticles == n.
    const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n) return;
    // fused pipeline: the cell table of this rebuild is cleared here (cellStart = "empty" for every bucket; FindCellStart runs
    // after the sort and nothing reads the table in between), which saves the fill launch in front of find_cell_start_kernel
    for (unsigned j = id; j < emptyCount; j += n) emptyTable[j] = 0xffffffffu;
    const vec3 p = positions(id);
    particleHash[id] = (unsigned)hash_coords(int_coord(p.x, cellSpacing), int_coord(p.y, cellSpacing),
                                             int_coord(p.z, cellSpacing), tableSize) +
                       (id / instanceParticles) * (unsigned)tableSize;
    particleIndex[id] = id;
}

// H3: FindCellStart_Kernel (SpatialHashGPU.cu L44-77).  The previous key comes straight from L1/L2 instead of
// a shared-memory stage; cellStart must have been filled with 0xffffffff, cellEnd is never cleared.
static __global__ void __launch_bounds__(256) find_cell_start_kernel(unsigned* __restrict__ cellStart,
                                                              unsigned* __restrict__ cellEnd,
                                                              const unsigned* __restrict__ particleHash, unsigned n)
{
    const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n) return;
    const unsigned hash = particleHash[id];
    if (id == 0) {
        cellStart[hash] = 0;
    } else {
        const unsigned prev = particleHash[id - 1];
        if (hash != prev) {
            cellStart[hash] = id;
            cellEnd[prev] = id;
        }
    }
    if (id == n - 1) cellEnd[hash] = id + 1;
}

// H4: CacheNeighbors_Kernel (SpatialHashGPU.cu L79-130).  One thread per particle id (the reference maps
// thread t to particleIndex[t]; the result is the same and id order makes the column-major stores
// neighbors[id + N*k] fully coalesced).  Traversal order x, y, z then bucket order is kept, so lists are
// bit-identical including duplicates (two of the 27 cells landing in one bucket) and the 64-entry caps.
template <class Pos, class Pos0>
__global__ void __launch_bounds__(256) cache_neighbors_kernel(unsigned* __restrict__ neighbors,
                                                              const unsigned* __restrict__ particleIndex,
                                                              const unsigned* __restrict__ cellStart,
                                                              const unsigned* __restrict__ cellEnd, Pos positions,
                                                              Pos0 originalPositions, VtHashParams hp)
{
    const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= hp.numObjects) return;
    const vec3 position = positions(id);
    const vec3 originalPos = originalPositions(id);
    const int ix = int_coord(position.x, hp.cellSpacing);
    const int iy = int_coord(position.y, hp.cellSpacing);
    const int iz = int_coord(position.z, hp.cellSpacing);

    const unsigned long long limit = (unsigned long long)hp.numObjects * hp.maxNumNeighbors;
    unsigned long long neighborIndex = id;
    for (int x = ix - 1; x <= ix + 1; x++)
        for (int y = iy - 1; y <= iy + 1; y++)
            for (int z = iz - 1; z <= iz + 1; z++) {
                const int h = hash_coords(x, y, z, hp.tableSize);
                const unsigned start = __ldg(cellStart + h);
                if (start == 0xffffffffu) continue;
                unsigned end = __ldg(cellEnd + h);
                if (start + hp.maxNumNeighbors < end) end = start + hp.maxNumNeighbors;
                for (unsigned i = start; i < end; i++) {
                    const unsigned nb = __ldg(particleIndex + i);
                    if (nb != id && (length2(position - positions(nb)) < hp.cellSpacing2) &&
                        (length2(originalPos - originalPositions(nb)) > hp.particleDiameter2)) {
                        neighbors[neighborIndex] = nb;
                        neighborIndex += hp.numObjects;
                        if (neighborIndex >= limit) return;
                    }
                }
            }
    if (neighborIndex < limit) neighbors[neighborIndex] = 0xffffffffu;
}


// ---------------------------------------------------------------- fused-pipeline variant of H4
//
// Same result as cache_neighbors_kernel, bit for bit, restructured for the SIMT machine (ncu on the direct
// transcription: 9 300 thread-instructions per particle, instruction-bound by divergent 27-cell loops and a runtime `%`):
//   * reorder pass: sortedPos[i] = (pred[particleIndex[i]], particle id in .w), sortedInit[i] likewise, so the
//     candidates of a bucket are contiguous 16-byte loads instead of an index -> position dependent gather;
//   * abs(h % tableSize) == |h| mod tableSize is computed with an exact multiply-high reduction (Lemire fastmod);
//   * phase 1 writes the non-empty bucket ranges of the 27 cells (x, y, z order) to a shared-memory strip,
//     phase 2 is ONE flat loop over all candidates, so a warp only diverges on the accept branches.
struct FastMod {
    unsigned long long M;  // ceil(2^64 / d)
    unsigned d;
    unsigned pow2Mask;     // d - 1 when d is a power of two (tableSize = 2 * maxNumObjects: every power-of-two cloth), else 0
};
inline FastMod make_fastmod(unsigned d)
{
    FastMod f;
    f.d = d;
    f.M = 0xFFFFFFFFFFFFFFFFull / d + 1;
    f.pow2Mask = (d & (d - 1)) == 0 ? d - 1 : 0u;
    return f;
}
__device__ __forceinline__ unsigned fastmod_u32(unsigned a, const FastMod f)
{
    if (f.pow2Mask) return a & f.pow2Mask;  // uniform branch: one AND instead of two 64-bit multiplies
    const unsigned long long low = f.M * a;
    return (unsigned)__umul64hi(low, (unsigned long long)f.d);
}
__device__ __forceinline__ unsigned hash_key_fast(int hx, int hy, int hz, const FastMod f)
{
    const int h = hx ^ hy ^ hz;
    const unsigned mag = h < 0 ? (unsigned)(-(long long)h) : (unsigned)h;  // |h|, INT_MIN -> 2^31
    return fastmod_u32(mag, f);
}

// One 32-byte record per sorted slot: {predicted xyz, cell tag} {registration-time xyz, particle id}.  Both distance tests
// of a candidate read the same 32-byte sector; the second half is only fetched for candidates that pass the first test.
//
// Cell tag: the candidate's integer cell coordinates, each reduced mod 256, in three 10-bit fields.
// A bucket of the reference's table mixes particles of unrelated cells (tableSize is a power of two: 207 936 occupied cells
// fold into 138 605 buckets on a flat 1024^2 sheet), so ~60 % of the candidates a particle meets cannot be neighbours at
// all.  tag + (258 - own coordinate mod 256) per field has its low eight bits in [0, 7] exactly when the candidate's cell is
// -2 .. +5 cells from the particle's on that axis (mod 256): ONE add and ONE masked compare reject every candidate whose
// cell is three or more cells away on some axis, i.e. at least 2 cell widths apart, which the reference's
// `distance^2 < cellSpacing^2` test rejects as well.  The filter is a superset test (a cell a multiple of 256 cells away
// passes too and is then rejected by the float test), so the lists stay bit-identical; it only removes the float work and
// the second half of the predicate for the foreign candidates.
// (Round 1 clamped the coordinates to +-256 cells instead of wrapping them: correct, but on a cloth wider than 512 cells --
// 4096^2 spans 1 820 -- most tags were clamped and the filter let nearly everything through.)
struct __align__(32) SortedParticle {
    float4 pos;   // w = cell tag bits
    float4 init;  // w = particle id bits
};
constexpr unsigned CN_TAG_FIELD = 10, CN_TAG_K = 258;
constexpr unsigned CN_TAG_MASK = 0xF8u | (0xF8u << 10) | (0xF8u << 20);
constexpr unsigned CN_TAG_WANT = 0u;

__device__ __forceinline__ unsigned cn_tag_coord(int i) { return (unsigned)i & 255u; }  // two's complement: i mod 256
__device__ __forceinline__ unsigned cn_cell_tag(int ix, int iy, int iz)
{
    return cn_tag_coord(ix) | (cn_tag_coord(iy) << CN_TAG_FIELD) | (cn_tag_coord(iz) << (2 * CN_TAG_FIELD));
}
// what a particle adds to a candidate's tag: per field in [3, 258], so tag + probe <= 513 never carries into the next field
__device__ __forceinline__ unsigned cn_cell_probe(int ix, int iy, int iz)
{
    return (CN_TAG_K - cn_tag_coord(ix)) | ((CN_TAG_K - cn_tag_coord(iy)) << CN_TAG_FIELD) |
           ((CN_TAG_K - cn_tag_coord(iz)) << (2 * CN_TAG_FIELD));
}

static __global__ void __launch_bounds__(256) reorder_sorted_kernel(SortedParticle* __restrict__ sorted,
                                                                    const unsigned* __restrict__ particleIndex,
                                                                    const float4* __restrict__ pred,
                                                                    const float4* __restrict__ init4, unsigned n,
                                                                    float cellSpacing, unsigned* __restrict__ cellStart,
                                                                    unsigned* __restrict__ cellEnd,
                                                                    const unsigned* __restrict__ particleHash)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (particleHash) {  // FindCellStart_Kernel (H3, find_cell_start_kernel above) for sorted slot i, in the same launch
        const unsigned hash = particleHash[i];
        if (i == 0) {
            cellStart[hash] = 0;
        } else {
            const unsigned prev = particleHash[i - 1];
            if (hash != prev) {
                cellStart[hash] = i;
                cellEnd[prev] = i;
            }
        }
        if (i == n - 1) cellEnd[hash] = i + 1;
    }
    const unsigned id = particleIndex[i];
    float4 p = __ldg(pred + id);
    float4 o = __ldg(init4 + id);
    p.w = __uint_as_float(cn_cell_tag(int_coord(p.x, cellSpacing), int_coord(p.y, cellSpacing), int_coord(p.z, cellSpacing)));
    o.w = __uint_as_float(id);
    sorted[i].pos = p;
    sorted[i].init = o;
}

constexpr int CN_THREADS = 256;

// The reference's hash maps many cells to one bucket when tableSize is a power of two (measured on a flat 1024^2 sheet:
// 207 936 cells -> 138 605 buckets, mean bucket 7.6, ~105 candidates per particle), and two of the 27 cells may share a
// bucket, which is then scanned twice: all of that defines the lists and is kept.
//
// Thread t handles the particle in sorted slot t (the reference's mapping, SpatialHashGPU.cu L87-88): the lanes of a
// warp sit in the same few buckets and walk the same candidate runs (mostly convergent loops, broadcast loads); the
// scattered 4-byte column stores neighbors[id + N*k] are absorbed by L2 (the touched part of the table is ~50 MB).
// The walk is software-pipelined over buckets (the range of the next non-empty bucket is fetched while the candidates of
// the current one are tested) and four candidates are in flight per trip.
// Domain-decomposed mode: the sorted slots whose particle this rank owns, compacted (warp ballot + one atomic per warp), so
// that the candidate walk below runs with full warps over numOwned threads.  Owned particles are scattered uniformly over
// the sorted order (the hash is a hash): walking all slots and returning early for the others kept ~1/world of the lanes of
// EVERY warp busy, and the walk barely got faster with more ranks (8 GPUs, 16.7M particles: the largest stage of the frame).
// The order of the compacted list depends on warp arrival; every thread writes only its own particle's list, so the
// neighbour table does not.
static __global__ void __launch_bounds__(256) compact_owned_slots_kernel(const SortedParticle* __restrict__ sorted,
                                                                         const unsigned char* __restrict__ ownedMask, unsigned n,
                                                                         unsigned* __restrict__ slots, unsigned* __restrict__ counter)
{
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    bool owned = false;
    if (t < n) owned = __ldg(ownedMask + __float_as_uint(__ldg(&sorted[t].init.w))) != 0;
    const unsigned ballot = __ballot_sync(0xffffffffu, owned);
    const unsigned lane = threadIdx.x & 31u;
    unsigned base = 0;
    if (lane == 0 && ballot) base = atomicAdd(counter, __popc(ballot));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (owned) slots[base + __popc(ballot & ((1u << lane) - 1u))] = t;
}

// Large cloths: the sorted slots of the particles [begin, begin + count), grouped into BANDS of bandSize consecutive particle
// indices (slots[b * bandSize ..] holds band b).  Sorted order is hash order, i.e. spatially random: walking 16.7 M slots in
// that order keeps every candidate record (537 MB) and the whole cell table in play at once and misses the L2 all the time
// (0.54 ns per particle against 0.33 ns at 1 M).  Particle indices of a grid cloth are row-major, so a band of 2^20 indices is a
// compact piece of cloth whose candidates fit the L2; inside a band the slots stay in sorted order (block by block), so the
// lanes of a warp still share their buckets.  counters[numBands] must be zero.
constexpr unsigned CN_MAX_BANDS = 1024;
static __global__ void __launch_bounds__(256) band_slots_kernel(const SortedParticle* __restrict__ sorted, unsigned n, unsigned begin,
                                                                unsigned count, unsigned bandSize, unsigned numBands,
                                                                unsigned* __restrict__ slots, unsigned* __restrict__ counters)
{
    __shared__ unsigned s_cnt[CN_MAX_BANDS], s_base[CN_MAX_BANDS];
    for (unsigned b = threadIdx.x; b < numBands; b += blockDim.x) s_cnt[b] = 0;
    __syncthreads();
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned band = 0xffffffffu, rank = 0;
    if (t < n) {
        const unsigned rel = __float_as_uint(__ldg(&sorted[t].init.w)) - begin;
        if (rel < count) {
            band = rel / bandSize;
            rank = atomicAdd(&s_cnt[band], 1u);
        }
    }
    __syncthreads();
    for (unsigned b = threadIdx.x; b < numBands; b += blockDim.x)
        if (s_cnt[b]) s_base[b] = atomicAdd(&counters[b], s_cnt[b]);
    __syncthreads();
    if (band != 0xffffffffu) slots[(size_t)band * bandSize + s_base[band] + rank] = t;
}

// SMEM_KEYS: see s_key below (the better trade up to a few million particles; beyond, the walk is bound by misses of the
// candidate records and the shared-memory carve-out costs more L1 than the saved instructions are worth)
template <bool SMEM_KEYS>
static __global__ void __launch_bounds__(CN_THREADS) cache_neighbors_sorted_kernel(
    unsigned* __restrict__ neighbors, const unsigned* __restrict__ cellStart, const unsigned* __restrict__ cellEnd,
    const SortedParticle* __restrict__ sorted, VtHashParams hp, FastMod fm, unsigned instanceParticles,
    const unsigned* __restrict__ ownedSlots, const unsigned numThreads)
{
    // the particle's 27 bucket keys, [bucket][thread] (conflict-free): phase 1 computes them with compile-time cell offsets,
    // phase 2 reads back the ones it visits -- recomputing a key there (runtime cell offset: two divisions, three selects, the
    // modulo) was 11 % of the kernel's instructions, and the key registers it kept alive cost a CTA of occupancy (48 -> 40)
    __shared__ unsigned s_key[SMEM_KEYS ? 27 * CN_THREADS : 1];
    const unsigned ti = blockIdx.x * CN_THREADS + threadIdx.x;
    if (ti >= numThreads) return;
    const unsigned t = ownedSlots ? __ldg(ownedSlots + ti) : ti;  // decomposed mode: only the slots of owned particles
    const float4 me = __ldg(&sorted[t].pos), me0 = __ldg(&sorted[t].init);
    const unsigned id = __float_as_uint(me0.w);
    const unsigned tableBase = (id / instanceParticles) * (unsigned)hp.tableSize;  // this instance's rows of the table
    const vec3 position = V3(me);
    const vec3 originalPos = V3(me0);
    const int ix = int_coord(position.x, hp.cellSpacing);
    const int iy = int_coord(position.y, hp.cellSpacing);
    const int iz = int_coord(position.z, hp.cellSpacing);
    const unsigned probe = cn_cell_probe(ix, iy, iz);
    const int hx0 = (int)((unsigned)(ix - 1) * 92837111u), hx1 = (int)((unsigned)ix * 92837111u), hx2 = (int)((unsigned)(ix + 1) * 92837111u);
    const int hy0 = (int)((unsigned)(iy - 1) * 689287499u), hy1 = (int)((unsigned)iy * 689287499u), hy2 = (int)((unsigned)(iy + 1) * 689287499u);
    const int hz0 = (int)((unsigned)(iz - 1) * 283923481u), hz1 = (int)((unsigned)iz * 283923481u), hz2 = (int)((unsigned)(iz + 1) * 283923481u);
    auto key_of = [&](int b) {  // !SMEM_KEYS: the key of bucket b, recomputed
        const int a = b / 9, m = (b / 3) % 3, c = b % 3;
        const int hx = a == 0 ? hx0 : a == 1 ? hx1 : hx2;
        const int hy = m == 0 ? hy0 : m == 1 ? hy1 : hy2;
        const int hz = c == 0 ? hz0 : c == 1 ? hz1 : hz2;
        return tableBase + hash_key_fast(hx, hy, hz, fm);
    };
    // phase 1: which of the 27 buckets (x, y, z traversal order = bit order) are non-empty; 27 independent loads in flight
    unsigned mask = 0;
#pragma unroll
    for (int b = 0; b < 27; b++) {
        const int hx = (b / 9) == 0 ? hx0 : (b / 9) == 1 ? hx1 : hx2;
        const int hy = ((b / 3) % 3) == 0 ? hy0 : ((b / 3) % 3) == 1 ? hy1 : hy2;
        const int hz = (b % 3) == 0 ? hz0 : (b % 3) == 1 ? hz1 : hz2;
        const unsigned key = tableBase + hash_key_fast(hx, hy, hz, fm);
        if (SMEM_KEYS) s_key[b * CN_THREADS + threadIdx.x] = key;
        if (__ldg(cellStart + key) != 0xffffffffu) mask |= 1u << b;
    }

    // phase 2: walk the non-empty buckets in traversal order
    const unsigned N = hp.numObjects, K = hp.maxNumNeighbors;
    const float cs2 = hp.cellSpacing2, pd2 = hp.particleDiameter2;
    unsigned* out = neighbors + id;
    unsigned k = 0;
    // ONE flat loop over the candidates of all buckets: a lane that finishes a bucket moves on to its next one inside the loop
    // body.  (With a loop over buckets around a loop over candidates, a warp -- whose lanes sit in ~4 different cells, each
    // with its own sequence of bucket lengths -- ran every bucket for as long as its slowest lane: 15.6 of 32 lanes busy.)
    // The range of the bucket after the current one is always in flight while the current one is tested.
    auto fetch_range = [&](unsigned& first, unsigned& last) {
        first = last = 0;
        if (mask) {
            const unsigned key = SMEM_KEYS ? s_key[(__ffs(mask) - 1) * CN_THREADS + threadIdx.x] : key_of(__ffs(mask) - 1);
            mask &= mask - 1;
            first = __ldg(cellStart + key);
            last = __ldg(cellEnd + key);
        }
    };
    unsigned cur, end, nextCur, nextEnd;
    fetch_range(cur, end);  // a listed bucket is never empty
    if (cur + K < end) end = cur + K;
    fetch_range(nextCur, nextEnd);
    while (cur < end) {
        float4 q[4];
#pragma unroll
        for (int j = 0; j < 4; j++) q[j] = __ldg(&sorted[cur + j < end ? cur + j : cur].pos);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (cur + j >= end) break;
            if (((__float_as_uint(q[j].w) + probe) & CN_TAG_MASK) != CN_TAG_WANT) continue;  // cell >= 3 cells away
            if (!(length2(position - V3(q[j])) < cs2)) continue;
            const float4 o = __ldg(&sorted[cur + j].init);  // same 32-byte sector as q[j]: an L1 hit
            const unsigned nb = __float_as_uint(o.w);
            if (nb != id && length2(originalPos - V3(o)) > pd2) {
                out[(size_t)k * N] = nb;
                if (++k >= K) return;
            }
        }
        cur += 4;
        if (cur >= end) {  // on to the next bucket (its range has arrived by now), and ask for the one after it
            cur = nextCur;
            end = nextEnd;
            if (cur + K < end) end = cur + K;
            fetch_range(nextCur, nextEnd);
        }
    }
    if (k < K) out[(size_t)k * N] = 0xffffffffu;
}

}  // namespace velvet
