// radix_sort.cu -- see radix_sort.cuh.  Hand-written onesweep LSD radix sort for sm_100a.
#include "radix_sort.cuh"

namespace velvet {

namespace {

constexpr int RS_BITS = 8;
constexpr int RS_RADIX = 1 << RS_BITS;
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_MAX_PASSES = 4;
constexpr unsigned RS_FLAG_AGG = 1u << 30;   // tile aggregate published
constexpr unsigned RS_FLAG_INCL = 2u << 30;  // inclusive prefix published
constexpr unsigned RS_VALUE_MASK = (1u << 30) - 1u;

__host__ __device__ inline unsigned pass_mask(int pass, int endBit)
{
    int bits = endBit - pass * RS_BITS;
    if (bits > RS_BITS) bits = RS_BITS;
    return (1u << bits) - 1u;
}

// One read of the keys builds the digit histograms of every pass.
__global__ void __launch_bounds__(RS_THREADS) rs_histogram_kernel(const unsigned* __restrict__ keys, unsigned n,
                                                                  int endBit, int passes, unsigned* __restrict__ hist)
{
    __shared__ unsigned sh[RS_MAX_PASSES * RS_RADIX];
    for (int i = threadIdx.x; i < passes * RS_RADIX; i += RS_THREADS) sh[i] = 0;
    __syncthreads();
    const unsigned stride = gridDim.x * RS_THREADS;
    for (unsigned i = blockIdx.x * RS_THREADS + threadIdx.x; i < n; i += stride) {
        const unsigned key = keys[i];
        for (int p = 0; p < passes; p++) atomicAdd(&sh[p * RS_RADIX + ((key >> (p * RS_BITS)) & pass_mask(p, endBit))], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * RS_RADIX; i += RS_THREADS)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

__device__ inline unsigned block_exclusive_scan_256(unsigned v, unsigned* s_warp /*[RS_WARPS]*/)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    unsigned warpBase = 0;
    for (int i = 0; i < w; i++) warpBase += s_warp[i];
    __syncthreads();
    return warpBase + incl - v;
}

// One digit pass.  Tile order is claimed dynamically so that look-back only ever waits on running tiles.
template <int ITEMS>
__global__ void __launch_bounds__(RS_THREADS)
rs_onesweep_kernel(const unsigned* __restrict__ keysIn, const unsigned* __restrict__ valsIn,
                   unsigned* __restrict__ keysOut, unsigned* __restrict__ valsOut, unsigned n, int shift,
                   unsigned mask, const unsigned* __restrict__ hist /* digit counts of this pass */, unsigned* lookback,
                   unsigned* tileCounter)
{
    constexpr int TILE = RS_THREADS * ITEMS;
    __shared__ unsigned s_tile;
    __shared__ unsigned s_whist[RS_WARPS][RS_RADIX];
    __shared__ unsigned s_goff[RS_RADIX];
    __shared__ unsigned s_tbase[RS_RADIX];
    __shared__ unsigned s_warp[RS_WARPS];
    __shared__ unsigned s_keys[TILE];
    __shared__ unsigned s_vals[TILE];

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(tileCounter, 1u);
    for (int i = tid; i < RS_WARPS * RS_RADIX; i += RS_THREADS) (&s_whist[0][0])[i] = 0;
    __syncthreads();
    const unsigned tile = s_tile;
    const unsigned tileStart = tile * (unsigned)TILE;
    const unsigned segStart = tileStart + (unsigned)(w * 32 * ITEMS);

    unsigned key[ITEMS];
    unsigned rank[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        const unsigned idx = segStart + k * 32 + lane;
        key[k] = idx < n ? keysIn[idx] : 0xffffffffu;
    }
    const unsigned ltMask = (1u << lane) - 1u;
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        const unsigned idx = segStart + k * 32 + lane;
        const bool valid = idx < n;
        const unsigned d = valid ? ((key[k] >> shift) & mask) : (unsigned)RS_RADIX;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        unsigned pre = 0;
        if (lane == leader && valid) {
            pre = s_whist[w][d];
            s_whist[w][d] = pre + __popc(peers);
        }
        pre = __shfl_sync(0xffffffffu, pre, leader);
        rank[k] = pre + __popc(peers & ltMask);
        __syncwarp();
    }
    __syncthreads();

    // thread d owns digit d: exclusive prefix across warps, tile count, decoupled look-back
    unsigned tcount = 0;
#pragma unroll
    for (int i = 0; i < RS_WARPS; i++) {
        const unsigned c = s_whist[i][tid];
        s_whist[i][tid] = tcount;
        tcount += c;
    }
    volatile unsigned* lb = lookback;
    unsigned excl = 0;
    if (tile == 0) {
        lb[tid] = tcount | RS_FLAG_INCL;
    } else {
        lb[tile * RS_RADIX + tid] = tcount | RS_FLAG_AGG;
        int p = (int)tile - 1;
        while (true) {
            unsigned v = lb[p * RS_RADIX + tid];
            while ((v >> 30) == 0) v = lb[p * RS_RADIX + tid];
            excl += v & RS_VALUE_MASK;
            if (v & RS_FLAG_INCL) break;
            p--;
        }
        lb[tile * RS_RADIX + tid] = (excl + tcount) | RS_FLAG_INCL;
    }
    const unsigned tbase = block_exclusive_scan_256(tcount, s_warp);
    // first output slot of digit `tid` in this pass: exclusive scan of the pass's histogram (every tile redoes this 256-bin
    // scan instead of a one-block kernel in front of the passes: one launch less per sort)
    const unsigned digitBase = block_exclusive_scan_256(hist[tid], s_warp);
    s_tbase[tid] = tbase;
    s_goff[tid] = digitBase + excl - tbase;
    __syncthreads();

#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        const unsigned idx = segStart + k * 32 + lane;
        if (idx < n) {
            const unsigned d = (key[k] >> shift) & mask;
            const unsigned pos = s_tbase[d] + s_whist[w][d] + rank[k];
            s_keys[pos] = key[k];
            s_vals[pos] = valsIn[idx];
        }
    }
    __syncthreads();

    const unsigned count = (n - tileStart) < (unsigned)TILE ? (n - tileStart) : (unsigned)TILE;
    for (unsigned i = tid; i < count; i += RS_THREADS) {
        const unsigned k = s_keys[i];
        const unsigned dst = s_goff[(k >> shift) & mask] + i;
        keysOut[dst] = k;
        valsOut[dst] = s_vals[i];
    }
}

inline unsigned tile_items(unsigned n) { return n > 262144u ? RS_THREADS * 16u : RS_THREADS * 4u; }
inline unsigned num_tiles(unsigned n) { return (n + tile_items(n) - 1) / tile_items(n); }

}  // namespace

void RadixSorter::reserve(unsigned n)
{
    unsigned tiles = num_tiles(n);
    if (tiles < 256u) tiles = 256u;
    if (tiles > m_reservedTiles) {
        m_scratch.allocate((size_t)RS_MAX_PASSES * RS_RADIX + RS_MAX_PASSES + (size_t)RS_MAX_PASSES * tiles * RS_RADIX);
        m_reservedTiles = tiles;
    }
}

int RadixSorter::sort(unsigned* keysA, unsigned* valsA, unsigned* keysB, unsigned* valsB, unsigned n, int endBit,
                      cudaStream_t stream)
{
    m_lastLaunches = 0;
    const int passes = numPasses(endBit);
    if (n == 0 || passes == 0) return 0;
    if (passes > RS_MAX_PASSES) throw Error(-1, "RadixSorter: endBit > 32");
    reserve(n);
    const unsigned tiles = num_tiles(n);
    unsigned* hist = m_scratch.data();
    unsigned* counters = hist + RS_MAX_PASSES * RS_RADIX;
    unsigned* lookback = counters + RS_MAX_PASSES;
    const size_t clearWords = (size_t)RS_MAX_PASSES * RS_RADIX + RS_MAX_PASSES + (size_t)passes * m_reservedTiles * RS_RADIX;
    VT_CUDA(cudaMemsetAsync(hist, 0, clearWords * sizeof(unsigned), stream));

    unsigned histBlocks = (n + RS_THREADS * 8 - 1) / (RS_THREADS * 8);
    if (histBlocks > 148u * 8u) histBlocks = 148u * 8u;
    rs_histogram_kernel<<<histBlocks, RS_THREADS, 0, stream>>>(keysA, n, endBit, passes, hist);
    m_lastLaunches += 1;

    unsigned *kin = keysA, *vin = valsA, *kout = keysB, *vout = valsB;
    for (int p = 0; p < passes; p++) {
        const int shift = p * RS_BITS;
        const unsigned mask = pass_mask(p, endBit);
        unsigned* lb = lookback + (size_t)p * m_reservedTiles * RS_RADIX;
        if (tile_items(n) == RS_THREADS * 16u)
            rs_onesweep_kernel<16><<<tiles, RS_THREADS, 0, stream>>>(kin, vin, kout, vout, n, shift, mask,
                                                                     hist + p * RS_RADIX, lb, counters + p);
        else
            rs_onesweep_kernel<4><<<tiles, RS_THREADS, 0, stream>>>(kin, vin, kout, vout, n, shift, mask,
                                                                    hist + p * RS_RADIX, lb, counters + p);
        m_lastLaunches++;
        unsigned* t;
        t = kin; kin = kout; kout = t;
        t = vin; vin = vout; vout = t;
    }
    VT_CUDA(cudaGetLastError());
    return passes & 1;
}

}  // namespace velvet
