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
                                                             unsigned n, float cellSpacing, int tableSize)
{
    const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n) return;
    const vec3 p = positions(id);
    particleHash[id] = (unsigned)hash_coords(int_coord(p.x, cellSpacing), int_coord(p.y, cellSpacing),
                                             int_coord(p.z, cellSpacing), tableSize);
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

}  // namespace velvet
