// fused_kernels.cuh -- launch interface of the fused sm_100a substep pipeline (see fused_kernels.cu).
#pragma once

#include <cuda_runtime.h>

#include "dd_peer.cuh"
#include "grid_plan.hpp"
#include "tile_plan.hpp"
#include "vt_math.cuh"

namespace velvet {

// Refreshed in device memory before every frame; graph topology never depends on these values.
struct FrameParams {
    VtSimParams P;
    float frameTime;
    float substepTime;
    float xpbdBend;  // bendCompliance / substepTime / substepTime (VtClothSolverGPU.cu L173)
    unsigned numColliders;
};

struct TilePlanDev {
    const TileDesc* tiles;
    const unsigned* ownedIds;
    const unsigned* haloIds;
    const uint16_t* cnt16;  // stretch | bend << 8 constraint counts per owned particle
    const unsigned* attOff;
    const uint2* stretchRec;
    const uint4* bendRec;
    const uint2* attachRec;
    unsigned numTiles, maxLocals, maxKS, maxKB, tileSize, maxBendPerTile, maxStretchPerTile;
    unsigned residentCtas;  // grid of the persistent iterate kernel: CTAs the device keeps resident at once
    unsigned threads;     // slot-row width: the power of two >= tileSize
    unsigned ctaThreads;  // CTA size: `threads`, or 1.25x that when a tile has more bending constraints than particles
    unsigned hasAttach;
};

// Grid cloths recognised among the registered constraints (grid_plan.hpp): everything iterate_grid_kernel reads.
struct GridPlanDev {
    const GridCloth* cloths;
    const float4* rest4;      // per particle: rest lengths of the stretch constraints generated at that vertex
    const float* restAngle;   // per particle: rest angle of the quad's bending constraint; NULL when all are `uniformAngle`
    float uniformAngle;       // (the reference registers 0 for every quad, VtClothObjectGPU.hpp L128-129)
    const unsigned* attOff;   // attach CSR by particle
    const uint2* attachRec;   // {slot id, distance bits}
    unsigned numCloths, numTiles, hasAttach;
    unsigned residentCtas;
    unsigned tilesY0;  // tiles along one side of cloth 0 (strip decomposition of a single cloth)
    unsigned tileX, tileY;  // owned particles per tile (grid_plan.hpp): 15 x 15, or 14 x 16
};

// candidate walk of a large grid cloth: in bands of 2^18 particle indices once the candidate records outgrow the L2
// (measured: no gain at 1.0 M particles, -23 % at 2.1 M, -26 % at 3.0 M, -31 % at 4.2 M, -23 % at 16.7 M)
constexpr unsigned VT_WALK_BAND_PARTICLES = 1u << 18, VT_WALK_BAND_MIN_PARTICLES = 3u << 19;
constexpr unsigned VT_WALK_SMEM_KEYS_MAX = 4u << 20;  // particles up to which the candidate walk keeps its bucket keys in shared memory
constexpr unsigned VT_GRID_MAX_ITERATIONS_PER_LAUNCH = 64;  // more Jacobi iterations than this per substep: one launch each
constexpr unsigned VT_MAX_COLLIDERS = 64;  // staged per block in shared memory (196 B each)
constexpr int VT_MAX_TILE = 512;           // particles (= threads) per Jacobi tile, upper bound

struct FusedLaunch {
    cudaStream_t stream;
    unsigned numParticles;  // all instances together
};

// Batched independent cloths (BASELINE config 4): `count` copies of one cloth topology, `particles` particles each.
// All instances share ONE tile plan / constraint set (it stays L2-resident) and differ only in their state, hash-table
// rows and attach-slot positions.  A plain solver is {1, numParticles, numSlots}.
struct Instancing {
    unsigned count;
    unsigned particles;
    unsigned slots;  // attach slots per instance
};

// The floating-point kernels are compiled twice from the same source (fused_kernels.cu):
//   exact_math  -fmad=false, IEEE division / square root: bit-identical to the CPU oracle (oracle/ref_jacobi_cpu.c)
//   fast_math   FMA contraction + approximate division / square root (nvcc -fmad=true -prec-div=false
//               -prec-sqrt=false): ~35 % fewer instructions in the Jacobi kernel, still deterministic (no atomics,
//               fixed summation order), within north_star's tolerance of the oracle but not bit-identical to it.
// The spatial-hash kernels exist only in exact_math: cell keys, sorted order and neighbour lists are integer work
// and must be bit-exact in every mode.
namespace exact_math {

// Per-frame inputs the host may have rewritten in managed memory: colliders (prepared once, with
// lastTransform * invCurTransform hoisted) and attach slot positions (copied to device scratch).
void launch_prepare_inputs(const FusedLaunch& L, const VtSDFCollider* colliders, PreparedCollider* prepared,
                           const float* slotPositions, float* slotPositionsOut, unsigned numSlotFloats,
                           const FrameParams* fp);

// AoS import + pre-stabilisation SDF pass (frame dt) + PredictPositions of substep 0.
void launch_begin_frame(const FusedLaunch& L, const float* positions, const float* velocities, const float* invMasses,
                        float4* pos4, float4* pred, const PreparedCollider* colliders,
                        const FrameParams* fp);

// CollideParticles + ApplyDeltas + CollideSDF (substep dt): predIn -> predOut.
void launch_collide(const FusedLaunch& L, const float4* predIn, float4* predOut, const float4* pos4,
                    const unsigned* neighbors, const PreparedCollider* colliders, const FrameParams* fp,
                    bool selfCollision, const unsigned* subset = nullptr, unsigned subsetCount = 0);
// the same for the contiguous particle range [begin, begin + count) (one strip of a decomposed grid cloth)
void launch_collide_range(const FusedLaunch& L, const float4* predIn, float4* predOut, const float4* pos4, const unsigned* neighbors,
                          const PreparedCollider* colliders, const FrameParams* fp, bool selfCollision, unsigned begin, unsigned count);

// One Jacobi iteration: SolveStretch + SolveAttachment + SolveBending + ApplyDeltas, predIn -> predOut.
void launch_iterate(const FusedLaunch& L, const float4* predIn, float4* predOut, const TilePlanDev& plan,
                    const float* attachSlotPositions, const FrameParams* fp, Instancing inst);
size_t iterate_smem_bytes(const TilePlanDev& plan);
// The same iteration for grid cloths without index records (iterate_grid_kernel): bit-identical to launch_iterate.
// iterations > 1 (no strip): that many Jacobi iterations in ONE launch, separated by grid-wide barriers on *gridBarrier; the
// result is in predOut when `iterations` is odd, in predIn when it is even (both arrays are overwritten along the way)
void launch_iterate_grid(const FusedLaunch& L, float4* predIn, float4* predOut, const GridPlanDev& plan,
                         const float* attachSlotPositions, const FrameParams* fp, Instancing inst,
                         const ddpeer::StripArgs* strip = nullptr,  // strip: this rank's rows of a decomposed cloth (dd_peer.cuh)
                         unsigned iterations = 1, unsigned* gridBarrier = nullptr);
unsigned configure_iterate_grid_kernel();  // returns resident CTAs on the device
unsigned configure_iterate_kernel(size_t smemBytes, unsigned threads, unsigned ctaThreads);  // opt in to > 48 KB smem; returns resident CTAs on the device

// Finalize of substep s fused with PredictPositions of substep s+1 (or, on the last substep, with the export
// of positions / velocities / predicted to the public packed-float3 buffers).
void launch_end_substep(const FusedLaunch& L, const float4* predIn, float4* pos4, float4* predNext,
                        bool last, float* positionsOut, float* velocitiesOut, float* predictedOut,
                        const FrameParams* fp);

// ComputeNormal as a per-vertex gather over incident triangles (ascending triangle id).
void launch_normals(const FusedLaunch& L, const float4* pos4, const unsigned* indices, const unsigned* vtxTriOff,
                    const unsigned* vtxTris, float* normalsOut, Instancing inst);

// spatial hash on float4 positions (keys/vals may be the alternate sort buffers)
// (also clears the cell table of the rebuild: cellStart[0, tableSize) = empty)
void launch_hash_particles(const FusedLaunch& L, unsigned* keys, unsigned* vals, const float4* pred, float cellSpacing,
                           int tableSizePerInstance, Instancing inst, unsigned* cellStart, int tableSize);
void launch_find_cell_start(const FusedLaunch& L, unsigned* cellStart, unsigned* cellEnd, const unsigned* particleHash);
void launch_cache_neighbors(const FusedLaunch& L, unsigned* neighbors, const unsigned* particleIndex,
                            const unsigned* cellStart, const unsigned* cellEnd, const float4* pred, const float4* init4,
                            VtHashParams hp);
// Restructured H4 (see hash_kernels.cuh): reorder pass + sorted-order candidate walk.  Returns the number of launches
// (kernels + memset nodes) it made, 0 when it cannot run (degenerate table); the caller then uses launch_cache_neighbors.
int launch_cache_neighbors_sorted(const FusedLaunch& L, unsigned* neighbors, const unsigned* particleIndex,
                                  const unsigned* cellStart, const unsigned* cellEnd, const float4* pred,
                                  const float4* init4, float4* sortedScratch /* cache_neighbors_scratch_float4(n) */, VtHashParams hp,
                                  Instancing inst,  // hp.tableSize = rows per instance
                                  const unsigned char* ownedMask = nullptr,  // decomposed mode: lists of the owned particles only ...
                                  unsigned numOwned = 0,                     // ... exactly this many of them
                                  const unsigned* sortedHashForCells = nullptr,  // also run FindCellStart (no launch_find_cell_start then)
                                  unsigned ownedBegin = 0xffffffffu,  // the owned particles are the index range [ownedBegin, +numOwned)
                                  unsigned bandParticles = 0);        // > 0: walk a contiguous range in bands of this many indices
size_t cache_neighbors_scratch_float4(size_t numParticles);
void launch_copy_words(cudaStream_t stream, const void* src, void* dst, size_t words);
void launch_pack_float4(const FusedLaunch& L, const float* packed3, float4* out, unsigned n);
void launch_unpack_float4(const FusedLaunch& L, const float4* in, float* packed3, unsigned n);  // xyz of n float4 -> packed float3
// halo exchange plumbing of the domain-decomposed mode: out[i] = src[ids[i]]  /  dst[ids[i]] = in[i]
void launch_gather_by_id(const FusedLaunch& L, const float4* src, const unsigned* ids, unsigned n, float4* out);
void launch_scatter_by_id(const FusedLaunch& L, const float4* in, const unsigned* ids, unsigned n, float4* dst);

}  // namespace exact_math

namespace fast_math {

// Per-frame inputs the host may have rewritten in managed memory: colliders (prepared once, with
// lastTransform * invCurTransform hoisted) and attach slot positions (copied to device scratch).
void launch_prepare_inputs(const FusedLaunch& L, const VtSDFCollider* colliders, PreparedCollider* prepared,
                           const float* slotPositions, float* slotPositionsOut, unsigned numSlotFloats,
                           const FrameParams* fp);

// AoS import + pre-stabilisation SDF pass (frame dt) + PredictPositions of substep 0.
void launch_begin_frame(const FusedLaunch& L, const float* positions, const float* velocities, const float* invMasses,
                        float4* pos4, float4* pred, const PreparedCollider* colliders,
                        const FrameParams* fp);

// CollideParticles + ApplyDeltas + CollideSDF (substep dt): predIn -> predOut.
void launch_collide(const FusedLaunch& L, const float4* predIn, float4* predOut, const float4* pos4,
                    const unsigned* neighbors, const PreparedCollider* colliders, const FrameParams* fp,
                    bool selfCollision, const unsigned* subset = nullptr, unsigned subsetCount = 0);
// the same for the contiguous particle range [begin, begin + count) (one strip of a decomposed grid cloth)
void launch_collide_range(const FusedLaunch& L, const float4* predIn, float4* predOut, const float4* pos4, const unsigned* neighbors,
                          const PreparedCollider* colliders, const FrameParams* fp, bool selfCollision, unsigned begin, unsigned count);

// One Jacobi iteration: SolveStretch + SolveAttachment + SolveBending + ApplyDeltas, predIn -> predOut.
void launch_iterate(const FusedLaunch& L, const float4* predIn, float4* predOut, const TilePlanDev& plan,
                    const float* attachSlotPositions, const FrameParams* fp, Instancing inst);
size_t iterate_smem_bytes(const TilePlanDev& plan);
// The same iteration for grid cloths without index records (iterate_grid_kernel): bit-identical to launch_iterate.
// iterations > 1 (no strip): that many Jacobi iterations in ONE launch, separated by grid-wide barriers on *gridBarrier; the
// result is in predOut when `iterations` is odd, in predIn when it is even (both arrays are overwritten along the way)
void launch_iterate_grid(const FusedLaunch& L, float4* predIn, float4* predOut, const GridPlanDev& plan,
                         const float* attachSlotPositions, const FrameParams* fp, Instancing inst,
                         const ddpeer::StripArgs* strip = nullptr,  // strip: this rank's rows of a decomposed cloth (dd_peer.cuh)
                         unsigned iterations = 1, unsigned* gridBarrier = nullptr);
unsigned configure_iterate_grid_kernel();  // returns resident CTAs on the device
unsigned configure_iterate_kernel(size_t smemBytes, unsigned threads, unsigned ctaThreads);  // opt in to > 48 KB smem; returns resident CTAs on the device

// Finalize of substep s fused with PredictPositions of substep s+1 (or, on the last substep, with the export
// of positions / velocities / predicted to the public packed-float3 buffers).
void launch_end_substep(const FusedLaunch& L, const float4* predIn, float4* pos4, float4* predNext,
                        bool last, float* positionsOut, float* velocitiesOut, float* predictedOut,
                        const FrameParams* fp);

// ComputeNormal as a per-vertex gather over incident triangles (ascending triangle id).
void launch_normals(const FusedLaunch& L, const float4* pos4, const unsigned* indices, const unsigned* vtxTriOff,
                    const unsigned* vtxTris, float* normalsOut, Instancing inst);

}  // namespace fast_math

// run-time selection of the math build
struct FusedOps {
    decltype(&exact_math::launch_prepare_inputs) prepare_inputs;
    decltype(&exact_math::launch_begin_frame) begin_frame;
    decltype(&exact_math::launch_collide) collide;
    decltype(&exact_math::launch_collide_range) collide_range;
    decltype(&exact_math::launch_iterate) iterate;
    decltype(&exact_math::launch_iterate_grid) iterate_grid;
    decltype(&exact_math::iterate_smem_bytes) iterate_smem_bytes;
    decltype(&exact_math::configure_iterate_kernel) configure_iterate_kernel;
    decltype(&exact_math::launch_end_substep) end_substep;
    decltype(&exact_math::launch_normals) normals;
};
inline FusedOps fused_ops(bool fastMath)
{
    if (fastMath)
        return FusedOps{&fast_math::launch_prepare_inputs, &fast_math::launch_begin_frame, &fast_math::launch_collide, &fast_math::launch_collide_range,
                        &fast_math::launch_iterate, &fast_math::launch_iterate_grid, &fast_math::iterate_smem_bytes, &fast_math::configure_iterate_kernel,
                        &fast_math::launch_end_substep, &fast_math::launch_normals};
    return FusedOps{&exact_math::launch_prepare_inputs, &exact_math::launch_begin_frame, &exact_math::launch_collide, &exact_math::launch_collide_range,
                    &exact_math::launch_iterate, &exact_math::launch_iterate_grid, &exact_math::iterate_smem_bytes, &exact_math::configure_iterate_kernel,
                    &exact_math::launch_end_substep, &exact_math::launch_normals};
}

}  // namespace velvet
