// tile_plan.hpp -- host preprocessing for the tile-fused Jacobi iteration kernel.
//
// The reference runs one kernel per constraint type and scatters every correction with 3 float + 1 int
// global atomics (VtClothSolverGPU.cu L65-251), then a fourth kernel averages (L253-264).  Here particles are
// partitioned into spatially compact tiles (Morton order of the registration-time positions); every
// constraint is assigned to the tile(s) owning one of its particles; inside a tile each (constraint,
// endpoint) pair gets a private shared-memory slot and each particle sums its slots in ascending
// constraint-id order (stretch, then attach, then bend).  That order is exactly the sequential order of
// the CPU oracle, so the result is deterministic and, up to libm, bit-identical to it.
//
// Constraints that straddle tiles are evaluated by each owning tile (same inputs, same code, same bits);
// only owned endpoints receive a slot.  Particles referenced but not owned form the tile's halo.
// Slot of (particle local index l, ordinal k) is slots[k * tileSize + l]: consecutive particles read consecutive
// 16-byte words, so the per-particle sums are free of shared-memory bank conflicts.
//
// The ORDER of a tile's records is free (a slot is addressed by ordinal, not by record position), and it decides the bank
// conflicts of the constraint threads: a warp-wide 16-byte access is served in quarter-warps of 8 lanes, conflict-free when
// the 8 local indices differ mod 8.  In constraint-id order the 4 stretch constraints generated per grid vertex sit in
// adjacent lanes and share an endpoint: their slot stores collide 4-way (ncu, round 1: 2.7x the ideal store wavefronts, the
// LSU pipe 57 % busy).  build_tile_plan therefore emits the records of a tile in a greedily conflict-avoiding order.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

namespace velvet {

struct TileDesc {
    unsigned ownedOff, nOwned;      // range in ownedIds
    unsigned haloOff, nHalo;        // range in haloIds
    unsigned stretchOff, nStretch;  // range in stretchRec
    unsigned bendOff, nBend;        // range in bendRec
    unsigned baseOff;               // first of nOwned+1 entries in attOff
    unsigned attachOff, nAttach;    // range in attachRec
    unsigned pad;
};

struct Rec2 {
    unsigned x, y;
};
struct Rec4 {
    unsigned x, y, z, w;
};

// endpoint encoding inside a record: (localIndex << 5) | slotOrdinal.  Halo endpoints (no slot of their own) carry the
// ordinal of a dump row (= maxKS for stretch, maxKB for bend) so that kernels can store unconditionally.
constexpr unsigned TP_ORD_BITS = 5;
constexpr unsigned TP_NO_SLOT = 31;
constexpr unsigned TP_MAX_LOCALS = 2047;

struct TilePlan {
    bool valid = false;
    std::string whyInvalid;
    int tileSize = 0;
    std::vector<TileDesc> tiles;
    std::vector<unsigned> ownedIds, haloIds;
    std::vector<uint8_t> sCnt, bCnt;     // stretch / bend constraints incident to each owned particle (index = ownedOff + local)
    std::vector<unsigned> attOff;        // attach CSR per owned particle (+1 per tile), relative to attachOff
    std::vector<Rec2> stretchRec;        // {ea | eb << 16, restLength bits}; every tile's range starts at an even index
    std::vector<Rec4> bendRec;           // {e0 | e1 << 16, e2 | e3 << 16, restAngle bits, constraint id}
    std::vector<Rec2> attachRec;         // {slot id, distance bits}
    unsigned maxLocals = 0;              // tileSize + max over tiles of nHalo (halo locals start at tileSize)
    unsigned maxBendPerTile = 0;         // max over tiles of nBend (bend records are staged in shared memory)
    unsigned maxStretchPerTile = 0;      // max over tiles of nStretch (staged in shared memory as well)
    unsigned maxKS = 0, maxKB = 0;       // max stretch / bend constraints on one particle; slot rows are [k][local], plus a dump row
    // statistics
    size_t numStretchEvaluated = 0, numBendEvaluated = 0, numHalo = 0;
    // shared-memory wavefronts of the constraint threads' 16-byte accesses (position loads + slot stores) per Jacobi
    // iteration: the minimum (4 per warp-wide access), with records in constraint-id order, and in the emitted order
    size_t smemWavefrontsIdeal = 0, smemWavefrontsIdOrder = 0, smemWavefronts = 0;
};

// positions: packed float3 (host) used only to order particles spatially.
TilePlan build_tile_plan(unsigned numParticles, const float* positions, const int* stretchIndices,
                         const float* stretchLengths, size_t numStretch, const unsigned* bendIndices,
                         const float* bendAngles, size_t numBend, const int* attachParticleIDs,
                         const int* attachSlotIDs, const float* attachDistances, size_t numAttach, int tileSize);

// ---- domain decomposition of ONE cloth over `world` ranks (north_star mode 2): rank r owns the contiguous tile range
// [tileBegin(r), tileEnd(r)) of the Morton-ordered plan, i.e. a spatially compact patch.  Before every Jacobi iteration a
// rank needs the predicted positions of the halo particles of its tiles that another rank owns.
struct ExchangePlan {
    int rank = 0, world = 1;
    unsigned tileBegin = 0, tileEnd = 0;
    std::vector<unsigned> tileBeginOf;              // [world + 1] tile range of every rank
    std::vector<std::vector<unsigned>> recvIds;     // [world] particle ids owned by peer q that this rank reads (ascending)
    std::vector<std::vector<unsigned>> sendIds;     // [world] particle ids owned by this rank that peer q reads (ascending)
};
ExchangePlan build_exchange_plan(const TilePlan& plan, unsigned numParticles, int rank, int world);

}  // namespace velvet
