// tile_plan.cpp -- see tile_plan.hpp.
#include "tile_plan.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>

namespace velvet {

namespace {

inline uint64_t spread21(uint64_t v)
{  // 21 bits -> every third bit
    v &= 0x1fffffull;
    v = (v | v << 32) & 0x1f00000000ffffull;
    v = (v | v << 16) & 0x1f0000ff0000ffull;
    v = (v | v << 8) & 0x100f00f00f00f00full;
    v = (v | v << 4) & 0x10c30c30c30c30c3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}

inline unsigned float_bits(float f)
{
    unsigned u;
    std::memcpy(&u, &f, 4);
    return u;
}

// ---- bank-conflict-avoiding record order (see tile_plan.hpp).  A record has K 16-bit endpoints (local << 5 | ordinal,
// ordinal == TP_NO_SLOT for halo endpoints).  Lane l of a warp handles record (base + l): records [8g, 8g+8) of a tile form
// a quarter-warp.  Endpoint k of those 8 records is one 16-byte shared-memory access (position load sp[local]; slot store
// slots[ordinal][local] for owned endpoints): its bank group is local mod 8.
template <int K>
struct EndpointsOf;
template <>
struct EndpointsOf<2> {
    static void get(const Rec2& r, unsigned* e) { e[0] = r.x & 0xffffu; e[1] = r.x >> 16; }
};
template <>
struct EndpointsOf<4> {
    static void get(const Rec4& r, unsigned* e) { e[0] = r.x & 0xffffu; e[1] = r.x >> 16; e[2] = r.y & 0xffffu; e[3] = r.y >> 16; }
};

// wavefronts of the loads + stores of one quarter-warp (8 records or fewer)
template <int K, class Rec>
unsigned quarter_wavefronts(const Rec* recs, unsigned n)
{
    unsigned total = 0;
    for (int k = 0; k < K; k++) {
        unsigned loadLocals[8][8], loadCnt[8] = {0}, storeCnt[8] = {0};
        for (unsigned i = 0; i < n; i++) {
            unsigned e[K];
            EndpointsOf<K>::get(recs[i], e);
            const unsigned local = e[k] >> TP_ORD_BITS, cls = local & 7u;
            bool seen = false;  // lanes reading the same address are served together
            for (unsigned j = 0; j < loadCnt[cls]; j++) seen |= loadLocals[cls][j] == local;
            if (!seen) loadLocals[cls][loadCnt[cls]++] = local;
            if ((e[k] & 31u) != TP_NO_SLOT) storeCnt[cls]++;
        }
        unsigned l = 0, st = 0;
        for (int c = 0; c < 8; c++) {
            l = std::max(l, loadCnt[c]);
            st = std::max(st, storeCnt[c]);
        }
        total += l + st;
    }
    return total;
}

template <int K, class Rec>
size_t tile_wavefronts(const Rec* recs, unsigned n)
{
    size_t total = 0;
    for (unsigned b = 0; b < n; b += 8) total += quarter_wavefronts<K>(recs + b, std::min(8u, n - b));
    return total;
}

template <int K, class Rec>
void reorder_for_banks(Rec* recs, unsigned n, std::vector<Rec>& scratch, std::vector<unsigned char>& taken)
{
    constexpr unsigned WINDOW = 128;
    scratch.assign(recs, recs + n);
    taken.assign(n, 0);
    unsigned head = 0, out = 0;
    while (out < n) {
        // owner[k][cls]: 0 = free, else local + 1 of the endpoint holding the bank group; halo endpoints of the same local
        // may share (one address, no store)
        unsigned owner[K][8] = {};
        bool ownedSlot[K][8] = {};
        unsigned cnt = 0;
        auto fits = [&](const Rec& r) {
            unsigned e[K];
            EndpointsOf<K>::get(r, e);
            for (int k = 0; k < K; k++) {
                const unsigned local = e[k] >> TP_ORD_BITS, cls = local & 7u;
                const bool halo = (e[k] & 31u) == TP_NO_SLOT;
                if (owner[k][cls] && !(halo && !ownedSlot[k][cls] && owner[k][cls] == local + 1)) return false;
            }
            return true;
        };
        auto take = [&](unsigned i) {
            unsigned e[K];
            EndpointsOf<K>::get(scratch[i], e);
            for (int k = 0; k < K; k++) {
                const unsigned local = e[k] >> TP_ORD_BITS, cls = local & 7u;
                owner[k][cls] = local + 1;
                ownedSlot[k][cls] = ownedSlot[k][cls] || (e[k] & 31u) != TP_NO_SLOT;
            }
            taken[i] = 1;
            recs[out++] = scratch[i];
            cnt++;
        };
        const unsigned want = std::min(8u, n - out);
        for (unsigned i = head; i < n && i < head + WINDOW && cnt < want; i++)
            if (!taken[i] && fits(scratch[i])) take(i);
        for (unsigned i = head; i < n && cnt < want; i++)  // nothing conflict-free left in the window: fill in id order
            if (!taken[i]) take(i);
        while (head < n && taken[head]) head++;
    }
}

}  // namespace

TilePlan build_tile_plan(unsigned N, const float* positions, const int* stretchIndices, const float* stretchLengths,
                         size_t S, const unsigned* bendIndices, const float* bendAngles, size_t B,
                         const int* attachParticleIDs, const int* attachSlotIDs, const float* attachDistances, size_t A,
                         int tileSize)
{
    TilePlan plan;
    plan.tileSize = tileSize;
    auto fail = [&](const std::string& why) {
        plan.valid = false;
        plan.whyInvalid = why;
        return plan;
    };
    if (N == 0) return fail("no particles");
    if (tileSize < 32 || tileSize > 1024 || (tileSize % 32) != 0) return fail("tile size must be a multiple of 32 in [32,1024]");

    // validate indices (the reference would read out of bounds; we refuse the fused plan instead)
    for (size_t c = 0; c < S; c++)
        if ((unsigned)stretchIndices[2 * c] >= N || (unsigned)stretchIndices[2 * c + 1] >= N) return fail("stretch index out of range");
    for (size_t c = 0; c < 4 * B; c++)
        if (bendIndices[c] >= N) return fail("bend index out of range");
    for (size_t c = 0; c < A; c++)
        if ((unsigned)attachParticleIDs[c] >= N) return fail("attach particle index out of range");

    // ---- 1. spatial order: Morton code of the quantised positions, ties by particle id
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (unsigned i = 0; i < N; i++)
        for (int k = 0; k < 3; k++) {
            const float v = positions[3 * (size_t)i + k];
            if (std::isfinite(v)) {
                lo[k] = std::min(lo[k], v);
                hi[k] = std::max(hi[k], v);
            }
        }
    float extent = 0;
    for (int k = 0; k < 3; k++)
        if (hi[k] >= lo[k]) extent = std::max(extent, hi[k] - lo[k]);
    const double scale = extent > 0 ? 2097151.0 / (double)extent : 0.0;
    std::vector<std::pair<uint64_t, unsigned>> order(N);
    for (unsigned i = 0; i < N; i++) {
        uint64_t q[3];
        for (int k = 0; k < 3; k++) {
            const float v = positions[3 * (size_t)i + k];
            double t = std::isfinite(v) ? ((double)v - (double)lo[k]) * scale : 0.0;
            if (t < 0) t = 0;
            if (t > 2097151.0) t = 2097151.0;
            q[k] = (uint64_t)t;
        }
        order[i] = {spread21(q[0]) | (spread21(q[1]) << 1) | (spread21(q[2]) << 2), i};
    }
    std::sort(order.begin(), order.end());

    const unsigned T = (unsigned)tileSize;
    const unsigned numTiles = (N + T - 1) / T;
    std::vector<unsigned> tileOf(N);
    plan.ownedIds.resize(N);
    for (unsigned r = 0; r < N; r++) {
        tileOf[order[r].second] = r / T;
        plan.ownedIds[r] = order[r].second;
    }
    order.clear();
    order.shrink_to_fit();
    for (unsigned t = 0; t < numTiles; t++) {
        const unsigned b = t * T, e = std::min(N, b + T);
        std::sort(plan.ownedIds.begin() + b, plan.ownedIds.begin() + e);  // ascending ids => coalesced loads
    }
    std::vector<unsigned> localOf(N);
    for (unsigned r = 0; r < N; r++) localOf[plan.ownedIds[r]] = r % T;

    // ---- 2. per-tile constraint lists (CSR, ascending constraint id inside each tile)
    auto distinct_tiles = [&](const unsigned* ids, int n, unsigned* out) {
        int m = 0;
        for (int i = 0; i < n; i++) {
            const unsigned t = tileOf[ids[i]];
            bool seen = false;
            for (int j = 0; j < m; j++) seen |= (out[j] == t);
            if (!seen) out[m++] = t;
        }
        return m;
    };
    std::vector<size_t> sOff(numTiles + 1, 0), bOff(numTiles + 1, 0), aOff(numTiles + 1, 0);
    unsigned tl[4];
    for (size_t c = 0; c < S; c++) {
        const unsigned ids[2] = {(unsigned)stretchIndices[2 * c], (unsigned)stretchIndices[2 * c + 1]};
        const int m = distinct_tiles(ids, 2, tl);
        for (int j = 0; j < m; j++) sOff[tl[j] + 1]++;
    }
    for (size_t c = 0; c < B; c++) {
        const int m = distinct_tiles(bendIndices + 4 * c, 4, tl);
        for (int j = 0; j < m; j++) bOff[tl[j] + 1]++;
    }
    for (size_t c = 0; c < A; c++) aOff[tileOf[(unsigned)attachParticleIDs[c]] + 1]++;
    for (unsigned t = 0; t < numTiles; t++) {
        sOff[t + 1] += sOff[t];
        bOff[t + 1] += bOff[t];
        aOff[t + 1] += aOff[t];
    }
    std::vector<unsigned> sList(sOff[numTiles]), bList(bOff[numTiles]), aList(aOff[numTiles]);
    {
        std::vector<size_t> sc(sOff.begin(), sOff.end() - 1), bc(bOff.begin(), bOff.end() - 1), ac(aOff.begin(), aOff.end() - 1);
        for (size_t c = 0; c < S; c++) {
            const unsigned ids[2] = {(unsigned)stretchIndices[2 * c], (unsigned)stretchIndices[2 * c + 1]};
            const int m = distinct_tiles(ids, 2, tl);
            for (int j = 0; j < m; j++) sList[sc[tl[j]]++] = (unsigned)c;
        }
        for (size_t c = 0; c < B; c++) {
            const int m = distinct_tiles(bendIndices + 4 * c, 4, tl);
            for (int j = 0; j < m; j++) bList[bc[tl[j]]++] = (unsigned)c;
        }
        for (size_t c = 0; c < A; c++) aList[ac[tileOf[(unsigned)attachParticleIDs[c]]]++] = (unsigned)c;
    }

    // ---- 3. per tile: halo discovery, slot ordinals, records
    plan.tiles.resize(numTiles);
    // stretch records of a tile start on a 16-byte boundary (pairs of 8-byte records are staged with 16-byte cp.async)
    std::vector<size_t> sRecOff(numTiles + 1, 0);
    for (unsigned t = 0; t < numTiles; t++) sRecOff[t + 1] = (sRecOff[t] + (sOff[t + 1] - sOff[t]) + 1) & ~(size_t)1;
    plan.stretchRec.assign(sRecOff[numTiles], Rec2{0, 0});
    plan.bendRec.resize(bList.size());
    plan.attachRec.resize(aList.size());
    plan.sCnt.resize(N);
    plan.bCnt.resize(N);
    plan.attOff.resize((size_t)N + numTiles);
    // Tiles are independent once the per-tile constraint lists exist: worker threads take contiguous tile ranges.  Every
    // output range is disjoint per tile; the halo lists are collected per worker (discovery order inside a tile is the
    // sequential order) and concatenated in tile order afterwards, so the plan does not depend on the thread count.
    struct Worker {
        std::vector<unsigned> halo;                 // halo ids of the worker's tiles, tile after tile
        std::vector<unsigned> haloKey, haloVal;     // open-addressing map particle -> halo local of the current tile
        std::vector<unsigned> cntS, cntB, cntA;
        std::vector<Rec2> scratch2;
        std::vector<Rec4> scratch4;
        std::vector<unsigned char> taken;
        unsigned maxKS = 0, maxKB = 0, maxLocals = 0, maxBend = 0, maxStretch = 0;
        size_t wfIdeal = 0, wfIdOrder = 0, wfEmitted = 0;
        std::string error;
    };
    unsigned numWorkers = std::max(1u, std::min(std::thread::hardware_concurrency(), 16u));
    if (const char* e = getenv("VELVET_PLAN_THREADS")) numWorkers = (unsigned)std::max(1, atoi(e));
    numWorkers = std::min(numWorkers, std::max(1u, numTiles / 64u));  // small plans: one thread
    std::vector<Worker> workers(numWorkers);
    std::vector<unsigned> haloCount(numTiles, 0);

    auto run = [&](unsigned wi) {
        Worker& W = workers[wi];
        const unsigned tBegin = (unsigned)((unsigned long long)numTiles * wi / numWorkers);
        const unsigned tEnd = (unsigned)((unsigned long long)numTiles * (wi + 1) / numWorkers);
        constexpr unsigned MAP = 4096, EMPTY = 0xffffffffu;  // > 2 x TP_MAX_LOCALS entries
        W.haloKey.assign(MAP, EMPTY);
        W.haloVal.resize(MAP);
        W.cntS.resize(T);
        W.cntB.resize(T);
        W.cntA.resize(T);
        std::vector<unsigned> usedSlots;
        for (unsigned t = tBegin; t < tEnd; t++) {
            TileDesc& td = plan.tiles[t];
            td.ownedOff = t * T;
            td.nOwned = std::min(N, td.ownedOff + T) - td.ownedOff;
            td.haloOff = 0;  // filled in after the join
            td.nHalo = 0;
            td.stretchOff = (unsigned)sRecOff[t];
            td.nStretch = (unsigned)(sOff[t + 1] - sOff[t]);
            td.bendOff = (unsigned)bOff[t];
            td.nBend = (unsigned)(bOff[t + 1] - bOff[t]);
            td.attachOff = (unsigned)aOff[t];
            td.nAttach = (unsigned)(aOff[t + 1] - aOff[t]);
            td.baseOff = td.ownedOff + t;
            td.pad = 0;
            std::fill(W.cntS.begin(), W.cntS.end(), 0u);
            std::fill(W.cntB.begin(), W.cntB.end(), 0u);
            std::fill(W.cntA.begin(), W.cntA.end(), 0u);
            for (unsigned slot : usedSlots) W.haloKey[slot] = EMPTY;
            usedSlots.clear();

            bool overflow = false, haloOverflow = false;
            auto endpoint = [&](unsigned p, std::vector<unsigned>& cnt) -> unsigned {
                if (tileOf[p] == t) {
                    const unsigned l = localOf[p];
                    const unsigned k = cnt[l]++;
                    if (k >= TP_NO_SLOT) overflow = true;
                    return (l << TP_ORD_BITS) | (k & 31u);
                }
                unsigned slot = (p * 2654435761u) >> 20;  // 12 bits
                while (W.haloKey[slot] != EMPTY && W.haloKey[slot] != p) slot = (slot + 1) & (MAP - 1);
                if (W.haloKey[slot] == EMPTY) {
                    if (T + td.nHalo >= TP_MAX_LOCALS) {  // reported below; keep the map from filling up
                        haloOverflow = true;
                        return (TP_MAX_LOCALS << TP_ORD_BITS) | TP_NO_SLOT;
                    }
                    W.haloKey[slot] = p;
                    W.haloVal[slot] = T + td.nHalo++;  // halo locals start at the tile size, also in a partially filled tile
                    usedSlots.push_back(slot);
                    W.halo.push_back(p);
                }
                return (W.haloVal[slot] << TP_ORD_BITS) | TP_NO_SLOT;
            };

            for (unsigned i = 0; i < td.nStretch; i++) {
                const unsigned c = sList[sOff[t] + i];
                const unsigned ea = endpoint((unsigned)stretchIndices[2 * (size_t)c], W.cntS);
                const unsigned eb = endpoint((unsigned)stretchIndices[2 * (size_t)c + 1], W.cntS);
                plan.stretchRec[td.stretchOff + i] = Rec2{ea | (eb << 16), float_bits(stretchLengths[c])};
            }
            for (unsigned i = 0; i < td.nBend; i++) {
                const unsigned c = bList[td.bendOff + i];
                unsigned e[4];
                for (int k = 0; k < 4; k++) e[k] = endpoint(bendIndices[4 * (size_t)c + k], W.cntB);
                plan.bendRec[td.bendOff + i] = Rec4{e[0] | (e[1] << 16), e[2] | (e[3] << 16), float_bits(bendAngles[c]), c};
            }
            haloCount[t] = td.nHalo;
            if (overflow) {
                W.error = "a particle has more than 30 stretch or bend constraints";
                return;
            }
            if (haloOverflow || T + td.nHalo > TP_MAX_LOCALS) {
                W.error = "tile halo too large (more than 2047 local particles)";
                return;
            }

            // record order inside the tile: free of shared-memory bank conflicts where possible (ordinals are already fixed)
            {
                Rec2* sr = plan.stretchRec.data() + td.stretchOff;
                Rec4* br = plan.bendRec.data() + td.bendOff;
                W.wfIdeal += (size_t)((td.nStretch + 7) / 8) * 4 + (size_t)((td.nBend + 7) / 8) * 8;
                W.wfIdOrder += tile_wavefronts<2>(sr, td.nStretch) + tile_wavefronts<4>(br, td.nBend);
                reorder_for_banks<2>(sr, td.nStretch, W.scratch2, W.taken);
                reorder_for_banks<4>(br, td.nBend, W.scratch4, W.taken);
                W.wfEmitted += tile_wavefronts<2>(sr, td.nStretch) + tile_wavefronts<4>(br, td.nBend);
            }

            // per-particle constraint counts (the slot of ordinal k of particle l is slots[k * T + l])
            for (unsigned l = 0; l < td.nOwned; l++) {
                plan.sCnt[td.ownedOff + l] = (uint8_t)W.cntS[l];
                plan.bCnt[td.ownedOff + l] = (uint8_t)W.cntB[l];
                W.maxKS = std::max(W.maxKS, W.cntS[l]);
                W.maxKB = std::max(W.maxKB, W.cntB[l]);
            }
            W.maxLocals = std::max(W.maxLocals, T + td.nHalo);
            W.maxBend = std::max(W.maxBend, td.nBend);
            W.maxStretch = std::max(W.maxStretch, td.nStretch);

            // attach: CSR by owned particle, ascending constraint id inside each particle
            for (unsigned i = 0; i < td.nAttach; i++) W.cntA[localOf[(unsigned)attachParticleIDs[aList[td.attachOff + i]]]]++;
            unsigned accA = 0;
            for (unsigned l = 0; l < td.nOwned; l++) {
                plan.attOff[td.baseOff + l] = accA;
                accA += W.cntA[l];
                W.cntA[l] = plan.attOff[td.baseOff + l];
            }
            plan.attOff[td.baseOff + td.nOwned] = accA;
            for (unsigned i = 0; i < td.nAttach; i++) {
                const unsigned c = aList[td.attachOff + i];
                const unsigned l = localOf[(unsigned)attachParticleIDs[c]];
                plan.attachRec[td.attachOff + W.cntA[l]++] = Rec2{(unsigned)attachSlotIDs[c], float_bits(attachDistances[c])};
            }
        }
    };
    if (numWorkers == 1) {
        run(0);
    } else {
        std::vector<std::thread> pool;
        for (unsigned wi = 0; wi < numWorkers; wi++) pool.emplace_back(run, wi);
        for (std::thread& th : pool) th.join();
    }
    for (const Worker& W : workers)
        if (!W.error.empty()) return fail(W.error);  // workers are scanned in tile order: the first failing range wins
    plan.haloIds.clear();
    {
        size_t total = 0;
        for (const Worker& W : workers) total += W.halo.size();
        plan.haloIds.reserve(total + 1);
        unsigned off = 0;
        for (unsigned t = 0; t < numTiles; t++) {
            plan.tiles[t].haloOff = off;
            off += haloCount[t];
        }
        for (const Worker& W : workers) {
            plan.haloIds.insert(plan.haloIds.end(), W.halo.begin(), W.halo.end());
            plan.maxKS = std::max(plan.maxKS, W.maxKS);
            plan.maxKB = std::max(plan.maxKB, W.maxKB);
            plan.maxLocals = std::max(plan.maxLocals, W.maxLocals);
            plan.maxBendPerTile = std::max(plan.maxBendPerTile, W.maxBend);
            plan.maxStretchPerTile = std::max(plan.maxStretchPerTile, W.maxStretch);
            plan.smemWavefrontsIdeal += W.wfIdeal;
            plan.smemWavefrontsIdOrder += W.wfIdOrder;
            plan.smemWavefronts += W.wfEmitted;
        }
    }
    // halo endpoints: replace the NO_SLOT marker by the dump-row ordinal of their constraint type
    if (plan.maxKS >= TP_NO_SLOT || plan.maxKB >= TP_NO_SLOT) return fail("a particle has more than 30 stretch or bend constraints");
    auto redirect = [](unsigned e, unsigned dumpRow) { return (e & 31u) == TP_NO_SLOT ? ((e & ~31u) | dumpRow) : e; };
    for (Rec2& r : plan.stretchRec)
        r.x = redirect(r.x & 0xffffu, plan.maxKS) | (redirect(r.x >> 16, plan.maxKS) << 16);
    for (Rec4& r : plan.bendRec) {
        r.x = redirect(r.x & 0xffffu, plan.maxKB) | (redirect(r.x >> 16, plan.maxKB) << 16);
        r.y = redirect(r.y & 0xffffu, plan.maxKB) | (redirect(r.y >> 16, plan.maxKB) << 16);
    }
    plan.numStretchEvaluated = sList.size();
    plan.numBendEvaluated = bList.size();
    plan.numHalo = plan.haloIds.size();
    if (plan.haloIds.empty()) plan.haloIds.push_back(0);  // keep device arrays non-empty
    plan.valid = true;
    return plan;
}

ExchangePlan build_exchange_plan(const TilePlan& plan, unsigned N, int rank, int world)
{
    ExchangePlan x;
    x.rank = rank;
    x.world = world;
    const unsigned numTiles = (unsigned)plan.tiles.size();
    x.tileBeginOf.resize((size_t)world + 1);
    for (int r = 0; r <= world; r++) x.tileBeginOf[r] = (unsigned)((unsigned long long)numTiles * r / world);
    x.tileBegin = x.tileBeginOf[rank];
    x.tileEnd = x.tileBeginOf[rank + 1];
    x.recvIds.assign(world, {});
    x.sendIds.assign(world, {});

    // owner rank of every particle: its tile's rank
    std::vector<int> rankOfTile(numTiles);
    for (int r = 0; r < world; r++)
        for (unsigned t = x.tileBeginOf[r]; t < x.tileBeginOf[r + 1]; t++) rankOfTile[t] = r;
    std::vector<int> ownerRank(N);
    for (unsigned t = 0; t < numTiles; t++)
        for (unsigned i = 0; i < plan.tiles[t].nOwned; i++) ownerRank[plan.ownedIds[plan.tiles[t].ownedOff + i]] = rankOfTile[t];

    // needs(r) = halo particles of r's tiles owned elsewhere; every rank evaluates every r, so both sides of each
    // message agree on its contents and order (ascending particle id) without any negotiation
    std::vector<unsigned char> seen(N);
    for (int r = 0; r < world; r++) {
        std::fill(seen.begin(), seen.end(), 0);
        std::vector<std::vector<unsigned>> need(world);
        for (unsigned t = x.tileBeginOf[r]; t < x.tileBeginOf[r + 1]; t++) {
            const TileDesc& td = plan.tiles[t];
            for (unsigned i = 0; i < td.nHalo; i++) {
                const unsigned p = plan.haloIds[td.haloOff + i];
                const int q = ownerRank[p];
                if (q != r && !seen[p]) {
                    seen[p] = 1;
                    need[q].push_back(p);
                }
            }
        }
        for (int q = 0; q < world; q++) std::sort(need[q].begin(), need[q].end());
        if (r == rank) x.recvIds = need;
        else x.sendIds[r] = need[rank];
    }
    return x;
}

}  // namespace velvet
