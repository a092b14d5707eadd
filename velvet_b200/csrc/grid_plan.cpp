// grid_plan.cpp -- see grid_plan.hpp.
#include "grid_plan.hpp"

#include <cmath>
#include <cstring>

namespace velvet {

void choose_grid_tile_shape(const std::vector<unsigned>& sides, bool squareOnly, unsigned& tileX, unsigned& tileY)
{
    auto tiles = [&](unsigned tx, unsigned ty) {
        unsigned long long n = 0;
        for (unsigned s : sides) n += (unsigned long long)((s + tx - 1) / tx) * ((s + ty - 1) / ty);
        return n;
    };
    tileX = tileY = GRID_TILE;
    if (squareOnly) return;
    const unsigned long long square = tiles(GRID_TILE, GRID_TILE), rect = tiles(GRID_TILE_RX, GRID_TILE_RY);
    if (rect * 100 <= square * 95) {
        tileX = GRID_TILE_RX;
        tileY = GRID_TILE_RY;
    }
}

unsigned lay_out_grid_tiles(std::vector<GridCloth>& cloths, unsigned tileX, unsigned tileY)
{
    unsigned tiles = 0;
    for (GridCloth& gc : cloths) {
        gc.tilesY = (gc.side + tileY - 1) / tileY;
        gc.firstTile = tiles;
        tiles += ((gc.side + tileX - 1) / tileX) * gc.tilesY;
    }
    return tiles;
}

GridPlan build_grid_plan(unsigned numParticles, const std::vector<ClothRange>& cloths, const int* stretchIndices,
                         const float* stretchLengths, size_t numStretch, const unsigned* bendIndices, const float* bendAngles,
                         size_t numBend, const int* attachParticleIDs, const int* attachSlotIDs, const float* attachDistances,
                         size_t numAttach, bool squareTilesOnly)
{
    GridPlan g;
    auto fail = [&](const char* why) {
        g.why = why;
        return g;
    };
    if (cloths.empty()) return fail("no cloth ranges");
    if (cloths.size() > GRID_MAX_CLOTHS) return fail("more cloths than the grid kernel's table holds");
    unsigned covered = 0;
    size_t wantStretch = 0, wantBend = 0;
    for (const ClothRange& c : cloths) {
        if (c.base != covered) return fail("cloth ranges are not contiguous");
        const unsigned side = (unsigned)std::lround(std::sqrt((double)c.count));
        if (side < 2 || (size_t)side * side != c.count) return fail("a cloth is not a square grid");
        const size_t R = side - 1;
        wantStretch += 4 * R * R + 2 * R;
        wantBend += R * R;
        covered += c.count;
    }
    if (covered != numParticles) return fail("particles outside the registered cloths");
    if (wantStretch != numStretch || wantBend != numBend) return fail("constraint counts differ from the grid pattern");

    g.rest4.assign(4 * (size_t)numParticles, 0.0f);
    g.restAngle.assign(numParticles, 0.0f);
    size_t s = 0, b = 0;
    for (const ClothRange& c : cloths) {
        const unsigned side = (unsigned)std::lround(std::sqrt((double)c.count));
        const int R = (int)side - 1, off = (int)c.base;
        auto at = [&](int x, int y) { return off + x * (int)side + y; };
        auto expect = [&](int a, int bIdx, int vertex, int kind) {
            if (stretchIndices[2 * s] != a || stretchIndices[2 * s + 1] != bIdx) return false;
            g.rest4[4 * (size_t)vertex + kind] = stretchLengths[s];
            s++;
            return true;
        };
        for (int x = 0; x <= R; x++)
            for (int y = 0; y <= R; y++) {
                const int v = at(x, y);
                if (y != R && !expect(v, at(x, y + 1), v, 0)) return fail("stretch constraints differ from the grid pattern");
                if (x != R && !expect(v, at(x + 1, y), v, 1)) return fail("stretch constraints differ from the grid pattern");
                if (y != R && x != R) {
                    if (!expect(v, at(x + 1, y + 1), v, 2)) return fail("stretch constraints differ from the grid pattern");
                    if (!expect(at(x, y + 1), at(x + 1, y), v, 3)) return fail("stretch constraints differ from the grid pattern");
                }
            }
        for (int x = 0; x < R; x++)
            for (int y = 0; y < R; y++, b++) {
                const unsigned* q = bendIndices + 4 * b;
                if (q[0] != (unsigned)at(x, y) || q[1] != (unsigned)at(x + 1, y + 1) || q[2] != (unsigned)at(x, y + 1) ||
                    q[3] != (unsigned)at(x + 1, y))
                    return fail("bending constraints differ from the grid pattern");
                g.restAngle[(size_t)at(x, y)] = bendAngles[b];
            }
        GridCloth gc;
        gc.base = c.base;
        gc.side = side;
        gc.tilesY = gc.firstTile = 0;
        g.cloths.push_back(gc);
    }
    {
        std::vector<unsigned> sides;
        for (const GridCloth& gc : g.cloths) sides.push_back(gc.side);
        choose_grid_tile_shape(sides, squareTilesOnly, g.tileX, g.tileY);
        g.numTiles = lay_out_grid_tiles(g.cloths, g.tileX, g.tileY);
    }

    // attach constraints by particle, ascending constraint id inside a particle (counting sort)
    g.attOff.assign((size_t)numParticles + 1, 0u);
    for (size_t a = 0; a < numAttach; a++) {
        if (attachParticleIDs[a] < 0 || (unsigned)attachParticleIDs[a] >= numParticles) return fail("attach particle index out of range");
        g.attOff[(size_t)attachParticleIDs[a] + 1]++;
    }
    for (unsigned p = 0; p < numParticles; p++) g.attOff[p + 1] += g.attOff[p];
    g.attachRec.assign(2 * numAttach + 2, 0u);
    std::vector<unsigned> cursor(g.attOff.begin(), g.attOff.end() - 1);
    for (size_t a = 0; a < numAttach; a++) {
        const unsigned at = cursor[(size_t)attachParticleIDs[a]]++;
        unsigned bits;
        std::memcpy(&bits, &attachDistances[a], 4);
        g.attachRec[2 * (size_t)at] = (unsigned)attachSlotIDs[a];
        g.attachRec[2 * (size_t)at + 1] = bits;
    }
    g.valid = true;
    return g;
}

}  // namespace velvet
