// grid_plan.hpp -- recognises grid cloths among the registered constraints and lays them out for the implicit-grid
// Jacobi kernel (iterate_grid_kernel, fused_kernels.cu).
//
// Every cloth the reference creates is a (R+1) x (R+1) grid whose constraints VtClothObjectGPU generates in a fixed
// pattern (VtClothObjectGPU.hpp L75-132; mesh indices from Scene.hpp L153-165): for vertex (x, y), index x*(R+1)+y,
//   stretch   (x,y)-(x,y+1) if y != R;  (x,y)-(x+1,y) if x != R;  (x,y)-(x+1,y+1) and (x,y+1)-(x+1,y) if both
//   bending   one per quad (x,y): particles (x,y), (x+1,y+1), (x,y+1), (x+1,y), quad id x*R + y.
// When the registered stretch / bending lists are EXACTLY that pattern for every cloth (checked entry by entry; attach
// constraints are free), the constraint topology is implied by the vertex coordinates and the kernel needs no index
// records: per vertex one float4 of rest lengths (vertical, horizontal, diagonal, anti-diagonal of the constraints
// generated there) and one rest angle.  Anything else (extra AddStretch calls, another triangulation, a generic mesh)
// keeps the record-driven tile plan (tile_plan.hpp).
#pragma once

#include <cstddef>
#include <string>
#include <vector>

namespace velvet {

constexpr int GRID_TILE = 15;                 // owned particles per tile side: (15+1)^2 = 256 constraint bundles = 256 threads
// Second tile shape: 14 x 16 owned particles = 15 x 17 = 255 bundles.  A cloth side that 15 divides badly wastes whole
// tile rows (side 64: 5 x 5 tiles of 15 for 4.27 x 4.27 tiles' worth of particles, 73 % of the threads useful; 14 x 16
// covers it with 5 x 4 tiles, 91 %).  Results do not depend on the shape (every particle sums its constraints in id order).
constexpr int GRID_TILE_RX = 14, GRID_TILE_RY = 16;
constexpr unsigned GRID_MAX_CLOTHS = 32;      // cloth table staged in shared memory

struct GridCloth {
    unsigned base;       // first particle
    unsigned side;       // R + 1
    unsigned tilesY;     // tiles along y (the fast index)
    unsigned firstTile;  // prefix sum of tile counts
};

struct GridPlan {
    bool valid = false;
    std::string why;  // when !valid
    std::vector<GridCloth> cloths;
    unsigned numTiles = 0;
    unsigned tileX = GRID_TILE, tileY = GRID_TILE;  // owned particles per tile along x (slow index) and y (fast index)
    std::vector<float> rest4;      // 4 per particle: rest lengths of the stretch constraints generated at that vertex
    std::vector<float> restAngle;  // 1 per particle: rest angle of the quad's bending constraint
    std::vector<unsigned> attOff;  // attach CSR by particle (numParticles + 1), constraint-id order inside a particle
    std::vector<unsigned> attachRec;  // 2 per attach constraint: {slot id, distance bits}
};

struct ClothRange {
    unsigned base, count;
};

// Tile shape for a set of square cloths (one shape per launch): 14 x 16 when it needs at least 5 % fewer tiles than
// 15 x 15, unless `squareOnly` (the strip decomposition of a single cloth counts in 15-row tiles).
void choose_grid_tile_shape(const std::vector<unsigned>& sides, bool squareOnly, unsigned& tileX, unsigned& tileY);
// Fills tilesY / firstTile of every cloth (side and base set) for the shape; returns the number of tiles.
unsigned lay_out_grid_tiles(std::vector<GridCloth>& cloths, unsigned tileX, unsigned tileY);

// `cloths`: the particle ranges of the AddCloth calls, in registration order.
GridPlan build_grid_plan(unsigned numParticles, const std::vector<ClothRange>& cloths, const int* stretchIndices,
                         const float* stretchLengths, size_t numStretch, const unsigned* bendIndices, const float* bendAngles,
                         size_t numBend, const int* attachParticleIDs, const int* attachSlotIDs, const float* attachDistances,
                         size_t numAttach, bool squareTilesOnly = false);

}  // namespace velvet
