// setup_kernels.cuh -- registration of a grid cloth on the device.
//
// Reference: VtClothObjectGPU::Start (VtClothObjectGPU.hpp L43-148) builds the stretch / attach / bending lists with one
// host call per constraint, and VtClothSolverGPU::AddCloth (VtClothSolverGPU.hpp L112-150) fills the per-particle arrays
// on the host.  Here the same lists, in the same order and with the same bits, are written by kernels straight into the
// device-resident pages of the public (managed) buffers, and the implicit-grid plan, the vertex -> triangle lists and the
// index checks are derived from them on the device as well: registration costs two host -> device copies (vertices, mesh
// indices) plus a dozen sub-millisecond kernels instead of ~0.2 s of host loops per million particles.
//
// Arithmetic: rest lengths are length_plain(world[a] - world[b]) of the world-space positions the device transform
// (seam::InitializePositions) just produced -- the same mul_point / length_plain the host generator applies to the same
// inputs, compiled without FMA contraction on both sides, so the lists are bit-identical to GenerateGridConstraints'
// (tests/test_setup_gpu.py compares every list).
#pragma once

#include <cuda_runtime.h>

#include <cstddef>

#include "vt_math.cuh"

namespace velvet {
namespace setup {

// 32-bit fills (zeroed velocities, unit inverse masses, ...)
void fill_words(void* dst, size_t words, unsigned value, cudaStream_t st);
// dst[i] = src[i] + offset (mesh indices of a cloth shifted to its place in the particle arrays)
void offset_indices(unsigned* dst, const unsigned* src, size_t n, unsigned offset, cudaStream_t st);

// Batched instances: positions[k * n + i] = model_k * vertices[i] for instance k (models: 16 floats per instance, device),
// and invMasses[k * n + pinned[j]] = 0 for the particles an attach constraint pins
void instance_positions(float* positions, const float* vertices, const float* models16, unsigned n, unsigned numInstances, cudaStream_t st);
void instance_pin(float* invMasses, const int* pinned, unsigned numPinned, unsigned n, unsigned numInstances, cudaStream_t st);

// GenerateStretch (L75-116): 4R^2 + 2R constraints of the cloth whose particles start at `base` (global index), emitted
// vertex by vertex in (x, y) order: structural y, structural x, the two shear diagonals.
void generate_stretch(int* idxPairs, float* lengths, const float* worldPositions, unsigned base, int R, cudaStream_t st);
// GenerateAttach (L134-148), one slot: every particle of the cloth, in index order; a particle at distance 0 from the slot
// is the attached one and loses its inverse mass (VtClothSolverGPU.hpp L172)
void generate_attach(int* particleIds, int* slotIds, float* distances, const float* worldPositions, float* invMasses, unsigned base,
                     unsigned numVertices, int slotId, vec3 slotPosition, cudaStream_t st);
// GenerateBending (L118-132): one dihedral per quad from the (already shifted) mesh indices (i, i+5, i+2, i+1), rest angle 0
void generate_bend(unsigned* idxQuads, float* angles, const unsigned* shiftedMeshIndices, size_t numQuads, cudaStream_t st);

// Implicit-grid plan (grid_plan.hpp) of a generated cloth: rest4[v] = lengths of the stretch constraints generated at
// vertex v (0 where the pattern has none), and *mismatch |= 1 unless the bending quads are the grid's
// ((x,y), (x+1,y+1), (x,y+1), (x+1,y) for quad x*R + y).
void grid_plan_from_lists(float4* rest4, const float* clothStretchLengths, const unsigned* clothBendIndices, unsigned base, int R,
                          int* mismatch, cudaStream_t st);
// Attach records of a generated cloth with `numSlots` slots registered slot by slot: particle p holds records
// [attBase + (p - base) * numSlots, + numSlots) = {firstSlot + s, distance bits}, which is constraint-id order.
void grid_attach_records(unsigned* attOff, uint2* attachRec, const float* clothAttachDistances, unsigned base, unsigned numVertices,
                         unsigned numSlots, unsigned firstSlot, unsigned attBase, cudaStream_t st);

// vertex -> incident triangles (ascending triangle id), CSR: off[numVertices + 1], tris[numIndices].  *badIndex |= 1 when an
// index is >= numVertices (the lists are then unspecified).  scratch: numVertices words.
void vertex_triangles(const unsigned* indices, size_t numIndices, unsigned numVertices, unsigned* off, unsigned* tris,
                      unsigned* scratch, int* badIndex, cudaStream_t st);

}  // namespace setup
}  // namespace velvet
