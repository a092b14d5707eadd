// seam.hpp -- C++ launchers of the kernel seam (explicit params + stream), shared by the C ABI wrappers in
// seam_kernels.cu and by the solver's reference-order pipeline.  One function per reference free function
// (VtClothSolverGPU.cuh L99-168, SpatialHashGPU.cuh L17-25); all pointers are packed-float3 device-accessible.
#pragma once

#include <cuda_runtime.h>

#include "../../include/velvet_b200.h"

namespace velvet {
namespace seam {

void InitializePositions(float* positions, int start, int count, const float* modelMatrix16, cudaStream_t st);
void PredictPositions(const VtSimParams& P, float* predicted, float* velocities, const float* positions, float dt,
                      cudaStream_t st);
void SolveStretch(float* predicted, float* deltas, int* deltaCounts, const int* stretchIndices,
                  const float* stretchLengths, const float* invMasses, unsigned n, cudaStream_t st);
void SolveBending(const VtSimParams& P, float* predicted, float* deltas, int* deltaCounts, const unsigned* bendIndices,
                  const float* bendAngles, const float* invMass, unsigned n, float dt, cudaStream_t st);
void SolveAttachment(const VtSimParams& P, float* predicted, float* deltas, int* deltaCounts, const float* invMass,
                     const int* attachParticleIDs, const int* attachSlotIDs, const float* attachSlotPositions,
                     const float* attachDistances, int n, cudaStream_t st);
void ApplyDeltas(const VtSimParams& P, float* predicted, float* deltas, int* deltaCounts, cudaStream_t st);
void CollideSDF(const VtSimParams& P, float* predicted, const VtSDFCollider* colliders, const float* positions,
                unsigned numColliders, float dt, cudaStream_t st);
void CollideParticles(const VtSimParams& P, float* deltas, int* deltaCounts, float* predicted, const float* invMasses,
                      const unsigned* neighbors, const float* positions, cudaStream_t st);
void Finalize(const VtSimParams& P, float* velocities, float* positions, const float* predicted, float dt,
              cudaStream_t st);
void ComputeNormal(const VtSimParams& P, float* normals, const float* positions, const unsigned* indices,
                   unsigned numTriangles, cudaStream_t st);
void HashObjects(unsigned* particleHash, unsigned* particleIndex, unsigned* cellStart, unsigned* cellEnd,
                 unsigned* neighbors, const float* positions, const float* originalPositions, VtHashParams hp,
                 cudaStream_t st);
// kernels launched by the most recent call on this thread (for launch accounting)
int LastLaunchCount();

}  // namespace seam
}  // namespace velvet
