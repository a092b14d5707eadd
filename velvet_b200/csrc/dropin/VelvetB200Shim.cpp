// VelvetB200Shim.cpp -- kernel-level drop-in (way A): compiled INSTEAD of the reference's
// VtClothSolverGPU.cu and SpatialHashGPU.cu.  It defines the twelve free functions those two files define
// (VtClothSolverGPU.cuh L99-168, SpatialHashGPU.cuh L17-25), with the reference's own declarations in scope, and forwards
// each to the extern "C" entry point of libvelvet_b200.so.  VtClothSolverGPU.hpp / SpatialHashGPU.hpp stay untouched.
//
// tests/test_dropin_gpu.py builds this file against the reference's headers (oracle/ref_cuda/build_ref_cuda.sh) and runs
// the reference-side orchestration through it.
#include "VtClothSolverGPU.cuh"  // reference declarations (namespace Velvet, glm types)
#include "SpatialHashGPU.cuh"

#define VELVET_B200_USE_REFERENCE_SIMPARAMS  // ::VtSimParams is the reference's (Common.hpp L19-47), same 80 bytes
#include <velvet_b200.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

static_assert(sizeof(::VtSimParams) == 80 && sizeof(Velvet::SDFCollider) == 196 && sizeof(Velvet::HashParams) == 24, "POD layout");
static_assert(sizeof(::VtSDFCollider) == sizeof(Velvet::SDFCollider) && sizeof(::VtHashParams) == sizeof(Velvet::HashParams), "POD layout");

namespace {
void ck(int st)
{
    if (st < 0) {
        std::fprintf(stderr, "velvet_b200: %s\n", velvet_last_error());
        std::exit(EXIT_FAILURE);
    }
}
float* f(glm::vec3* p) { return reinterpret_cast<float*>(p); }
const float* f(const glm::vec3* p) { return reinterpret_cast<const float*>(p); }
}  // namespace

namespace Velvet {

void SetSimulationParams(VtSimParams* hostParams) { ck(velvet_SetSimulationParams(hostParams)); }

void InitializePositions(glm::vec3* positions, const int start, const int count, const glm::mat4 modelMatrix)
{
    ck(velvet_InitializePositions(f(positions), start, count, &modelMatrix[0][0]));
}

void PredictPositions(glm::vec3* predicted, glm::vec3* velocities, CONST(glm::vec3*) positions, const float deltaTime)
{
    ck(velvet_PredictPositions(f(predicted), f(velocities), f(positions), deltaTime));
}

void SolveStretch(glm::vec3* predicted, glm::vec3* deltas, int* deltaCounts, CONST(int*) stretchIndices, CONST(float*) stretchLengths,
                  CONST(float*) invMasses, const uint numConstraints)
{
    ck(velvet_SolveStretch(f(predicted), f(deltas), deltaCounts, stretchIndices, stretchLengths, invMasses, numConstraints));
}

void SolveBending(glm::vec3* predicted, glm::vec3* deltas, int* deltaCounts, CONST(uint*) bendingIndices, CONST(float*) bendingAngles,
                  CONST(float*) invMass, const uint numConstraints, const float deltaTime)
{
    ck(velvet_SolveBending(f(predicted), f(deltas), deltaCounts, bendingIndices, bendingAngles, invMass, numConstraints, deltaTime));
}

void SolveAttachment(glm::vec3* predicted, glm::vec3* deltas, int* deltaCounts, CONST(float*) invMass, CONST(int*) attachParticleIDs,
                     CONST(int*) attachSlotIDs, CONST(glm::vec3*) attachSlotPositions, CONST(float*) attachDistances,
                     const int numConstraints)
{
    ck(velvet_SolveAttachment(f(predicted), f(deltas), deltaCounts, invMass, attachParticleIDs, attachSlotIDs, f(attachSlotPositions),
                              attachDistances, numConstraints));
}

void ApplyDeltas(glm::vec3* predicted, glm::vec3* deltas, int* deltaCounts) { ck(velvet_ApplyDeltas(f(predicted), f(deltas), deltaCounts)); }

void CollideSDF(glm::vec3* predicted, CONST(SDFCollider*) colliders, CONST(glm::vec3*) positions, const uint numColliders,
                const float deltaTime)
{
    ck(velvet_CollideSDF(f(predicted), reinterpret_cast<const ::VtSDFCollider*>(colliders), f(positions), numColliders, deltaTime));
}

void CollideParticles(glm::vec3* deltas, int* deltaCounts, glm::vec3* predicted, CONST(float*) invMasses, CONST(uint*) neighbors,
                      CONST(glm::vec3*) positions)
{
    ck(velvet_CollideParticles(f(deltas), deltaCounts, f(predicted), invMasses, neighbors, f(positions)));
}

void Finalize(glm::vec3* velocities, glm::vec3* positions, CONST(glm::vec3*) predicted, const float deltaTime)
{
    ck(velvet_Finalize(f(velocities), f(positions), f(predicted), deltaTime));
}

void ComputeNormal(glm::vec3* normals, CONST(glm::vec3*) positions, CONST(uint*) indices, const uint numTriangles)
{
    ck(velvet_ComputeNormal(f(normals), f(positions), indices, numTriangles));
}

void HashObjects(uint* particleHash, uint* particleIndex, uint* cellStart, uint* cellEnd, uint* neighbors, CONST(glm::vec3*) positions,
                 CONST(glm::vec3*) originalPositions, const HashParams params)
{
    ::VtHashParams q;
    std::memcpy(&q, &params, sizeof(q));
    ck(velvet_HashObjects(particleHash, particleIndex, cellStart, cellEnd, neighbors, f(positions), f(originalPositions), q));
}

}  // namespace Velvet
