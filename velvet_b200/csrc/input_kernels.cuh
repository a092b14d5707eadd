// input_kernels.cuh -- interactive inputs of the solver as device operations.
//
// Reference: MouseGrabber (MouseGrabber.hpp L31-110), which the cloth scenes run every frame next to the solver
// (VtClothSolverGPU.hpp L60-66): a HOST loop over every particle position picks the vertex closest to the mouse ray
// (FindClosestVertexToRay, L92-110), then host code pins it (invMass = 0, L50-52), drags it (positions / velocities write,
// L66-79) and restores the mass on release (L57-62) -- all through managed memory, i.e. the whole position array migrates to
// the host and back on every pick.  Here the pick is one reduction kernel over the device-resident positions and the
// pin / drag / release are one-thread kernels on the solver stream; only the result of a pick (index, distance) returns.
//
// Arithmetic (compiled without FMA contraction, like the host code it replaces): glm::dot = (x*x' + y*y') + z*z',
// glm::cross = (a.y*b.z - b.y*a.z, a.z*b.x - b.z*a.x, a.x*b.y - b.x*a.y), glm::length = sqrt(dot(v, v)).
#pragma once

#include <cuda_runtime.h>

#include "vt_math.cuh"

namespace velvet {
namespace input {

struct GrabState {            // device-resident
    unsigned long long best;  // packed (ordered distanceToView, index) of the pick in progress
    int index;                // grabbed particle, -1 when none
    float distanceToOrigin;   // RaycastCollision::distanceToOrigin of the pick (L26)
    float savedInvMass;       // m_grabbedVertexMass (L83)
    int grabbing;
};

// FindClosestVertexToRay + the pinning of HandleMouseInteraction (L40-55): afterwards state->index / distanceToOrigin hold the
// pick (index -1: the ray met no particle within one particle diameter) and the picked particle's inverse mass is 0.
void grab(GrabState* state, const float* positions, float* invMasses, unsigned numParticles, vec3 rayOrigin, vec3 rayDirection,
          float particleDiameter, cudaStream_t st);
// UpdateGrappedVertex (L66-79): target = Lerp(origin + direction * distanceToOrigin, current, 0.8); position = target,
// velocity = (target - current) / fixedDeltaTime.  No-op unless a particle is grabbed.
void drag(const GrabState* state, float* positions, float* velocities, vec3 rayOrigin, vec3 rayDirection, float fixedDeltaTime,
          cudaStream_t st);
// mouse-up branch of HandleMouseInteraction (L57-62): the grabbed particle gets its inverse mass back
void release(GrabState* state, float* invMasses, cudaStream_t st);

}  // namespace input
}  // namespace velvet
