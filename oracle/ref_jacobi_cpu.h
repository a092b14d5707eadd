/*
 * oracle/ref_jacobi_cpu.h -- O1: CPU restatement of vitalight/Velvet's GPU XPBD path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is product code: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product (velvet_b200/) never links or calls it.
 *
 * PARITY UNPINNED by the reference: vitalight/Velvet ships no tests, golden vectors
 * or fixtures for this path.  This restatement is pinned instead by
 *   (1) property tests (brute-force neighbour truth, pinned-particle behaviour, ...),
 *   (2) the reference's own CUDA kernels compiled on stand-in headers (oracle/ref_cuda,
 *       built into oracle/_ref/) and run on the GPU box on the same inputs.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference/Velvet).  Arithmetic is fp32, sequentialised in ascending thread id,
 * deltas accumulated in constraint-id order stretch -> attach -> bend; compile with
 * -ffp-contract=off.  glm (unpinned vcpkg dependency, >= 0.9.9) semantics are restated:
 * dot = (x*x' + y*y') + z*z'; length = sqrt(dot); normalize = v * (1/sqrt(dot));
 * vec/scalar = per-component division; mat4*vec4 = (m0*x + m1*y) + (m2*z + m3*w).
 */
#ifndef VELVET_ORACLE_REF_JACOBI_CPU_H
#define VELVET_ORACLE_REF_JACOBI_CPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Common.hpp L19-47: 80-byte POD, offsets asserted in the .c file. */
typedef struct O1SimParams {
    int32_t numSubsteps;
    int32_t numIterations;
    int32_t maxNumNeighbors;
    float maxSpeed;
    float gravity[3];
    float bendCompliance;
    float damping;
    float relaxationFactor;
    float longRangeStretchiness;
    float collisionMargin;
    float friction;
    uint8_t enableSelfCollision;
    uint8_t _pad[3];
    int32_t interleavedHash;
    uint32_t numParticles;
    float particleDiameter;
    float deltaTime;
    float particleDiameterScalar;
    float hashCellSizeScalar;
} O1SimParams;

/* VtClothSolverGPU.cuh L8-18: 196-byte POD (ColliderType: Sphere 0, Plane 1, Cube 2). */
typedef struct O1SDFCollider {
    int32_t type;
    float position[3];
    float scale[3];
    float deltaTime;
    float curTransform[9];     /* mat3, column-major */
    float invCurTransform[16]; /* mat4, column-major */
    float lastTransform[16];
} O1SDFCollider;

/* SpatialHashGPU.cuh L7-15: 24-byte POD. */
typedef struct O1HashParams {
    uint32_t numObjects;
    uint32_t maxNumNeighbors;
    float cellSpacing;
    float cellSpacing2;
    int32_t tableSize;
    float particleDiameter2;
} O1HashParams;

void o1_default_params(O1SimParams* p);

/* ---- kernel-level restatements (AoS float3 buffers, host memory) ---- */
void o1_initialize_positions(float* positions, int start, int count, const float* model16);
void o1_predict_positions(const O1SimParams* P, float* predicted, float* velocities,
                          const float* positions, float dt);
void o1_solve_stretch(float* predicted, float* deltas, int* deltaCounts, const int* stretchIndices,
                      const float* stretchLengths, const float* invMasses, uint32_t n);
void o1_solve_bending(const O1SimParams* P, float* predicted, float* deltas, int* deltaCounts,
                      const uint32_t* bendIndices, const float* bendAngles, const float* invMass,
                      uint32_t n, float dt);
void o1_solve_attachment(const O1SimParams* P, float* predicted, float* deltas, int* deltaCounts,
                         const float* invMass, const int* attachParticleIDs, const int* attachSlotIDs,
                         const float* attachSlotPositions, const float* attachDistances, int n);
void o1_apply_deltas(const O1SimParams* P, float* predicted, float* deltas, int* deltaCounts);
void o1_collide_sdf(const O1SimParams* P, float* predicted, const O1SDFCollider* colliders,
                    const float* positions, uint32_t numColliders, float dt);
void o1_collide_particles(const O1SimParams* P, float* deltas, int* deltaCounts, float* predicted,
                          const float* invMasses, const uint32_t* neighbors, const float* positions);
void o1_finalize(const O1SimParams* P, float* velocities, float* positions, const float* predicted,
                 float dt);
void o1_compute_normal(const O1SimParams* P, float* normals, const float* positions,
                       const uint32_t* indices, uint32_t numTriangles);
void o1_hash_objects(uint32_t* particleHash, uint32_t* particleIndex, uint32_t* cellStart,
                     uint32_t* cellEnd, uint32_t* neighbors, const float* positions,
                     const float* originalPositions, O1HashParams hp);
int o1_hash_position(const float* p3, float cellSpacing, int tableSize);
float o1_acosf(float x);

/* ---- host-side helpers (glm restatements) ---- */
void o1_transform_matrix(const float* position3, const float* rotationDeg3, const float* scale3,
                         float* out16);
void o1_mat4_inverse(const float* m16, float* out16);
void o1_make_collider(int type, const float* position3, const float* scale3, const float* cur16,
                      const float* last16, float deltaTime, O1SDFCollider* out);
void o1_generate_cloth_mesh(int resolution, float* vertices /*3*(R+1)^2*/,
                            uint32_t* indices /*6*R^2*/);

/* ---- solver object: VtClothSolverGPU.hpp restated ---- */
typedef struct O1Solver O1Solver;
O1Solver* o1_solver_create(const O1SimParams* params);
void o1_solver_destroy(O1Solver* s);
O1SimParams* o1_solver_params(O1Solver* s);
int o1_solver_add_cloth(O1Solver* s, const float* vertices, int numVertices, const uint32_t* indices,
                        int numIndices, const float* model16, float particleDiameter);
void o1_solver_add_stretch(O1Solver* s, int idx1, int idx2, float distance);
void o1_solver_add_attach_slot(O1Solver* s, const float* pos3);
void o1_solver_add_attach(O1Solver* s, int particleIndex, int slotIndex, float distance);
void o1_solver_add_bend(O1Solver* s, uint32_t i1, uint32_t i2, uint32_t i3, uint32_t i4, float angle);
void o1_solver_set_colliders(O1Solver* s, const O1SDFCollider* colliders, int n);
void o1_solver_simulate(O1Solver* s);
void o1_solver_hash(O1Solver* s); /* SpatialHashGPU::Hash(predicted) */

/* VtClothObjectGPU::Start restated: registers one grid cloth incl. constraints. Returns offset. */
int o1_cloth_object_start(O1Solver* s, int resolution, const float* vertices, const uint32_t* indices,
                          const float* model16, const int* attachedIndices, int numAttached);

/* buffer access; `which` names follow VtClothSolverGPU.hpp L209-231 / SpatialHashGPU.hpp L54-60 */
enum {
    O1_BUF_POSITIONS = 0, O1_BUF_NORMALS, O1_BUF_INDICES, O1_BUF_VELOCITIES, O1_BUF_PREDICTED,
    O1_BUF_DELTAS, O1_BUF_DELTACOUNTS, O1_BUF_INVMASSES, O1_BUF_STRETCHINDICES,
    O1_BUF_STRETCHLENGTHS, O1_BUF_BENDINDICES, O1_BUF_BENDANGLES, O1_BUF_ATTACHPARTICLEIDS,
    O1_BUF_ATTACHSLOTIDS, O1_BUF_ATTACHDISTANCES, O1_BUF_ATTACHSLOTPOSITIONS,
    O1_BUF_NEIGHBORS, O1_BUF_INITIALPOSITIONS, O1_BUF_PARTICLEHASH, O1_BUF_PARTICLEINDEX,
    O1_BUF_CELLSTART, O1_BUF_CELLEND
};
void* o1_solver_buffer(O1Solver* s, int which, uint64_t* count /* in elements of 4 bytes */);

#ifdef __cplusplus
}
#endif
#endif
