/*
 * oracle/ref_gs_cpu.h -- O2: headless CPU restatement of the reference's own CPU solver
 * (VtClothSolverCPU.hpp L42-405 + SpatialHashCPU.hpp L14-123 + Collider.hpp L43-77), minus GL.
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY: bench.py times it as the reference's CPU implementation of the path
 * (cpu_baseline and --impl reference); nothing in velvet_b200/ links or calls it.
 *
 * It is a DIFFERENT ALGORITHM from the GPU path (single-thread Gauss-Seidel in place, unilateral stretch,
 * acos-gradient bending clamped to [0,1], hard-set attachments, no long-range attachments, collisions inside
 * the iteration loop, hash once per frame on a frame-dt prediction, no speed clamp) -- SURVEY.md section 8c --
 * so it is a timing baseline and a behavioural envelope, not a parity oracle.  The reference class is
 * bit-rotted upstream (Scene.hpp L278 names a non-existent class); parity unpinned.
 */
#ifndef VELVET_ORACLE_REF_GS_CPU_H
#define VELVET_ORACLE_REF_GS_CPU_H

#include <stdint.h>

#include "ref_jacobi_cpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct O2Solver O2Solver;

/* VtClothSolverCPU(resolution) + SetAttachedIndices + Initialize(mesh, modelMatrix), hpp L32-67 */
O2Solver* o2_create(const O1SimParams* params, int resolution, const float* vertices, const uint32_t* indices,
                    const float* model16, const int* attachedIndices, int numAttached);
void o2_destroy(O2Solver* s);
/* Collider components found in the scene (Collider.hpp L12-41): type (0 sphere, 1 plane), position, scale.x */
void o2_set_colliders(O2Solver* s, const int* types, const float* positions3, const float* scalesX, int n);
void o2_simulate(O2Solver* s); /* hpp L69-100 */
float* o2_positions(O2Solver* s);
float* o2_normals(O2Solver* s);
int o2_num_particles(O2Solver* s);

#ifdef __cplusplus
}
#endif
#endif
