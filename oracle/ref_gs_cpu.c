/*
 * oracle/ref_gs_cpu.c -- O2, see ref_gs_cpu.h.  TEST / BASELINE INFRASTRUCTURE ONLY.  Parity unpinned.
 * Single thread, fp32, in-place Gauss-Seidel exactly in the reference's loop order.
 */
#include "ref_gs_cpu.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float x, y, z; } v3;
static inline v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 add(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 muls(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static inline v3 divs(v3 a, float s) { return V(a.x / s, a.y / s, a.z / s); }
static inline v3 neg(v3 a) { return V(-a.x, -a.y, -a.z); }
static inline float dot3(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline float len3(v3 a) { return sqrtf(dot3(a, a)); }
static inline v3 normalize3(v3 a) { return muls(a, 1.0f / sqrtf(dot3(a, a))); }
static inline v3 cross3(v3 a, v3 b) { return V(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
static inline float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

#define K_EPSILON 1e-6f /* VtClothSolverCPU.hpp L388 */

typedef struct { int a, b; float d; } Stretch;
typedef struct { int i; v3 p; } Attach;
typedef struct { int i1, i2, i3, i4; float angle; } Bend;

struct O2Solver {
    O1SimParams P;
    int resolution, n, numIndices;
    v3 *positions, *predicted, *velocities, *normals;
    float* invMass;
    uint32_t* indices;
    Stretch* stretch; int numStretch;
    Attach* attach; int numAttach;
    Bend* bend; int numBend;
    float particleDiameter;
    /* colliders */
    int numColliders; int* colType; v3* colPos; float* colScaleX;
    /* SpatialHashCPU */
    float spacing; int tableSize;
    int *cellStart, *cellEntries;
    int *nbOff, *nbList; size_t nbCap;
};

static inline int int_coord(const O2Solver* s, float v) { return (int)floorf(v / s->spacing); }
static inline int hash_coords(const O2Solver* s, int x, int y, int z)
{ /* SpatialHashCPU.hpp L72-76 */
    int32_t h = (int32_t)((uint32_t)x * 92837111u) ^ (int32_t)((uint32_t)y * 689287499u) ^ (int32_t)((uint32_t)z * 283923481u);
    int32_t r = h % s->tableSize;
    return r < 0 ? -r : r;
}
static inline int hash_position(const O2Solver* s, v3 p)
{
    return hash_coords(s, int_coord(s, p.x), int_coord(s, p.y), int_coord(s, p.z));
}

/* SpatialHashCPU::HashObjects, L24-54; CacheNeighbors/QueryNeighbors, L88-122 */
static void hash_objects(O2Solver* s, const v3* positions)
{
    const int n = s->n, T = s->tableSize;
    memset(s->cellStart, 0, sizeof(int) * (size_t)(T + 1));
    memset(s->cellEntries, 0, sizeof(int) * (size_t)n);
    for (int i = 0; i < n; i++) s->cellStart[hash_position(s, positions[i])]++;
    int start = 0;
    for (int i = 0; i < T; i++) { start += s->cellStart[i]; s->cellStart[i] = start; }
    s->cellStart[T] = start;
    for (int i = 0; i < n; i++) {
        int c = hash_position(s, positions[i]);
        s->cellStart[c]--;
        s->cellEntries[s->cellStart[c]] = i;
    }
    size_t total = 0;
    for (int i = 0; i < n; i++) {
        s->nbOff[i] = (int)total;
        v3 p = positions[i];
        int ix = int_coord(s, p.x), iy = int_coord(s, p.y), iz = int_coord(s, p.z);
        for (int x = ix - 1; x <= ix + 1; x++)
            for (int y = iy - 1; y <= iy + 1; y++)
                for (int z = iz - 1; z <= iz + 1; z++) {
                    int h = hash_coords(s, x, y, z);
                    int b = s->cellStart[h], e = s->cellStart[h + 1];
                    if (total + (size_t)(e - b) > s->nbCap) {
                        s->nbCap = (total + (size_t)(e - b)) * 3 / 2 + 1024;
                        s->nbList = (int*)realloc(s->nbList, sizeof(int) * s->nbCap);
                    }
                    for (int k = b; k < e; k++) s->nbList[total++] = s->cellEntries[k];
                }
    }
    s->nbOff[n] = (int)total;
}

/* VtClothSolverCPU::ComputeFriction, L331-346 */
static v3 compute_friction(const O2Solver* s, v3 correction, v3 relVel)
{
    v3 friction = V(0, 0, 0);
    float correctionLength = len3(correction);
    if (s->P.friction > 0 && correctionLength > 0) {
        v3 norm = divs(correction, correctionLength);
        v3 tanVel = sub(relVel, muls(norm, dot3(relVel, norm)));
        float tanLength = len3(tanVel);
        float maxTan = correctionLength * s->P.friction;
        friction = muls(neg(tanVel), fminf(maxTan / tanLength, 1.0f));
    }
    return friction;
}

/* Collider::ComputeSDF / ComputePlaneSDF / ComputeSphereSDF, Collider.hpp L43-77 */
static v3 collider_sdf(const O2Solver* s, int c, v3 p)
{
    const float margin = s->P.collisionMargin;
    if (s->colType[c] == 1) {
        if (p.y < margin) return V(0, margin - p.y, 0);
        return V(0, 0, 0);
    }
    float radius = s->colScaleX[c] + margin;
    v3 diff = sub(p, s->colPos[c]);
    float distance = len3(diff);
    if (distance < radius) return muls(divs(diff, distance), radius - distance);
    return V(0, 0, 0);
}

/* L240-256 */
static void collide_sdf(O2Solver* s, v3* positions)
{
    for (int i = 0; i < s->n; i++)
        for (int c = 0; c < s->numColliders; c++) {
            v3 pos = positions[i];
            v3 correction = collider_sdf(s, c, pos);
            positions[i] = add(positions[i], correction);
            v3 relVel = sub(positions[i], s->positions[i]);
            positions[i] = add(positions[i], compute_friction(s, correction, relVel));
        }
}

/* L185-192 */
static void predict_positions(O2Solver* s, float dt)
{
    v3 g = V(s->P.gravity[0], s->P.gravity[1], s->P.gravity[2]);
    for (int i = 0; i < s->n; i++) {
        s->velocities[i] = add(s->velocities[i], muls(g, dt));
        s->predicted[i] = add(s->positions[i], muls(s->velocities[i], dt));
    }
}

/* L194-219: unilateral (distance > expected) */
static void solve_stretch(O2Solver* s)
{
    for (int k = 0; k < s->numStretch; k++) {
        int a = s->stretch[k].a, b = s->stretch[k].b;
        float expected = s->stretch[k].d;
        v3 diff = sub(s->predicted[a], s->predicted[b]);
        float distance = len3(diff);
        float w1 = s->invMass[a], w2 = s->invMass[b];
        if (distance > expected && w1 + w2 > 0) {
            v3 gradient = divs(diff, distance + K_EPSILON);
            float denom = w1 + w2;
            float lambda = (distance - expected) / denom;
            s->predicted[a] = sub(s->predicted[a], muls(gradient, w1 * lambda));
            s->predicted[b] = add(s->predicted[b], muls(gradient, w2 * lambda));
        }
    }
}

/* L221-272 */
static void solve_bending(O2Solver* s, float dt)
{
    float xpbd_bend = s->P.bendCompliance / dt / dt;
    for (int k = 0; k < s->numBend; k++) {
        int idx1 = s->bend[k].i3, idx2 = s->bend[k].i2, idx3 = s->bend[k].i1, idx4 = s->bend[k].i4; /* get<2>,<1>,<0>,<3> */
        float expectedAngle = s->bend[k].angle;
        float w1 = s->invMass[idx1], w2 = s->invMass[idx2], w3 = s->invMass[idx3], w4 = s->invMass[idx4];
        v3 p1 = s->predicted[idx1];
        v3 p2 = sub(s->predicted[idx2], p1), p3 = sub(s->predicted[idx3], p1), p4 = sub(s->predicted[idx4], p1);
        v3 n1 = normalize3(cross3(p2, p3));
        v3 n2 = normalize3(cross3(p2, p4));
        float d = clampf(dot3(n1, n2), 0.0f, 1.0f);
        float angle = acosf(d);
        if (angle < K_EPSILON || isnan(d)) continue;
        v3 q3 = divs(add(cross3(p2, n2), muls(cross3(n1, p2), d)), len3(cross3(p2, p3)) + K_EPSILON);
        v3 q4 = divs(add(cross3(p2, n1), muls(cross3(n2, p2), d)), len3(cross3(p2, p4)) + K_EPSILON);
        v3 q2 = sub(neg(divs(add(cross3(p3, n2), muls(cross3(n1, p3), d)), len3(cross3(p2, p3)) + K_EPSILON)),
                    divs(add(cross3(p4, n1), muls(cross3(n2, p4), d)), len3(cross3(p2, p4)) + K_EPSILON));
        v3 q1 = sub(sub(neg(q2), q3), q4);
        float denom = xpbd_bend + (w1 * dot3(q1, q1) + w2 * dot3(q2, q2) + w3 * dot3(q3, q3) + w4 * dot3(q4, q4));
        if (denom < K_EPSILON) continue;
        float lambda = sqrtf(1.0f - d * d) * (angle - expectedAngle) / denom;
        s->predicted[idx1] = add(s->predicted[idx1], muls(q1, w1 * lambda));
        s->predicted[idx2] = add(s->predicted[idx2], muls(q2, w2 * lambda));
        s->predicted[idx3] = add(s->predicted[idx3], muls(q3, w3 * lambda));
        s->predicted[idx4] = add(s->predicted[idx4], muls(q4, w4 * lambda));
    }
}

/* L282-315 */
static void collide_particles(O2Solver* s)
{
    for (int i = 0; i < s->n; i++)
        for (int k = s->nbOff[i]; k < s->nbOff[i + 1]; k++) {
            int j = s->nbList[k];
            if (i >= j) continue;
            float expected = s->particleDiameter;
            v3 diff = sub(s->predicted[i], s->predicted[j]);
            float distance = len3(diff);
            float w1 = s->invMass[i], w2 = s->invMass[j];
            if (distance < expected && w1 + w2 > 0) {
                v3 gradient = divs(diff, distance + K_EPSILON);
                float denom = w1 + w2;
                float lambda = (distance - expected) / denom;
                v3 common = muls(gradient, lambda);
                s->predicted[i] = sub(s->predicted[i], muls(common, w1));
                s->predicted[j] = add(s->predicted[j], muls(common, w2));
                v3 relVel = sub(sub(s->predicted[i], s->positions[i]), sub(s->predicted[j], s->positions[j]));
                v3 friction = compute_friction(s, common, relVel);
                s->predicted[i] = add(s->predicted[i], muls(friction, w1));
                s->predicted[j] = sub(s->predicted[j], muls(friction, w2));
            }
        }
}

/* L69-100 */
void o2_simulate(O2Solver* s)
{
    const float frameTime = 1.0f / 60.0f;
    const float substepTime = frameTime / (float)s->P.numSubsteps;
    collide_sdf(s, s->positions);
    predict_positions(s, frameTime);
    hash_objects(s, s->predicted);
    for (int substep = 0; substep < s->P.numSubsteps; substep++) {
        predict_positions(s, substepTime);
        for (int it = 0; it < s->P.numIterations; it++) {
            solve_stretch(s);
            solve_bending(s, substepTime);
            collide_particles(s);
            collide_sdf(s, s->predicted);
            for (int k = 0; k < s->numAttach; k++) s->predicted[s->attach[k].i] = s->attach[k].p; /* L258-266 */
        }
        for (int i = 0; i < s->n; i++) { /* Finalize L317-327 */
            s->velocities[i] = muls(divs(sub(s->predicted[i], s->positions[i]), substepTime), 1 - s->P.damping * substepTime);
            s->positions[i] = s->predicted[i];
        }
    }
    /* ComputeNormals L348-371 */
    memset(s->normals, 0, sizeof(v3) * (size_t)s->n);
    for (int i = 0; i + 2 < s->numIndices; i += 3) {
        uint32_t a = s->indices[i], b = s->indices[i + 1], c = s->indices[i + 2];
        v3 nrm = cross3(sub(s->positions[b], s->positions[a]), sub(s->positions[c], s->positions[a]));
        s->normals[a] = add(s->normals[a], nrm); s->normals[b] = add(s->normals[b], nrm); s->normals[c] = add(s->normals[c], nrm);
    }
    for (int i = 0; i < s->n; i++) s->normals[i] = normalize3(s->normals[i]);
}

O2Solver* o2_create(const O1SimParams* params, int resolution, const float* vertices, const uint32_t* indices,
                    const float* M, const int* attachedIndices, int numAttached)
{
    O2Solver* s = (O2Solver*)calloc(1, sizeof(O2Solver));
    s->P = *params;
    s->resolution = resolution;
    const int S = resolution + 1;
    s->n = S * S;
    s->numIndices = 6 * resolution * resolution;
    s->positions = (v3*)calloc((size_t)s->n, sizeof(v3));
    s->predicted = (v3*)calloc((size_t)s->n, sizeof(v3));
    s->velocities = (v3*)calloc((size_t)s->n, sizeof(v3));
    s->normals = (v3*)calloc((size_t)s->n, sizeof(v3));
    s->invMass = (float*)malloc(sizeof(float) * (size_t)s->n);
    s->indices = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)s->numIndices);
    memcpy(s->indices, indices, sizeof(uint32_t) * (size_t)s->numIndices);
    for (int i = 0; i < s->n; i++) { /* L47-52 */
        float x = vertices[3 * i], y = vertices[3 * i + 1], z = vertices[3 * i + 2];
        s->positions[i] = V((M[0] * x + M[4] * y) + (M[8] * z + M[12]), (M[1] * x + M[5] * y) + (M[9] * z + M[13]),
                            (M[2] * x + M[6] * y) + (M[10] * z + M[14]));
        s->invMass[i] = 1.0f;
    }
    s->particleDiameter = len3(sub(s->positions[0], s->positions[S])); /* L61 */
    s->spacing = s->particleDiameter;                                   /* L62, SpatialHashCPU L16-17 */
    s->tableSize = 2 * s->n;
    s->cellStart = (int*)calloc((size_t)s->tableSize + 1, sizeof(int));
    s->cellEntries = (int*)calloc((size_t)s->n, sizeof(int));
    s->nbOff = (int*)calloc((size_t)s->n + 1, sizeof(int));
    s->nbCap = (size_t)s->n * 32; s->nbList = (int*)malloc(sizeof(int) * s->nbCap);

    s->stretch = (Stretch*)malloc(sizeof(Stretch) * (size_t)(4 * resolution * resolution + 2 * resolution + 4));
#define VAT(x, y) ((x) * S + (y))
#define PUSH(a_, b_) do { Stretch c; c.a = (a_); c.b = (b_); c.d = len3(sub(s->positions[c.a], s->positions[c.b])); s->stretch[s->numStretch++] = c; } while (0)
    for (int x = 0; x < S; x++) /* GenerateStretch L109-151 */
        for (int y = 0; y < S; y++) {
            if (y != resolution) PUSH(VAT(x, y), VAT(x, y + 1));
            if (x != resolution) PUSH(VAT(x, y), VAT(x + 1, y));
            if (y != resolution && x != resolution) { PUSH(VAT(x, y), VAT(x + 1, y + 1)); PUSH(VAT(x, y + 1), VAT(x + 1, y)); }
        }
#undef PUSH
#undef VAT
    s->attach = (Attach*)malloc(sizeof(Attach) * (size_t)(numAttached + 1)); /* GenerateAttachment L153-160 */
    for (int k = 0; k < numAttached; k++) {
        s->attach[k].i = attachedIndices[k];
        s->attach[k].p = s->positions[attachedIndices[k]];
        s->invMass[attachedIndices[k]] = 0;
    }
    s->numAttach = numAttached;
    s->bend = (Bend*)malloc(sizeof(Bend) * (size_t)(resolution * resolution + 1)); /* GenerateBending L162-176 */
    for (int i = 0; i + 5 < s->numIndices; i += 6) {
        Bend b; b.i1 = (int)indices[i]; b.i2 = (int)indices[i + 1]; b.i3 = (int)indices[i + 2]; b.i4 = (int)indices[i + 5]; b.angle = 0;
        s->bend[s->numBend++] = b;
    }
    return s;
}

void o2_set_colliders(O2Solver* s, const int* types, const float* positions3, const float* scalesX, int n)
{
    s->colType = (int*)realloc(s->colType, sizeof(int) * (size_t)(n + 1));
    s->colPos = (v3*)realloc(s->colPos, sizeof(v3) * (size_t)(n + 1));
    s->colScaleX = (float*)realloc(s->colScaleX, sizeof(float) * (size_t)(n + 1));
    for (int i = 0; i < n; i++) {
        s->colType[i] = types[i];
        s->colPos[i] = V(positions3[3 * i], positions3[3 * i + 1], positions3[3 * i + 2]);
        s->colScaleX[i] = scalesX[i];
    }
    s->numColliders = n;
}

void o2_destroy(O2Solver* s)
{
    if (!s) return;
    free(s->positions); free(s->predicted); free(s->velocities); free(s->normals); free(s->invMass); free(s->indices);
    free(s->stretch); free(s->attach); free(s->bend); free(s->colType); free(s->colPos); free(s->colScaleX);
    free(s->cellStart); free(s->cellEntries); free(s->nbOff); free(s->nbList);
    free(s);
}

float* o2_positions(O2Solver* s) { return (float*)s->positions; }
float* o2_normals(O2Solver* s) { return (float*)s->normals; }
int o2_num_particles(O2Solver* s) { return s->n; }
