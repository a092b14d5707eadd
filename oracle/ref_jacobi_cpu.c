/*
 * oracle/ref_jacobi_cpu.c -- O1 parity oracle.  TEST INFRASTRUCTURE ONLY (see the header).
 * The reference ships no golden vectors; pinned against committed outputs of the reference's own CUDA
 * kernels (tests/golden/refcuda_*.npz), by those kernels run live (oracle/ref_cuda) and by property tests.
 *
 * Sequential fp32 restatement of the kernels of VtClothSolverGPU.cu / SpatialHashGPU.cu and
 * of the host orchestration in VtClothSolverGPU.hpp / SpatialHashGPU.hpp /
 * VtClothObjectGPU.hpp.  Build: gcc -O2 -ffp-contract=off -fno-fast-math (twice: portable, and with -mfma so that
 * fmaf() is one instruction; both give the same bits).
 */
#include "ref_jacobi_cpu.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define O1_EPSILON 1e-6f /* Common.cuh L21 */

_Static_assert(sizeof(O1SimParams) == 80, "VtSimParams layout");
_Static_assert(sizeof(O1SDFCollider) == 196, "SDFCollider layout");
_Static_assert(sizeof(O1HashParams) == 24, "HashParams layout");
_Static_assert(__builtin_offsetof(O1SimParams, enableSelfCollision) == 52, "offset");
_Static_assert(__builtin_offsetof(O1SimParams, interleavedHash) == 56, "offset");
_Static_assert(__builtin_offsetof(O1SDFCollider, curTransform) == 32, "offset");
_Static_assert(__builtin_offsetof(O1SDFCollider, invCurTransform) == 68, "offset");
_Static_assert(__builtin_offsetof(O1SDFCollider, lastTransform) == 132, "offset");

/* ------------------------------------------------------------------ glm restated */
typedef struct { float x, y, z; } v3;

static inline v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 ld3(const float* p, size_t i) { return V(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
static inline void st3(float* p, size_t i, v3 v) { p[3 * i] = v.x; p[3 * i + 1] = v.y; p[3 * i + 2] = v.z; }
static inline v3 add(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 mulv(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 muls(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static inline v3 divs(v3 a, float s) { return V(a.x / s, a.y / s, a.z / s); }
static inline v3 neg(v3 a) { return V(-a.x, -a.y, -a.z); }
/* The reference's nvcc build contracts a*b+c into FMAs wherever its compiler chooses, which cannot be reproduced off the
 * device.  The oracle and the product therefore WRITE the contractions: the device-side dot and cross products, the weighted
 * sums of the bending constraint and the acos polynomials are explicit fmaf() calls, identical on both sides
 * (velvet_b200/csrc/vt_math.cuh); everything else is compiled without contraction (-ffp-contract=off / -fmad=false).
 * The spatial hash's candidate tests and the host-side constraint generation (MSVC host code in the reference) use the
 * plain forms. */
static inline float dot3(v3 a, v3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
static inline float dot3_plain(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline float len3(v3 a) { return sqrtf(dot3(a, a)); }
static inline float len3_plain(v3 a) { return sqrtf(dot3_plain(a, a)); }
static inline v3 normalize3(v3 a) { return muls(a, 1.0f / sqrtf(dot3(a, a))); }
static inline v3 cross3(v3 a, v3 b)
{
    return V(fmaf(a.y, b.z, -(b.y * a.z)), fmaf(a.z, b.x, -(b.z * a.x)), fmaf(a.x, b.y, -(b.x * a.y)));
}
/* s*a + t*b per component; weighted sum of four scalars (contracted left to right like the dot product) */
static inline v3 lincomb3(float s, v3 a, float t, v3 b)
{
    return V(fmaf(t, b.x, s * a.x), fmaf(t, b.y, s * a.y), fmaf(t, b.z, s * a.z));
}
static inline float wsum4(float w0, float a0, float w1, float a1, float w2, float a2, float w3, float a3)
{
    return fmaf(w3, a3, fmaf(w2, a2, fmaf(w1, a1, w0 * a0)));
}
static inline float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

/* mat4 (column-major) * (v,w): (m0*x + m1*y) + (m2*z + m3*w); returns xyz */
static inline v3 mat4_mul_point(const float* m, v3 p, float w)
{
    float r[3];
    for (int k = 0; k < 3; k++)
        r[k] = (m[0 + k] * p.x + m[4 + k] * p.y) + (m[8 + k] * p.z + m[12 + k] * w);
    return V(r[0], r[1], r[2]);
}

/* glm mat4*mat4: Result[c] = ((A0*B[c][0] + A1*B[c][1]) + A2*B[c][2]) + A3*B[c][3] */
static void mat4_mul(const float* A, const float* B, float* R)
{
    float out[16];
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++)
            out[4 * c + r] = ((A[0 + r] * B[4 * c + 0] + A[4 + r] * B[4 * c + 1]) + A[8 + r] * B[4 * c + 2])
                             + A[12 + r] * B[4 * c + 3];
    memcpy(R, out, sizeof(out));
}

void o1_default_params(O1SimParams* p)
{ /* Common.hpp L21-46 host initialisers */
    memset(p, 0, sizeof(*p));
    p->numSubsteps = 2;
    p->numIterations = 4;
    p->maxNumNeighbors = 64;
    p->maxSpeed = 50.0f;
    p->gravity[0] = 0.0f; p->gravity[1] = -9.8f; p->gravity[2] = 0.0f;
    p->bendCompliance = 0.0f;
    p->damping = 0.25f;
    p->relaxationFactor = 1.0f;
    p->longRangeStretchiness = 1.2f;
    p->collisionMargin = 0.06f;
    p->friction = 0.1f;
    p->enableSelfCollision = 1;
    p->interleavedHash = 3;
    p->particleDiameterScalar = 1.5f;
    p->hashCellSizeScalar = 1.5f;
}

/* ------------------------------------------------------------------ kernels */

/* VtClothSolverGPU.cu L30-34 */
void o1_initialize_positions(float* positions, int start, int count, const float* model16)
{
    for (int id = 0; id < count; id++)
        st3(positions, (size_t)(start + id), mat4_mul_point(model16, ld3(positions, (size_t)(start + id)), 1.0f));
}

/* VtClothSolverGPU.cu L42-53: applies to every particle, pinned ones included */
void o1_predict_positions(const O1SimParams* P, float* predicted, float* velocities,
                          const float* positions, float dt)
{
    v3 g = V(P->gravity[0], P->gravity[1], P->gravity[2]);
    for (uint32_t id = 0; id < P->numParticles; id++) {
        v3 v = add(ld3(velocities, id), muls(g, dt));
        st3(velocities, id, v);
        st3(predicted, id, add(ld3(positions, id), muls(v, dt)));
    }
}

static inline void acc3(float* deltas, size_t i, v3 c)
{
    deltas[3 * i] += c.x; deltas[3 * i + 1] += c.y; deltas[3 * i + 2] += c.z;
}

/* VtClothSolverGPU.cu L65-102 (bilateral, counts bumped even when w == 0) */
void o1_solve_stretch(float* predicted, float* deltas, int* deltaCounts, const int* stretchIndices,
                      const float* stretchLengths, const float* invMasses, uint32_t n)
{
    for (uint32_t id = 0; id < n; id++) {
        int idx1 = stretchIndices[2 * id], idx2 = stretchIndices[2 * id + 1];
        float expected = stretchLengths[id];
        v3 diff = sub(ld3(predicted, (size_t)idx1), ld3(predicted, (size_t)idx2));
        float distance = len3(diff);
        float w1 = invMasses[idx1], w2 = invMasses[idx2];
        if (distance != expected && w1 + w2 > 0) {
            v3 gradient = divs(diff, distance + O1_EPSILON);
            float denom = w1 + w2;
            float lambda = (distance - expected) / denom;
            v3 common = muls(gradient, lambda);
            acc3(deltas, (size_t)idx1, muls(common, -w1));
            acc3(deltas, (size_t)idx2, muls(common, w2));
            deltaCounts[idx1] += 1;
            deltaCounts[idx2] += 1;
        }
    }
}

static inline int f2i(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float i2f(int i) { float f; memcpy(&f, &i, 4); return f; }

/* acos(float) of VtClothSolverGPU.cu L163.  The reference calls CUDA's acosf (<= 2 ulp, unspecified last bit);
 * the oracle and the product both use the published fdlibm e_acosf.c algorithm (< 1 ulp, only + - * / sqrt) so
 * that the bending constraint is bit-reproducible between CPU and GPU. */
static inline float acos_p(float z)
{ /* Horner with one rounding per step */
    const float pS0 = 1.6666667163e-01f, pS1 = -3.2556581497e-01f, pS2 = 2.0121252537e-01f, pS3 = -4.0055535734e-02f,
                pS4 = 7.9153501429e-04f, pS5 = 3.4793309169e-05f;
    return z * fmaf(z, fmaf(z, fmaf(z, fmaf(z, fmaf(z, pS5, pS4), pS3), pS2), pS1), pS0);
}
static inline float acos_q(float z)
{
    const float qS1 = -2.4033949375e+00f, qS2 = 2.0209457874e+00f, qS3 = -6.8828397989e-01f, qS4 = 7.7038154006e-02f;
    return fmaf(z, fmaf(z, fmaf(z, fmaf(z, qS4, qS3), qS2), qS1), 1.0f);
}
float o1_acosf(float x)
{
    const float one = 1.0000000000e+00f, pi = 3.1415925026e+00f, pio2_hi = 1.5707962513e+00f, pio2_lo = 7.5497894159e-08f;
    const int hx = f2i(x);
    const int ix = hx & 0x7fffffff;
    if (ix == 0x3f800000) return hx > 0 ? 0.0f : pi + 2.0f * pio2_lo;
    if (ix > 0x3f800000) return (x - x) / (x - x);
    if (ix < 0x3f000000) {
        if (ix <= 0x23000000) return pio2_hi + pio2_lo;
        const float z = x * x;
        const float p = acos_p(z);
        const float q = acos_q(z);
        const float r = p / q;
        return pio2_hi - (x - fmaf(-x, r, pio2_lo));
    }
    if (hx < 0) {
        const float z = (one + x) * 0.5f;
        const float p = acos_p(z);
        const float q = acos_q(z);
        const float s = sqrtf(z);
        const float r = p / q;
        const float w = fmaf(r, s, -pio2_lo);
        return pi - 2.0f * (s + w);
    }
    const float z = (one - x) * 0.5f;
    const float s = sqrtf(z);
    const float df = i2f(f2i(s) & (int)0xfffff000);
    const float c = fmaf(-df, df, z) / (s + df);
    const float p = acos_p(z);
    const float q = acos_q(z);
    const float r = p / q;
    const float w = fmaf(r, s, c);
    return 2.0f * (df + w);
}

/* VtClothSolverGPU.cu L117-189 */
void o1_solve_bending(const O1SimParams* P, float* predicted, float* deltas, int* deltaCounts,
                      const uint32_t* bendIndices, const float* bendAngles, const float* invMass,
                      uint32_t n, float dt)
{
    for (uint32_t id = 0; id < n; id++) {
        uint32_t i0 = bendIndices[4 * id], i1 = bendIndices[4 * id + 1];
        uint32_t i2 = bendIndices[4 * id + 2], i3 = bendIndices[4 * id + 3];
        float rest = bendAngles[id];
        float w0 = invMass[i0], w1 = invMass[i1], w2 = invMass[i2], w3 = invMass[i3];
        v3 p0 = ld3(predicted, i0), p1 = ld3(predicted, i1), p2 = ld3(predicted, i2), p3 = ld3(predicted, i3);

        v3 e = sub(p3, p2);
        float elen = len3(e);
        if (elen < O1_EPSILON) continue;
        float invElen = 1.0f / elen;

        v3 n1 = cross3(sub(p2, p0), sub(p3, p0)); n1 = divs(n1, dot3(n1, n1));
        v3 n2 = cross3(sub(p3, p1), sub(p2, p1)); n2 = divs(n2, dot3(n2, n2));

        v3 d0 = muls(n1, elen);
        v3 d1 = muls(n2, elen);
        v3 d2 = lincomb3(dot3(sub(p0, p3), e) * invElen, n1, dot3(sub(p1, p3), e) * invElen, n2);
        v3 d3 = lincomb3(dot3(sub(p2, p0), e) * invElen, n1, dot3(sub(p2, p1), e) * invElen, n2);

        n1 = normalize3(n1);
        n2 = normalize3(n2);
        float d = clampf(dot3(n1, n2), -1.0f, 1.0f);
        float phi = o1_acosf(d);

        float lambda = wsum4(w0, dot3(d0, d0), w1, dot3(d1, d1), w2, dot3(d2, d2), w3, dot3(d3, d3));
        if (lambda < O1_EPSILON) continue;

        float xpbd_bend = P->bendCompliance / dt / dt;
        lambda = (phi - rest) / (lambda + xpbd_bend);
        if (dot3(cross3(n1, n2), e) > 0.0f) lambda = -lambda;

        acc3(deltas, i0, muls(d0, -w0 * lambda));
        acc3(deltas, i1, muls(d1, -w1 * lambda));
        acc3(deltas, i2, muls(d2, -w2 * lambda));
        acc3(deltas, i3, muls(d3, -w3 * lambda));
        deltaCounts[i0] += 1; deltaCounts[i1] += 1; deltaCounts[i2] += 1; deltaCounts[i3] += 1;
    }
}

/* VtClothSolverGPU.cu L205-235 */
void o1_solve_attachment(const O1SimParams* P, float* predicted, float* deltas, int* deltaCounts,
                         const float* invMass, const int* attachParticleIDs, const int* attachSlotIDs,
                         const float* attachSlotPositions, const float* attachDistances, int n)
{
    for (int id = 0; id < n; id++) {
        uint32_t pid = (uint32_t)attachParticleIDs[id];
        v3 slotPos = ld3(attachSlotPositions, (size_t)attachSlotIDs[id]);
        float targetDist = attachDistances[id] * P->longRangeStretchiness;
        if (invMass[pid] == 0 && targetDist > 0) continue;
        v3 diff = sub(ld3(predicted, pid), slotPos);
        float dist = len3(diff);
        if (dist > targetDist) {
            v3 correction = add(neg(diff), muls(divs(diff, dist), targetDist));
            acc3(deltas, pid, correction);
            deltaCounts[pid] += 1;
        }
    }
}

/* VtClothSolverGPU.cu L253-264 */
void o1_apply_deltas(const O1SimParams* P, float* predicted, float* deltas, int* deltaCounts)
{
    for (uint32_t id = 0; id < P->numParticles; id++) {
        float count = (float)deltaCounts[id];
        if (count > 0) {
            st3(predicted, id, add(ld3(predicted, id), muls(divs(ld3(deltas, id), count), P->relaxationFactor)));
            st3(deltas, id, V(0, 0, 0));
            deltaCounts[id] = 0;
        }
    }
}

/* VtClothSolverGPU.cu L272-287 */
static v3 compute_friction(const O1SimParams* P, v3 correction, v3 relVel)
{
    v3 friction = V(0, 0, 0);
    float correctionLength = len3(correction);
    if (P->friction > 0 && correctionLength > 0) {
        v3 norm = divs(correction, correctionLength);
        v3 tanVel = sub(relVel, muls(norm, dot3(relVel, norm)));
        float tanLength = len3(tanVel);
        float maxTanLength = correctionLength * P->friction;
        friction = muls(neg(tanVel), fminf(maxTanLength / tanLength, 1.0f));
    }
    return friction;
}

static inline float sgnf(float v) { return (v > 0) ? 1.0f : (v < 0 ? -1.0f : 0.0f); }

/* VtClothSolverGPU.cuh L22-89 */
static v3 sdf_compute(const O1SDFCollider* c, v3 target, float margin)
{
    if (c->type == 1) { /* Plane */
        float offset = target.y - (c->position[1] + margin);
        if (offset < 0) return V(0, -offset, 0);
    } else if (c->type == 0) { /* Sphere */
        float radius = c->scale[0] + margin;
        v3 diff = sub(target, V(c->position[0], c->position[1], c->position[2]));
        float distance = len3(diff);
        float offset = distance - radius;
        if (offset < 0) {
            v3 direction = divs(diff, distance);
            return muls(direction, -offset);
        }
    } else if (c->type == 2) { /* Cube */
        v3 correction = V(0, 0, 0);
        v3 lp = mat4_mul_point(c->invCurTransform, target, 1.0f);
        v3 cubeSize = add(V(0.5f, 0.5f, 0.5f), V(margin / c->scale[0], margin / c->scale[1], margin / c->scale[2]));
        v3 offset = sub(V(fabsf(lp.x), fabsf(lp.y), fabsf(lp.z)), cubeSize);
        float maxVal = fmaxf(offset.x, fmaxf(offset.y, offset.z));
        float minVal = fminf(offset.x, fminf(offset.y, offset.z));
        float midVal = offset.x + offset.y + offset.z - maxVal - minVal;
        float scalar = 1.0f;
        if (maxVal < 0) {
            float m = 0.03f;
            if (midVal > -m) scalar = 0.2f;
            if (minVal > -m) {
                v3 mask;
                mask.x = offset.x < 0 ? sgnf(lp.x) : 0;
                mask.y = offset.y < 0 ? sgnf(lp.y) : 0;
                mask.z = offset.z < 0 ? sgnf(lp.z) : 0;
                v3 vec = add(offset, V(m, m, m));
                float len = len3(vec);
                if (len < m) correction = muls(mulv(mask, normalize3(vec)), m - len);
            } else if (offset.x == maxVal) {
                correction = V(copysignf(-offset.x, lp.x), 0, 0);
            } else if (offset.y == maxVal) {
                correction = V(0, copysignf(-offset.y, lp.y), 0);
            } else if (offset.z == maxVal) {
                correction = V(0, 0, copysignf(-offset.z, lp.z));
            }
        }
        /* curTransform * scalar * correction == (mat3*scalar)*vec3, glm left-to-right sums */
        const float* m3 = c->curTransform;
        v3 r;
        r.x = (m3[0] * scalar) * correction.x + (m3[3] * scalar) * correction.y + (m3[6] * scalar) * correction.z;
        r.y = (m3[1] * scalar) * correction.x + (m3[4] * scalar) * correction.y + (m3[7] * scalar) * correction.z;
        r.z = (m3[2] * scalar) * correction.x + (m3[5] * scalar) * correction.y + (m3[8] * scalar) * correction.z;
        return r;
    }
    return V(0, 0, 0);
}

/* VtClothSolverGPU.cuh L91-96 */
static v3 sdf_velocity_at(const O1SDFCollider* c, v3 target)
{
    float M[16];
    mat4_mul(c->lastTransform, c->invCurTransform, M);
    v3 lastPos = mat4_mul_point(M, target, 1.0f);
    return divs(sub(target, lastPos), c->deltaTime);
}

/* VtClothSolverGPU.cu L289-314; predicted may alias positions (pre-stabilisation) */
void o1_collide_sdf(const O1SimParams* P, float* predicted, const O1SDFCollider* colliders,
                    const float* positions, uint32_t numColliders, float dt)
{
    if (numColliders == 0) return;
    for (uint32_t id = 0; id < P->numParticles; id++) {
        v3 pos = ld3(positions, id);
        v3 pred = ld3(predicted, id);
        for (uint32_t i = 0; i < numColliders; i++) {
            const O1SDFCollider* c = &colliders[i];
            v3 correction = sdf_compute(c, pred, P->collisionMargin);
            pred = add(pred, correction);
            if (dot3(correction, correction) > 0) {
                v3 relVel = sub(sub(pred, pos), muls(sdf_velocity_at(c, pred), dt));
                pred = add(pred, compute_friction(P, correction, relVel));
            }
        }
        st3(predicted, id, pred);
    }
}

/* VtClothSolverGPU.cu L329-386 (CollideParticles_Kernel then ApplyDeltas_Kernel) */
void o1_collide_particles(const O1SimParams* P, float* deltas, int* deltaCounts, float* predicted,
                          const float* invMasses, const uint32_t* neighbors, const float* positions)
{
    const uint32_t N = P->numParticles;
    for (uint32_t id = 0; id < N; id++) {
        v3 positionDelta = V(0, 0, 0);
        int deltaCount = 0;
        v3 pred_i = ld3(predicted, id);
        v3 vel_i = sub(pred_i, ld3(positions, id));
        float w_i = invMasses[id];
        for (uint64_t nb = id; nb < (uint64_t)N * (uint32_t)P->maxNumNeighbors; nb += N) {
            uint32_t j = neighbors[nb];
            if (j > N) break;
            float w_j = invMasses[j];
            float denom = w_i + w_j;
            if (denom <= 0) continue;
            v3 pred_j = ld3(predicted, j);
            v3 diff = sub(pred_i, pred_j);
            float distance = len3(diff);
            if (distance >= P->particleDiameter) continue;
            v3 gradient = divs(diff, distance + O1_EPSILON);
            float lambda = (distance - P->particleDiameter) / denom;
            v3 common = muls(gradient, lambda);
            deltaCount++;
            positionDelta = sub(positionDelta, muls(common, w_i));
            v3 relativeVelocity = sub(vel_i, sub(pred_j, ld3(positions, j)));
            v3 friction = compute_friction(P, common, relativeVelocity);
            positionDelta = add(positionDelta, muls(friction, w_i));
        }
        st3(deltas, id, positionDelta);
        deltaCounts[id] = deltaCount;
    }
    o1_apply_deltas(P, predicted, deltas, deltaCounts);
}

/* VtClothSolverGPU.cu L388-407 */
void o1_finalize(const O1SimParams* P, float* velocities, float* positions, const float* predicted,
                 float dt)
{
    for (uint32_t id = 0; id < P->numParticles; id++) {
        v3 new_pos = ld3(predicted, id);
        v3 pos = ld3(positions, id);
        v3 raw_vel = divs(sub(new_pos, pos), dt);
        float raw_vel_len = len3(raw_vel);
        if (raw_vel_len > P->maxSpeed) {
            raw_vel = muls(divs(raw_vel, raw_vel_len), P->maxSpeed);
            new_pos = add(pos, muls(raw_vel, dt));
        }
        st3(velocities, id, muls(raw_vel, 1 - P->damping * dt));
        st3(positions, id, new_pos);
    }
}

/* VtClothSolverGPU.cu L419-465; triangle contributions summed in triangle-id order */
void o1_compute_normal(const O1SimParams* P, float* normals, const float* positions,
                       const uint32_t* indices, uint32_t numTriangles)
{
    if (!P->numParticles) return;
    memset(normals, 0, (size_t)P->numParticles * 3 * sizeof(float));
    for (uint32_t id = 0; id < numTriangles; id++) {
        uint32_t a = indices[3 * id], b = indices[3 * id + 1], c = indices[3 * id + 2];
        v3 p1 = ld3(positions, a), p2 = ld3(positions, b), p3 = ld3(positions, c);
        v3 nrm = cross3(sub(p2, p1), sub(p3, p1));
        acc3(normals, a, nrm); acc3(normals, b, nrm); acc3(normals, c, nrm);
    }
    for (uint32_t id = 0; id < P->numParticles; id++) st3(normals, id, normalize3(ld3(normals, id)));
}

/* ------------------------------------------------------------------ spatial hash */

/* SpatialHashGPU.cu L13-16 */
static inline int int_coord(float value, float cellSpacing) { return (int)floorf(value / cellSpacing); }

/* SpatialHashGPU.cu L18-22: wrapping int32 products, C remainder, abs */
static inline int hash_coords(int x, int y, int z, int tableSize)
{
    int32_t h = (int32_t)((uint32_t)x * 92837111u) ^ (int32_t)((uint32_t)y * 689287499u) ^ (int32_t)((uint32_t)z * 283923481u);
    int32_t r = h % tableSize;
    return r < 0 ? -r : r;
}

int o1_hash_position(const float* p3, float cellSpacing, int tableSize)
{
    return hash_coords(int_coord(p3[0], cellSpacing), int_coord(p3[1], cellSpacing), int_coord(p3[2], cellSpacing), tableSize);
}

typedef struct { uint32_t key, val; } kv_t;

/* stable LSD radix sort on the low `bits` bits == cub::DeviceRadixSort::SortPairs(.., 0, maxBit)
 * (SpatialHashGPU.cu L133-157); keys < tableSize <= 2^bits so this is a full stable sort. */
static void stable_sort_pairs(uint32_t* keys, uint32_t* vals, uint32_t n, int bits)
{
    if (n == 0) return;
    kv_t* a = (kv_t*)malloc(sizeof(kv_t) * n);
    kv_t* b = (kv_t*)malloc(sizeof(kv_t) * n);
    for (uint32_t i = 0; i < n; i++) { a[i].key = keys[i]; a[i].val = vals[i]; }
    for (int shift = 0; shift < bits; shift += 8) {
        int nb = bits - shift < 8 ? bits - shift : 8;
        uint32_t mask = (1u << nb) - 1u;
        uint32_t count[257];
        memset(count, 0, sizeof(count));
        for (uint32_t i = 0; i < n; i++) count[((a[i].key >> shift) & mask) + 1]++;
        for (int d = 0; d < 256; d++) count[d + 1] += count[d];
        for (uint32_t i = 0; i < n; i++) b[count[(a[i].key >> shift) & mask]++] = a[i];
        kv_t* t = a; a = b; b = t;
    }
    for (uint32_t i = 0; i < n; i++) { keys[i] = a[i].key; vals[i] = a[i].val; }
    free(a); free(b);
}

static inline float length2_3(v3 v) { return dot3_plain(v, v); } /* Common.cuh L48-51 */

/* SpatialHashGPU.cu L159-196 (H1 L34-42, H2 L133-157, H3 L44-77 + memset L185, H4 L79-130) */
void o1_hash_objects(uint32_t* particleHash, uint32_t* particleIndex, uint32_t* cellStart,
                     uint32_t* cellEnd, uint32_t* neighbors, const float* positions,
                     const float* originalPositions, O1HashParams hp)
{
    const uint32_t N = hp.numObjects;
    if (N == 0) return;
    for (uint32_t id = 0; id < N; id++) {
        particleHash[id] = (uint32_t)o1_hash_position(positions + 3 * (size_t)id, hp.cellSpacing, hp.tableSize);
        particleIndex[id] = id;
    }
    int maxBit = (int)ceil(log2((double)hp.tableSize));
    stable_sort_pairs(particleHash, particleIndex, N, maxBit);

    /* memset(cellStart, 0xff, 4*(tableSize+1)) in the reference writes one element past size();
     * the oracle clears exactly tableSize entries (the extra one is never read). cellEnd is not cleared. */
    memset(cellStart, 0xff, sizeof(uint32_t) * (size_t)hp.tableSize);
    for (uint32_t id = 0; id < N; id++) {
        uint32_t hash = particleHash[id];
        if (id == 0 || hash != particleHash[id - 1]) {
            cellStart[hash] = id;
            if (id > 0) cellEnd[particleHash[id - 1]] = id;
        }
        if (id == N - 1) cellEnd[hash] = id + 1;
    }

    for (uint32_t t = 0; t < N; t++) {
        uint32_t id = particleIndex[t];
        v3 position = ld3(positions, id);
        v3 originalPos = ld3(originalPositions, id);
        int ix = int_coord(position.x, hp.cellSpacing);
        int iy = int_coord(position.y, hp.cellSpacing);
        int iz = int_coord(position.z, hp.cellSpacing);
        uint64_t neighborIndex = id;
        const uint64_t limit = (uint64_t)N * hp.maxNumNeighbors;
        int full = 0;
        for (int x = ix - 1; x <= ix + 1 && !full; x++)
            for (int y = iy - 1; y <= iy + 1 && !full; y++)
                for (int z = iz - 1; z <= iz + 1 && !full; z++) {
                    int h = hash_coords(x, y, z, hp.tableSize);
                    uint32_t start = cellStart[h];
                    if (start == 0xffffffffu) continue;
                    uint32_t end = cellEnd[h];
                    if (start + hp.maxNumNeighbors < end) end = start + hp.maxNumNeighbors;
                    for (uint32_t i = start; i < end; i++) {
                        uint32_t nb = particleIndex[i];
                        if (nb != id &&
                            (length2_3(sub(position, ld3(positions, nb))) < hp.cellSpacing2) &&
                            (length2_3(sub(originalPos, ld3(originalPositions, nb))) > hp.particleDiameter2)) {
                            neighbors[neighborIndex] = nb;
                            neighborIndex += N;
                            if (neighborIndex >= limit) { full = 1; break; }
                        }
                    }
                }
        if (!full && neighborIndex < limit) neighbors[neighborIndex] = 0xffffffffu;
    }
}

/* ------------------------------------------------------------------ host helpers */

/* glm::inverse(mat4): cofactor expansion as published in glm/detail/func_matrix.inl */
void o1_mat4_inverse(const float* m, float* out)
{
#define M(c, r) m[4 * (c) + (r)]
    float c00 = M(2,2) * M(3,3) - M(3,2) * M(2,3);
    float c02 = M(1,2) * M(3,3) - M(3,2) * M(1,3);
    float c03 = M(1,2) * M(2,3) - M(2,2) * M(1,3);
    float c04 = M(2,1) * M(3,3) - M(3,1) * M(2,3);
    float c06 = M(1,1) * M(3,3) - M(3,1) * M(1,3);
    float c07 = M(1,1) * M(2,3) - M(2,1) * M(1,3);
    float c08 = M(2,1) * M(3,2) - M(3,1) * M(2,2);
    float c10 = M(1,1) * M(3,2) - M(3,1) * M(1,2);
    float c11 = M(1,1) * M(2,2) - M(2,1) * M(1,2);
    float c12 = M(2,0) * M(3,3) - M(3,0) * M(2,3);
    float c14 = M(1,0) * M(3,3) - M(3,0) * M(1,3);
    float c15 = M(1,0) * M(2,3) - M(2,0) * M(1,3);
    float c16 = M(2,0) * M(3,2) - M(3,0) * M(2,2);
    float c18 = M(1,0) * M(3,2) - M(3,0) * M(1,2);
    float c19 = M(1,0) * M(2,2) - M(2,0) * M(1,2);
    float c20 = M(2,0) * M(3,1) - M(3,0) * M(2,1);
    float c22 = M(1,0) * M(3,1) - M(3,0) * M(1,1);
    float c23 = M(1,0) * M(2,1) - M(2,0) * M(1,1);

    float F0[4] = {c00, c00, c02, c03}, F1[4] = {c04, c04, c06, c07}, F2[4] = {c08, c08, c10, c11};
    float F3[4] = {c12, c12, c14, c15}, F4[4] = {c16, c16, c18, c19}, F5[4] = {c20, c20, c22, c23};
    float V0[4] = {M(1,0), M(0,0), M(0,0), M(0,0)}, V1[4] = {M(1,1), M(0,1), M(0,1), M(0,1)};
    float V2[4] = {M(1,2), M(0,2), M(0,2), M(0,2)}, V3[4] = {M(1,3), M(0,3), M(0,3), M(0,3)};
    static const float SA[4] = {+1, -1, +1, -1}, SB[4] = {-1, +1, -1, +1};
    float inv[16];
    for (int k = 0; k < 4; k++) {
        inv[0 + k]  = ((V1[k] * F0[k] - V2[k] * F1[k]) + V3[k] * F2[k]) * SA[k];
        inv[4 + k]  = ((V0[k] * F0[k] - V2[k] * F3[k]) + V3[k] * F4[k]) * SB[k];
        inv[8 + k]  = ((V0[k] * F1[k] - V1[k] * F3[k]) + V3[k] * F5[k]) * SA[k];
        inv[12 + k] = ((V0[k] * F2[k] - V1[k] * F4[k]) + V2[k] * F5[k]) * SB[k];
    }
    float d0 = M(0,0) * inv[0], d1 = M(0,1) * inv[4], d2 = M(0,2) * inv[8], d3 = M(0,3) * inv[12];
    float det = (d0 + d1) + (d2 + d3);
    float ood = 1.0f / det;
    for (int k = 0; k < 16; k++) out[k] = inv[k] * ood;
#undef M
}

/* glm::rotate(m, angle, axis) for a unit axis (ext/matrix_transform.inl) */
static void glm_rotate(float* m, float angle, const float* axis)
{
    float c = cosf(angle), s = sinf(angle);
    float temp[3] = {(1.0f - c) * axis[0], (1.0f - c) * axis[1], (1.0f - c) * axis[2]};
    float R[3][3];
    R[0][0] = c + temp[0] * axis[0]; R[0][1] = temp[0] * axis[1] + s * axis[2]; R[0][2] = temp[0] * axis[2] - s * axis[1];
    R[1][0] = temp[1] * axis[0] - s * axis[2]; R[1][1] = c + temp[1] * axis[1]; R[1][2] = temp[1] * axis[2] + s * axis[0];
    R[2][0] = temp[2] * axis[0] + s * axis[1]; R[2][1] = temp[2] * axis[1] - s * axis[0]; R[2][2] = c + temp[2] * axis[2];
    float out[16];
    for (int col = 0; col < 3; col++)
        for (int r = 0; r < 4; r++)
            out[4 * col + r] = (m[0 + r] * R[col][0] + m[4 + r] * R[col][1]) + m[8 + r] * R[col][2];
    for (int r = 0; r < 4; r++) out[12 + r] = m[12 + r];
    memcpy(m, out, sizeof(out));
}

/* Transform.hpp L22-29 + Helper.cpp L8-15: T * Ry * Rz * Rx * S */
void o1_transform_matrix(const float* position3, const float* rotationDeg3, const float* scale3, float* out16)
{
    float m[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    float t[4];
    for (int r = 0; r < 4; r++)
        t[r] = ((m[0 + r] * position3[0] + m[4 + r] * position3[1]) + m[8 + r] * position3[2]) + m[12 + r];
    for (int r = 0; r < 4; r++) m[12 + r] = t[r];
    const float k = 0.01745329251994329576923690768489f; /* glm::radians */
    static const float ay[3] = {0, 1, 0}, az[3] = {0, 0, 1}, ax[3] = {1, 0, 0};
    glm_rotate(m, rotationDeg3[1] * k, ay);
    glm_rotate(m, rotationDeg3[2] * k, az);
    glm_rotate(m, rotationDeg3[0] * k, ax);
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 4; r++) m[4 * c + r] = m[4 * c + r] * scale3[c];
    memcpy(out16, m, sizeof(m));
}

/* VtClothSolverGPU.hpp L195-203 */
void o1_make_collider(int type, const float* position3, const float* scale3, const float* cur16,
                      const float* last16, float deltaTime, O1SDFCollider* out)
{
    memset(out, 0, sizeof(*out));
    out->type = type;
    memcpy(out->position, position3, 12);
    memcpy(out->scale, scale3, 12);
    out->deltaTime = deltaTime;
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++) out->curTransform[3 * c + r] = cur16[4 * c + r];
    o1_mat4_inverse(cur16, out->invCurTransform);
    memcpy(out->lastTransform, last16, 64);
}

/* Scene.hpp L131-168 */
void o1_generate_cloth_mesh(int resolution, float* vertices, uint32_t* indices)
{
    const float clothSize = 2.0f;
    size_t k = 0;
    for (int y = 0; y <= resolution; y++)
        for (int x = 0; x <= resolution; x++) {
            vertices[k++] = clothSize * ((float)x / (float)resolution - 0.5f);
            vertices[k++] = clothSize * (-(float)y / (float)resolution);
            vertices[k++] = clothSize * 0.0f;
        }
    k = 0;
    const uint32_t S = (uint32_t)resolution + 1;
    for (uint32_t x = 0; x < (uint32_t)resolution; x++)
        for (uint32_t y = 0; y < (uint32_t)resolution; y++) {
            indices[k++] = x * S + y;       indices[k++] = (x + 1) * S + y; indices[k++] = x * S + y + 1;
            indices[k++] = x * S + y + 1;   indices[k++] = (x + 1) * S + y; indices[k++] = (x + 1) * S + y + 1;
        }
}

/* ------------------------------------------------------------------ solver object */
typedef struct { void* p; size_t n, cap, esz; } vec_t;
static void vec_init(vec_t* v, size_t esz) { v->p = NULL; v->n = v->cap = 0; v->esz = esz; }
static void vec_reserve(vec_t* v, size_t n)
{
    if (n > v->cap) { size_t c = n * 3 / 2 + 8; v->p = realloc(v->p, c * v->esz); v->cap = c; }
}
static void* vec_push(vec_t* v, const void* e)
{
    vec_reserve(v, v->n + 1);
    void* d = (char*)v->p + v->n * v->esz;
    memcpy(d, e, v->esz);
    v->n++;
    return d;
}
static void vec_resize0(vec_t* v, size_t n)
{
    vec_reserve(v, n);
    if (n > v->n) memset((char*)v->p + v->n * v->esz, 0, (n - v->n) * v->esz);
    v->n = n;
}

struct O1Solver {
    O1SimParams P;
    vec_t positions, normals, indices, velocities, predicted, deltas, deltaCounts, invMasses;
    vec_t stretchIndices, stretchLengths, bendIndices, bendAngles;
    vec_t attachParticleIDs, attachSlotIDs, attachDistances, attachSlotPositions;
    O1SDFCollider* colliders; int numColliders;
    /* SpatialHashGPU */
    float hashSpacing; int hashTableSize;
    vec_t neighbors, initialPositions, particleHash, particleIndex, cellStart, cellEnd;
};

O1Solver* o1_solver_create(const O1SimParams* params)
{
    O1Solver* s = (O1Solver*)calloc(1, sizeof(O1Solver));
    if (params) s->P = *params; else o1_default_params(&s->P);
    s->P.numParticles = 0; /* VtClothSolverGPU.hpp L29 */
    vec_t* f4[] = {&s->positions, &s->normals, &s->indices, &s->velocities, &s->predicted, &s->deltas,
                   &s->deltaCounts, &s->invMasses, &s->stretchIndices, &s->stretchLengths, &s->bendIndices,
                   &s->bendAngles, &s->attachParticleIDs, &s->attachSlotIDs, &s->attachDistances,
                   &s->attachSlotPositions, &s->neighbors, &s->initialPositions, &s->particleHash,
                   &s->particleIndex, &s->cellStart, &s->cellEnd};
    for (size_t i = 0; i < sizeof(f4) / sizeof(f4[0]); i++) vec_init(f4[i], 4);
    return s;
}

void o1_solver_destroy(O1Solver* s)
{
    if (!s) return;
    vec_t* f4[] = {&s->positions, &s->normals, &s->indices, &s->velocities, &s->predicted, &s->deltas,
                   &s->deltaCounts, &s->invMasses, &s->stretchIndices, &s->stretchLengths, &s->bendIndices,
                   &s->bendAngles, &s->attachParticleIDs, &s->attachSlotIDs, &s->attachDistances,
                   &s->attachSlotPositions, &s->neighbors, &s->initialPositions, &s->particleHash,
                   &s->particleIndex, &s->cellStart, &s->cellEnd};
    for (size_t i = 0; i < sizeof(f4) / sizeof(f4[0]); i++) free(f4[i]->p);
    free(s->colliders);
    free(s);
}

O1SimParams* o1_solver_params(O1Solver* s) { return &s->P; }

/* VtClothSolverGPU.hpp L114-156 (+ SpatialHashGPU ctor/SetInitialPositions, SpatialHashGPU.hpp L18-39) */
int o1_solver_add_cloth(O1Solver* s, const float* vertices, int numVertices, const uint32_t* indices,
                        int numIndices, const float* model16, float particleDiameter)
{
    int prev = (int)s->P.numParticles;
    const float fixedDt = 1.0f / 60.0f; /* Timer.hpp L235 */
    s->P.numParticles += (uint32_t)numVertices;
    s->P.particleDiameter = particleDiameter;
    s->P.deltaTime = fixedDt;
    s->P.maxSpeed = 2 * particleDiameter / fixedDt * s->P.numSubsteps;
    const size_t N = s->P.numParticles;

    vec_resize0(&s->positions, 3 * N);
    memcpy((float*)s->positions.p + 3 * (size_t)prev, vertices, sizeof(float) * 3 * (size_t)numVertices);
    vec_resize0(&s->normals, 3 * N);
    for (int i = 0; i < numIndices; i++) { uint32_t v = indices[i] + (uint32_t)prev; vec_push(&s->indices, &v); }
    vec_resize0(&s->velocities, 3 * N);
    vec_resize0(&s->predicted, 3 * N);
    vec_resize0(&s->deltas, 3 * N);
    vec_resize0(&s->deltaCounts, N);
    size_t oldn = s->invMasses.n;
    vec_resize0(&s->invMasses, N);
    for (size_t i = oldn; i < N; i++) ((float*)s->invMasses.p)[i] = 1.0f;

    o1_initialize_positions((float*)s->positions.p, prev, numVertices, model16);

    s->hashSpacing = particleDiameter * s->P.hashCellSizeScalar;
    s->hashTableSize = 2 * (int)N;
    vec_resize0(&s->neighbors, N * (size_t)s->P.maxNumNeighbors);
    vec_resize0(&s->particleHash, N);
    vec_resize0(&s->particleIndex, N);
    vec_resize0(&s->cellStart, (size_t)s->hashTableSize);
    vec_resize0(&s->cellEnd, (size_t)s->hashTableSize);
    vec_resize0(&s->initialPositions, 3 * N);
    memcpy(s->initialPositions.p, s->positions.p, sizeof(float) * 3 * N);
    return prev;
}

void o1_solver_add_stretch(O1Solver* s, int idx1, int idx2, float distance)
{ /* VtClothSolverGPU.hpp L158-163 */
    vec_push(&s->stretchIndices, &idx1); vec_push(&s->stretchIndices, &idx2);
    vec_push(&s->stretchLengths, &distance);
}
void o1_solver_add_attach_slot(O1Solver* s, const float* pos3)
{ /* L165-168 */
    vec_push(&s->attachSlotPositions, &pos3[0]); vec_push(&s->attachSlotPositions, &pos3[1]);
    vec_push(&s->attachSlotPositions, &pos3[2]);
}
void o1_solver_add_attach(O1Solver* s, int particleIndex, int slotIndex, float distance)
{ /* L170-176 */
    if (distance == 0) ((float*)s->invMasses.p)[particleIndex] = 0;
    vec_push(&s->attachParticleIDs, &particleIndex); vec_push(&s->attachSlotIDs, &slotIndex);
    vec_push(&s->attachDistances, &distance);
}
void o1_solver_add_bend(O1Solver* s, uint32_t i1, uint32_t i2, uint32_t i3, uint32_t i4, float angle)
{ /* L178-185 */
    vec_push(&s->bendIndices, &i1); vec_push(&s->bendIndices, &i2); vec_push(&s->bendIndices, &i3);
    vec_push(&s->bendIndices, &i4); vec_push(&s->bendAngles, &angle);
}

void o1_solver_set_colliders(O1Solver* s, const O1SDFCollider* colliders, int n)
{
    s->colliders = (O1SDFCollider*)realloc(s->colliders, sizeof(O1SDFCollider) * (size_t)(n > 0 ? n : 1));
    if (n > 0) memcpy(s->colliders, colliders, sizeof(O1SDFCollider) * (size_t)n);
    s->numColliders = n;
}

/* SpatialHashGPU.hpp L41-52 */
void o1_solver_hash(O1Solver* s)
{
    O1HashParams hp;
    hp.numObjects = s->P.numParticles;
    hp.cellSpacing = s->hashSpacing;
    hp.cellSpacing2 = s->hashSpacing * s->hashSpacing;
    hp.tableSize = s->hashTableSize;
    hp.maxNumNeighbors = (uint32_t)s->P.maxNumNeighbors;
    hp.particleDiameter2 = s->P.particleDiameter * s->P.particleDiameter;
    o1_hash_objects((uint32_t*)s->particleHash.p, (uint32_t*)s->particleIndex.p, (uint32_t*)s->cellStart.p,
                    (uint32_t*)s->cellEnd.p, (uint32_t*)s->neighbors.p, (const float*)s->predicted.p,
                    (const float*)s->initialPositions.p, hp);
}

/* VtClothSolverGPU.hpp L56-111 */
void o1_solver_simulate(O1Solver* s)
{
    const O1SimParams* P = &s->P;
    float frameTime = 1.0f / 60.0f;
    float substepTime = frameTime / (float)P->numSubsteps;
    float *positions = (float*)s->positions.p, *predicted = (float*)s->predicted.p;
    float *velocities = (float*)s->velocities.p, *deltas = (float*)s->deltas.p;
    int* deltaCounts = (int*)s->deltaCounts.p;
    const float* invMasses = (const float*)s->invMasses.p;

    o1_collide_sdf(P, positions, s->colliders, positions, (uint32_t)s->numColliders, frameTime);
    for (int substep = 0; substep < P->numSubsteps; substep++) {
        o1_predict_positions(P, predicted, velocities, positions, substepTime);
        if (P->enableSelfCollision) {
            if (substep % P->interleavedHash == 0) o1_solver_hash(s);
            o1_collide_particles(P, deltas, deltaCounts, predicted, invMasses, (const uint32_t*)s->neighbors.p, positions);
        }
        o1_collide_sdf(P, predicted, s->colliders, positions, (uint32_t)s->numColliders, substepTime);
        for (int it = 0; it < P->numIterations; it++) {
            o1_solve_stretch(predicted, deltas, deltaCounts, (const int*)s->stretchIndices.p,
                             (const float*)s->stretchLengths.p, invMasses, (uint32_t)s->stretchLengths.n);
            o1_solve_attachment(P, predicted, deltas, deltaCounts, invMasses, (const int*)s->attachParticleIDs.p,
                                (const int*)s->attachSlotIDs.p, (const float*)s->attachSlotPositions.p,
                                (const float*)s->attachDistances.p, (int)s->attachParticleIDs.n);
            o1_solve_bending(P, predicted, deltas, deltaCounts, (const uint32_t*)s->bendIndices.p,
                             (const float*)s->bendAngles.p, invMasses, (uint32_t)s->bendAngles.n, substepTime);
            o1_apply_deltas(P, predicted, deltas, deltaCounts);
        }
        o1_finalize(P, velocities, positions, predicted, substepTime);
    }
    o1_compute_normal(P, (float*)s->normals.p, positions, (const uint32_t*)s->indices.p, (uint32_t)(s->indices.n / 3));
}

/* VtClothObjectGPU.hpp L43-148 */
int o1_cloth_object_start(O1Solver* s, int resolution, const float* vertices, const uint32_t* indices,
                          const float* model16, const int* attachedIndices, int numAttached)
{
    const int S = resolution + 1;
    const int nv = S * S, ni = 6 * resolution * resolution;
    float diameter = len3_plain(sub(ld3(vertices, 0), ld3(vertices, 1))) * s->P.particleDiameterScalar;
    int off = o1_solver_add_cloth(s, vertices, nv, indices, ni, model16, diameter);

    float* pos = (float*)malloc(sizeof(float) * 3 * (size_t)nv); /* ApplyTransform L67-73 */
    for (int i = 0; i < nv; i++) st3(pos, (size_t)i, mat4_mul_point(model16, ld3(vertices, (size_t)i), 1.0f));

#define VAT(x, y) ((x) * S + (y))
#define DIST(a, b) len3_plain(sub(ld3(pos, (size_t)(a)), ld3(pos, (size_t)(b))))
    for (int x = 0; x < S; x++) /* GenerateStretch L75-116 */
        for (int y = 0; y < S; y++) {
            int a, b;
            if (y != resolution) { a = VAT(x, y); b = VAT(x, y + 1); o1_solver_add_stretch(s, off + a, off + b, DIST(a, b)); }
            if (x != resolution) { a = VAT(x, y); b = VAT(x + 1, y); o1_solver_add_stretch(s, off + a, off + b, DIST(a, b)); }
            if (y != resolution && x != resolution) {
                a = VAT(x, y); b = VAT(x + 1, y + 1); o1_solver_add_stretch(s, off + a, off + b, DIST(a, b));
                a = VAT(x, y + 1); b = VAT(x + 1, y); o1_solver_add_stretch(s, off + a, off + b, DIST(a, b));
            }
        }
    for (int slot = 0; slot < numAttached; slot++) { /* GenerateAttach L134-148 */
        v3 slotPos = ld3(pos, (size_t)attachedIndices[slot]);
        float sp[3] = {slotPos.x, slotPos.y, slotPos.z};
        o1_solver_add_attach_slot(s, sp);
        for (int i = 0; i < nv; i++)
            o1_solver_add_attach(s, off + i, slot, len3_plain(sub(slotPos, ld3(pos, (size_t)i))));
    }
    for (int i = 0; i < ni; i += 6) /* GenerateBending L118-132 */
        o1_solver_add_bend(s, (uint32_t)off + indices[i], (uint32_t)off + indices[i + 5],
                           (uint32_t)off + indices[i + 2], (uint32_t)off + indices[i + 1], 0.0f);
#undef VAT
#undef DIST
    free(pos);
    return off;
}

void* o1_solver_buffer(O1Solver* s, int which, uint64_t* count)
{
    vec_t* v = NULL;
    switch (which) {
    case O1_BUF_POSITIONS: v = &s->positions; break;
    case O1_BUF_NORMALS: v = &s->normals; break;
    case O1_BUF_INDICES: v = &s->indices; break;
    case O1_BUF_VELOCITIES: v = &s->velocities; break;
    case O1_BUF_PREDICTED: v = &s->predicted; break;
    case O1_BUF_DELTAS: v = &s->deltas; break;
    case O1_BUF_DELTACOUNTS: v = &s->deltaCounts; break;
    case O1_BUF_INVMASSES: v = &s->invMasses; break;
    case O1_BUF_STRETCHINDICES: v = &s->stretchIndices; break;
    case O1_BUF_STRETCHLENGTHS: v = &s->stretchLengths; break;
    case O1_BUF_BENDINDICES: v = &s->bendIndices; break;
    case O1_BUF_BENDANGLES: v = &s->bendAngles; break;
    case O1_BUF_ATTACHPARTICLEIDS: v = &s->attachParticleIDs; break;
    case O1_BUF_ATTACHSLOTIDS: v = &s->attachSlotIDs; break;
    case O1_BUF_ATTACHDISTANCES: v = &s->attachDistances; break;
    case O1_BUF_ATTACHSLOTPOSITIONS: v = &s->attachSlotPositions; break;
    case O1_BUF_NEIGHBORS: v = &s->neighbors; break;
    case O1_BUF_INITIALPOSITIONS: v = &s->initialPositions; break;
    case O1_BUF_PARTICLEHASH: v = &s->particleHash; break;
    case O1_BUF_PARTICLEINDEX: v = &s->particleIndex; break;
    case O1_BUF_CELLSTART: v = &s->cellStart; break;
    case O1_BUF_CELLEND: v = &s->cellEnd; break;
    default: if (count) *count = 0; return NULL;
    }
    if (count) *count = v->n;
    return v->p;
}
