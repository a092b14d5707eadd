// Stand-in for the subset of glm (unpinned vcpkg dependency of vitalight/Velvet, not present under /root/reference)
// that the reference's VtClothSolverGPU.{cu,cuh}, SpatialHashGPU.cu, Common.{cuh,hpp} use.  TEST INFRASTRUCTURE:
// only oracle/ref_cuda builds against it.  Evaluation order restates glm >= 0.9.9 (detail/func_geometric.inl,
// type_mat4x4.inl, type_mat3x3.inl): dot = (x*x' + y*y') + z*z', normalize = v * (1/sqrt(dot)),
// mat4*vec4 = (m0*x + m1*y) + (m2*z + m3*w), mat3*vec3 and mat4*mat4 summed left to right.
// Default constructors are trivial (required for `__constant__ VtSimParams`).
#pragma once
#include <cmath>
#include <cuda_runtime.h>

#define GLMS __host__ __device__ inline

namespace glm {

struct vec4;
struct vec3 {
    float x, y, z;
    vec3() = default;
    GLMS explicit vec3(float s) : x(s), y(s), z(s) {}
    GLMS explicit vec3(int s) : x((float)s), y((float)s), z((float)s) {}
    template <class A, class B, class C> GLMS vec3(A a, B b, C c) : x((float)a), y((float)b), z((float)c) {}
    GLMS vec3(const vec4& v);
    GLMS float& operator[](int i) { return (&x)[i]; }
    GLMS const float& operator[](int i) const { return (&x)[i]; }
    GLMS vec3& operator+=(const vec3& b) { x += b.x; y += b.y; z += b.z; return *this; }
    GLMS vec3& operator-=(const vec3& b) { x -= b.x; y -= b.y; z -= b.z; return *this; }
    GLMS vec3& operator/=(float s) { x /= s; y /= s; z /= s; return *this; }
    GLMS vec3& operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
};
struct vec4 {
    float x, y, z, w;
    vec4() = default;
    template <class A, class B, class C, class D> GLMS vec4(A a, B b, C c, D d) : x((float)a), y((float)b), z((float)c), w((float)d) {}
    template <class W> GLMS vec4(const vec3& v, W w_) : x(v.x), y(v.y), z(v.z), w((float)w_) {}
    GLMS float& operator[](int i) { return (&x)[i]; }
    GLMS const float& operator[](int i) const { return (&x)[i]; }
};
GLMS vec3::vec3(const vec4& v) : x(v.x), y(v.y), z(v.z) {}
struct vec2 { float x, y; vec2() = default; GLMS vec2(float a, float b) : x(a), y(b) {} };
struct ivec3 { int x, y, z; ivec3() = default; GLMS ivec3(int a, int b, int c) : x(a), y(b), z(c) {} };

GLMS vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
GLMS vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
GLMS vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
GLMS vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
GLMS vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
GLMS vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }
GLMS vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
GLMS vec3 operator/(float s, const vec3& a) { return vec3(s / a.x, s / a.y, s / a.z); }
GLMS vec3 operator+(const vec3& a, float s) { return vec3(a.x + s, a.y + s, a.z + s); }
GLMS vec4 operator+(const vec4& a, const vec4& b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
GLMS vec4 operator*(const vec4& a, float s) { return vec4(a.x * s, a.y * s, a.z * s, a.w * s); }

GLMS float dot(const vec3& a, const vec3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
GLMS float length(const vec3& a) { return sqrtf(dot(a, a)); }
GLMS vec3 normalize(const vec3& a) { return a * (1.0f / sqrtf(dot(a, a))); }
GLMS vec3 cross(const vec3& x, const vec3& y) { return vec3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
GLMS vec3 abs(const vec3& a) { return vec3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
GLMS float clamp(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

struct mat3 {
    vec3 c[3];
    mat3() = default;
    GLMS vec3& operator[](int i) { return c[i]; }
    GLMS const vec3& operator[](int i) const { return c[i]; }
};
struct mat4 {
    vec4 c[4];
    mat4() = default;
    GLMS vec4& operator[](int i) { return c[i]; }
    GLMS const vec4& operator[](int i) const { return c[i]; }
};
GLMS mat3 operator*(const mat3& m, float s) { mat3 r; r.c[0] = m.c[0] * s; r.c[1] = m.c[1] * s; r.c[2] = m.c[2] * s; return r; }
GLMS vec3 operator*(const mat3& m, const vec3& v)
{
    return vec3(m.c[0].x * v.x + m.c[1].x * v.y + m.c[2].x * v.z, m.c[0].y * v.x + m.c[1].y * v.y + m.c[2].y * v.z,
                m.c[0].z * v.x + m.c[1].z * v.y + m.c[2].z * v.z);
}
GLMS vec4 operator*(const mat4& m, const vec4& v)
{
    const vec4 Mul0 = m.c[0] * v.x, Mul1 = m.c[1] * v.y, Add0 = Mul0 + Mul1;
    const vec4 Mul2 = m.c[2] * v.z, Mul3 = m.c[3] * v.w, Add1 = Mul2 + Mul3;
    return Add0 + Add1;
}
GLMS mat4 operator*(const mat4& a, const mat4& b)
{
    mat4 r;
    for (int i = 0; i < 4; i++) r.c[i] = ((a.c[0] * b.c[i].x + a.c[1] * b.c[i].y) + a.c[2] * b.c[i].z) + a.c[3] * b.c[i].w;
    return r;
}

}  // namespace glm
