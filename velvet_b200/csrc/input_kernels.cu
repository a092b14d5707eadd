// input_kernels.cu -- see input_kernels.cuh.
#include "input_kernels.cuh"

#include "vt_buffer.hpp"

namespace velvet {
namespace input {

namespace {

constexpr int PB = 256;

__device__ __forceinline__ vec3 cross_plain(vec3 a, vec3 b)
{
    return V3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}

// float -> unsigned whose order is the float order (-0 folded onto +0: the host's `<` sees them as equal)
__device__ __forceinline__ unsigned ordered_bits(float f)
{
    const unsigned b = __float_as_uint(f + 0.0f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float from_ordered_bits(unsigned o)
{
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

__global__ void grab_reset_kernel(GrabState* s)
{
    s->best = ~0ull;
}

// The host loop keeps the FIRST particle of the smallest distanceToView (strict `<`, ascending index): the minimum of the
// packed pair (ordered distance, index).  NaN distances never compare true on the host and are skipped here.
__global__ void __launch_bounds__(PB) grab_pick_kernel(GrabState* __restrict__ s, const float* __restrict__ positions, unsigned n,
                                                       vec3 o, vec3 d, float diameter)
{
    unsigned long long mine = ~0ull;
    for (unsigned i = blockIdx.x * PB + threadIdx.x; i < n; i += gridDim.x * PB) {
        const vec3 rel = load3(positions, i) - o;
        const float distanceToView = dot_plain(d, rel);
        const float distanceToRay = length_plain(cross_plain(d, rel));
        if (distanceToRay < diameter && distanceToView == distanceToView && distanceToView < 3.402823466e+38f) {
            const unsigned long long key = ((unsigned long long)ordered_bits(distanceToView) << 32) | i;
            if (key < mine) mine = key;
        }
    }
    for (int off = 16; off > 0; off >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, mine, off);
        if (other < mine) mine = other;
    }
    if ((threadIdx.x & 31) == 0 && mine != ~0ull) atomicMin(&s->best, mine);
}

__global__ void grab_pin_kernel(GrabState* s, float* invMasses)
{
    const unsigned long long best = s->best;
    if (best == ~0ull) {
        s->index = -1;
        s->distanceToOrigin = 3.402823466e+38f;  // FLT_MAX, what the host loop leaves in minDistanceToView
        return;                                  // (a grab that was in progress stays in progress: L46-47 only set the flag on a hit)
    }
    const int id = (int)(unsigned)(best & 0xffffffffu);
    s->index = id;
    s->distanceToOrigin = from_ordered_bits((unsigned)(best >> 32));
    s->grabbing = 1;
    s->savedInvMass = invMasses[id];
    invMasses[id] = 0.0f;
}

__global__ void drag_kernel(const GrabState* s, float* positions, float* velocities, vec3 o, vec3 d, float fixedDeltaTime)
{
    if (!s->grabbing) return;
    const int id = s->index;
    const vec3 mousePos = o + d * s->distanceToOrigin;
    const vec3 curPos = load3(positions, (size_t)id);
    const float a = 0.8f;  // Helper::Lerp(value1, value2, a) = a * value2 + (1 - a) * value1, Helper.hpp L36-40
    const vec3 target = a * curPos + (1 - a) * mousePos;
    store3(positions, (size_t)id, target);
    store3(velocities, (size_t)id, (target - curPos) / fixedDeltaTime);
}

__global__ void release_kernel(GrabState* s, float* invMasses)
{
    if (!s->grabbing) return;
    s->grabbing = 0;
    invMasses[s->index] = s->savedInvMass;
}

}  // namespace

void grab(GrabState* state, const float* positions, float* invMasses, unsigned numParticles, vec3 rayOrigin, vec3 rayDirection,
          float particleDiameter, cudaStream_t st)
{
    grab_reset_kernel<<<1, 1, 0, st>>>(state);
    if (numParticles) {
        unsigned blocks = (numParticles + PB - 1) / PB;
        if (blocks > 148u * 8u) blocks = 148u * 8u;
        grab_pick_kernel<<<blocks, PB, 0, st>>>(state, positions, numParticles, rayOrigin, rayDirection, particleDiameter);
    }
    grab_pin_kernel<<<1, 1, 0, st>>>(state, invMasses);
    VT_CUDA(cudaGetLastError());
}

void drag(const GrabState* state, float* positions, float* velocities, vec3 rayOrigin, vec3 rayDirection, float fixedDeltaTime,
          cudaStream_t st)
{
    drag_kernel<<<1, 1, 0, st>>>(state, positions, velocities, rayOrigin, rayDirection, fixedDeltaTime);
    VT_CUDA(cudaGetLastError());
}

void release(GrabState* state, float* invMasses, cudaStream_t st)
{
    release_kernel<<<1, 1, 0, st>>>(state, invMasses);
    VT_CUDA(cudaGetLastError());
}

}  // namespace input
}  // namespace velvet
