// setup_kernels.cu -- see setup_kernels.cuh.  Compiled without FMA contraction like every EXACT unit: the rest lengths must
// carry the bits of the host generator (solver.cu: GenerateGridConstraints).
#include "setup_kernels.cuh"

#include <cub/device/device_scan.cuh>

#include "vt_buffer.hpp"

namespace velvet {
namespace setup {

namespace {

constexpr int PB = 256;
inline unsigned blocks_for(size_t n) { return (unsigned)((n + PB - 1) / PB); }

__global__ void __launch_bounds__(PB) fill_words_kernel(unsigned* __restrict__ dst, size_t n, unsigned value)
{
    const size_t i = (size_t)blockIdx.x * PB + threadIdx.x;
    if (i < n) dst[i] = value;
}

__global__ void __launch_bounds__(PB) offset_indices_kernel(unsigned* __restrict__ dst, const unsigned* __restrict__ src, size_t n,
                                                            unsigned offset)
{
    const size_t i = (size_t)blockIdx.x * PB + threadIdx.x;
    if (i < n) dst[i] = src[i] + offset;
}

__global__ void __launch_bounds__(PB) instance_positions_kernel(float* __restrict__ positions, const float* __restrict__ vertices,
                                                                const float* __restrict__ models, unsigned n)
{
    const unsigned i = blockIdx.x * PB + threadIdx.x;
    if (i >= n) return;
    const size_t k = blockIdx.y;
    store3(positions, k * n + i, mul_point(models + 16 * k, load3(vertices, i), 1.0f));
}

__global__ void __launch_bounds__(PB) instance_pin_kernel(float* __restrict__ invMasses, const int* __restrict__ pinned, unsigned numPinned,
                                                          unsigned n, unsigned numInstances)
{
    const size_t t = (size_t)blockIdx.x * PB + threadIdx.x;
    if (t >= (size_t)numPinned * numInstances) return;
    invMasses[(t / numPinned) * n + (size_t)pinned[t % numPinned]] = 0.0f;
}

// where the constraints generated at vertex (x, y) start in the cloth's stretch list
__device__ __forceinline__ size_t stretch_slot(int x, int y, int R)
{
    return x < R ? (size_t)x * (4 * (size_t)R + 1) + 4 * (size_t)y : (size_t)R * (4 * (size_t)R + 1) + (size_t)y;
}

// one thread per vertex: its up-to-four constraints are consecutive in the list
__global__ void __launch_bounds__(PB) generate_stretch_kernel(int* __restrict__ idx, float* __restrict__ len,
                                                              const float* __restrict__ world, unsigned base, int R)
{
    const int side = R + 1;
    const unsigned v = blockIdx.x * PB + threadIdx.x;
    if (v >= (unsigned)side * (unsigned)side) return;
    const int x = (int)(v / (unsigned)side), y = (int)(v % (unsigned)side);
    size_t s = stretch_slot(x, y, R);
    const unsigned g = base + v;
    auto emit = [&](unsigned a, unsigned b) {
        idx[2 * s] = (int)a;
        idx[2 * s + 1] = (int)b;
        len[s] = length_plain(load3(world, a) - load3(world, b));
        s++;
    };
    if (y != R) emit(g, g + 1);
    if (x != R) emit(g, g + (unsigned)side);
    if (y != R && x != R) {
        emit(g, g + (unsigned)side + 1);
        emit(g + 1, g + (unsigned)side);
    }
}

__global__ void __launch_bounds__(PB) generate_attach_kernel(int* __restrict__ pid, int* __restrict__ slot, float* __restrict__ dist,
                                                             const float* __restrict__ world, float* __restrict__ invMass,
                                                             unsigned base, unsigned nv, int slotId, vec3 slotPos)
{
    const unsigned i = blockIdx.x * PB + threadIdx.x;
    if (i >= nv) return;
    const float d = length_plain(slotPos - load3(world, base + i));
    pid[i] = (int)(base + i);
    slot[i] = slotId;
    dist[i] = d;
    if (d == 0) invMass[base + i] = 0;
}

__global__ void __launch_bounds__(PB) generate_bend_kernel(unsigned* __restrict__ quads, float* __restrict__ angles,
                                                           const unsigned* __restrict__ mesh, size_t numQuads)
{
    const size_t q = (size_t)blockIdx.x * PB + threadIdx.x;
    if (q >= numQuads) return;
    const unsigned* t = mesh + 6 * q;
    reinterpret_cast<uint4*>(quads)[q] = make_uint4(t[0], t[5], t[2], t[1]);
    angles[q] = 0.0f;
}

__global__ void __launch_bounds__(PB) grid_plan_kernel(float4* __restrict__ rest4, const float* __restrict__ len,
                                                       const unsigned* __restrict__ bend, unsigned base, int R, int* mismatch)
{
    const int side = R + 1;
    const unsigned v = blockIdx.x * PB + threadIdx.x;
    if (v >= (unsigned)side * (unsigned)side) return;
    const int x = (int)(v / (unsigned)side), y = (int)(v % (unsigned)side);
    size_t s = stretch_slot(x, y, R);
    float4 r = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (y != R) r.x = len[s++];
    if (x != R) r.y = len[s++];
    if (y != R && x != R) {
        r.z = len[s];
        r.w = len[s + 1];
        const uint4 q = reinterpret_cast<const uint4*>(bend)[(size_t)x * R + y];
        const unsigned g = base + v;
        if (q.x != g || q.y != g + (unsigned)side + 1 || q.z != g + 1 || q.w != g + (unsigned)side) atomicOr(mismatch, 1);
    }
    rest4[base + v] = r;
}

__global__ void __launch_bounds__(PB) grid_attach_kernel(unsigned* __restrict__ attOff, uint2* __restrict__ rec,
                                                         const float* __restrict__ dist, unsigned base, unsigned nv, unsigned numSlots,
                                                         unsigned firstSlot, unsigned attBase)
{
    const unsigned i = blockIdx.x * PB + threadIdx.x;
    if (i >= nv) return;
    const unsigned first = attBase + i * numSlots;
    attOff[base + i] = first;
    for (unsigned s = 0; s < numSlots; s++) rec[first + s] = make_uint2(firstSlot + s, __float_as_uint(dist[(size_t)s * nv + i]));
}

__global__ void __launch_bounds__(PB) count_incidence_kernel(const unsigned* __restrict__ indices, size_t n, unsigned nv,
                                                             unsigned* __restrict__ count /* off + 1 */, int* badIndex)
{
    const size_t i = (size_t)blockIdx.x * PB + threadIdx.x;
    if (i >= n) return;
    const unsigned v = indices[i];
    if (v >= nv) {
        atomicOr(badIndex, 1);
        return;
    }
    atomicAdd(count + v, 1u);
}

__global__ void __launch_bounds__(PB) place_incidence_kernel(const unsigned* __restrict__ indices, size_t n, unsigned nv,
                                                             unsigned* __restrict__ cursor, unsigned* __restrict__ tris)
{
    const size_t i = (size_t)blockIdx.x * PB + threadIdx.x;
    if (i >= n) return;
    const unsigned v = indices[i];
    if (v >= nv) return;
    tris[atomicAdd(cursor + v, 1u)] = (unsigned)(i / 3);
}

// arrival order of the atomics -> ascending triangle id (a vertex of a grid has at most six triangles)
__global__ void __launch_bounds__(PB) sort_incidence_kernel(const unsigned* __restrict__ off, unsigned nv, unsigned* __restrict__ tris)
{
    const unsigned v = blockIdx.x * PB + threadIdx.x;
    if (v >= nv) return;
    const unsigned a = off[v], b = off[v + 1];
    for (unsigned i = a + 1; i < b; i++) {
        const unsigned t = tris[i];
        unsigned j = i;
        for (; j > a && tris[j - 1] > t; j--) tris[j] = tris[j - 1];
        tris[j] = t;
    }
}

}  // namespace

void fill_words(void* dst, size_t words, unsigned value, cudaStream_t st)
{
    if (!words) return;
    fill_words_kernel<<<blocks_for(words), PB, 0, st>>>(static_cast<unsigned*>(dst), words, value);
    VT_CUDA(cudaGetLastError());
}

void offset_indices(unsigned* dst, const unsigned* src, size_t n, unsigned offset, cudaStream_t st)
{
    if (!n) return;
    offset_indices_kernel<<<blocks_for(n), PB, 0, st>>>(dst, src, n, offset);
    VT_CUDA(cudaGetLastError());
}

void instance_positions(float* positions, const float* vertices, const float* models16, unsigned n, unsigned numInstances, cudaStream_t st)
{
    if (!n || !numInstances) return;
    instance_positions_kernel<<<dim3(blocks_for(n), numInstances), PB, 0, st>>>(positions, vertices, models16, n);
    VT_CUDA(cudaGetLastError());
}

void instance_pin(float* invMasses, const int* pinned, unsigned numPinned, unsigned n, unsigned numInstances, cudaStream_t st)
{
    if (!numPinned || !numInstances) return;
    instance_pin_kernel<<<blocks_for((size_t)numPinned * numInstances), PB, 0, st>>>(invMasses, pinned, numPinned, n, numInstances);
    VT_CUDA(cudaGetLastError());
}

void generate_stretch(int* idxPairs, float* lengths, const float* worldPositions, unsigned base, int R, cudaStream_t st)
{
    const size_t nv = (size_t)(R + 1) * (R + 1);
    generate_stretch_kernel<<<blocks_for(nv), PB, 0, st>>>(idxPairs, lengths, worldPositions, base, R);
    VT_CUDA(cudaGetLastError());
}

void generate_attach(int* particleIds, int* slotIds, float* distances, const float* worldPositions, float* invMasses, unsigned base,
                     unsigned numVertices, int slotId, vec3 slotPosition, cudaStream_t st)
{
    generate_attach_kernel<<<blocks_for(numVertices), PB, 0, st>>>(particleIds, slotIds, distances, worldPositions, invMasses, base,
                                                                   numVertices, slotId, slotPosition);
    VT_CUDA(cudaGetLastError());
}

void generate_bend(unsigned* idxQuads, float* angles, const unsigned* shiftedMeshIndices, size_t numQuads, cudaStream_t st)
{
    if (!numQuads) return;
    generate_bend_kernel<<<blocks_for(numQuads), PB, 0, st>>>(idxQuads, angles, shiftedMeshIndices, numQuads);
    VT_CUDA(cudaGetLastError());
}

void grid_plan_from_lists(float4* rest4, const float* clothStretchLengths, const unsigned* clothBendIndices, unsigned base, int R,
                          int* mismatch, cudaStream_t st)
{
    const size_t nv = (size_t)(R + 1) * (R + 1);
    grid_plan_kernel<<<blocks_for(nv), PB, 0, st>>>(rest4, clothStretchLengths, clothBendIndices, base, R, mismatch);
    VT_CUDA(cudaGetLastError());
}

void grid_attach_records(unsigned* attOff, uint2* attachRec, const float* clothAttachDistances, unsigned base, unsigned numVertices,
                         unsigned numSlots, unsigned firstSlot, unsigned attBase, cudaStream_t st)
{
    grid_attach_kernel<<<blocks_for(numVertices), PB, 0, st>>>(attOff, attachRec, clothAttachDistances, base, numVertices, numSlots,
                                                               firstSlot, attBase);
    VT_CUDA(cudaGetLastError());
}

void vertex_triangles(const unsigned* indices, size_t numIndices, unsigned numVertices, unsigned* off, unsigned* tris,
                      unsigned* scratch, int* badIndex, cudaStream_t st)
{
    VT_CUDA(cudaMemsetAsync(off, 0, ((size_t)numVertices + 1) * sizeof(unsigned), st));
    if (!numIndices) return;
    count_incidence_kernel<<<blocks_for(numIndices), PB, 0, st>>>(indices, numIndices, numVertices, off + 1, badIndex);
    // off[v + 1] = number of incidences of vertices 0..v
    size_t tempBytes = 0;
    VT_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tempBytes, off + 1, off + 1, (int)numVertices, st));
    void* temp = nullptr;
    VT_CUDA(cudaMallocAsync(&temp, tempBytes ? tempBytes : 4, st));
    VT_CUDA(cub::DeviceScan::InclusiveSum(temp, tempBytes, off + 1, off + 1, (int)numVertices, st));
    VT_CUDA(cudaFreeAsync(temp, st));
    VT_CUDA(cudaMemcpyAsync(scratch, off, (size_t)numVertices * sizeof(unsigned), cudaMemcpyDeviceToDevice, st));
    place_incidence_kernel<<<blocks_for(numIndices), PB, 0, st>>>(indices, numIndices, numVertices, scratch, tris);
    sort_incidence_kernel<<<blocks_for(numVertices), PB, 0, st>>>(off, numVertices, tris);
    VT_CUDA(cudaGetLastError());
}

}  // namespace setup
}  // namespace velvet
