// Unit test of velvet::VtBuffer / VtMergedBuffer / VtRegisteredBuffer (velvet_b200/csrc/vt_buffer.hpp) against the surface of
// the reference's VtBuffer.hpp L7-236: push_back, push_back(n, value), reserve with 1.5x growth, exact resize, operator[] on
// the host (managed memory), implicit conversion to T*, use from a kernel, destroy; the merged buffer's ranges and its sync()
// into registered (externally owned) device arrays.  Built and run by tests/test_vt_buffer_gpu.py; exit code 0 = pass.
#include <cstdio>
#include <vector>

#include "vt_buffer.hpp"

using namespace velvet;

#define CHECK(cond)                                                               \
    do {                                                                          \
        if (!(cond)) {                                                            \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);         \
            return 1;                                                             \
        }                                                                         \
    } while (0)

__global__ void add_one(int* p, size_t n)
{
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] += 1;
}

struct V3 {
    float x, y, z;
};

int main()
{
    {  // VtBuffer: growth, host indexing, kernels
        VtBuffer<int> b;
        CHECK(b.size() == 0 && b.capacity() == 0 && b.data() == nullptr);
        b.push_back(7);
        CHECK(b.size() == 1 && b.capacity() >= 1 && b[0] == 7);
        const unsigned g0 = b.generation();
        for (int i = 1; i < 1000; i++) b.push_back(i);
        CHECK(b.size() == 1000 && b[999] == 999 && b[0] == 7);
        CHECK(b.generation() > g0);  // reallocations are visible to a solver that cached the pointer
        const size_t cap = b.capacity();
        b.reserve(cap);  // no-op
        CHECK(b.capacity() == cap);
        b.reserve(cap + 1);  // reference growth: 1.5x of the request
        CHECK(b.capacity() == (cap + 1) * 3 / 2);
        b.push_back(5, -3);
        CHECK(b.size() == 1005 && b[1004] == -3 && b[1000] == -3);
        std::vector<int> more = {11, 12, 13};
        b.push_back(more);
        CHECK(b.size() == 1008 && b[1007] == 13);
        int* raw = b;  // operator T*
        add_one<<<(unsigned)((b.size() + 255) / 256), 256>>>(raw, b.size());
        CHECK(cudaDeviceSynchronize() == cudaSuccess);
        CHECK(b[0] == 8 && b[999] == 1000 && b[1007] == 14);
        b.resize(10);  // shrinking keeps the block
        CHECK(b.size() == 10 && b.capacity() >= 1008);
        b.resize(5000, 42);  // exact growth, fill of the new tail
        CHECK(b.size() == 5000 && b.capacity() == 5000 && b[9] == 10 && b[10] == 42 && b[4999] == 42);
        b.destroy();
        CHECK(b.size() == 0 && b.data() == nullptr);
        b.push_back(1);  // usable again
        CHECK(b.size() == 1 && b[0] == 1);
    }
    {  // device-only placement (hash arrays) and back
        VtBuffer<unsigned> d;
        d.setDeviceOnly();
        d.resize(256);
        CHECK(d.deviceOnly() && d.size() == 256);
        cudaPointerAttributes attr;
        CHECK(cudaPointerGetAttributes(&attr, d.data()) == cudaSuccess && attr.type == cudaMemoryTypeDevice);
        bool threw = false;
        try { d.setManaged(); } catch (const Error&) { threw = true; }
        CHECK(threw);
        d.destroy();
        d.setManaged();
        d.resize(4);
        CHECK(cudaPointerGetAttributes(&attr, d.data()) == cudaSuccess && attr.type == cudaMemoryTypeManaged);
    }
    {  // VtMergedBuffer: one array, one range per cloth; sync() mirrors ranges into registered arrays
        VtMergedBuffer<V3> m;
        std::vector<V3> a(100, V3{1, 2, 3}), c(50, V3{4, 5, 6});
        CHECK(m.registerNewBuffer(a.data(), a.size()) == 0);
        CHECK(m.registerNewBuffer(c.data(), c.size()) == 100);
        CHECK(m.registerNewBuffer(nullptr, 10) == 150);
        CHECK(m.size() == 160 && m.numRanges() == 3 && m.rangeOffset(1) == 100 && m.rangeCount(1) == 50);
        CHECK(m[0].x == 1 && m[100].y == 5 && m[155].z == 0);
        V3 *r0 = nullptr, *r1 = nullptr;
        CHECK(cudaMalloc(&r0, 100 * sizeof(V3)) == cudaSuccess && cudaMalloc(&r1, 50 * sizeof(V3)) == cudaSuccess);
        m.attachRegistered(0, r0);
        m.attachRegistered(1, r1);
        CHECK(m.numRegistered() == 2);
        m[3] = V3{9, 9, 9};
        m[149] = V3{8, 8, 8};
        m.sync(0);
        CHECK(cudaDeviceSynchronize() == cudaSuccess);
        std::vector<V3> h0(100), h1(50);
        cudaMemcpy(h0.data(), r0, 100 * sizeof(V3), cudaMemcpyDeviceToHost);
        cudaMemcpy(h1.data(), r1, 50 * sizeof(V3), cudaMemcpyDeviceToHost);
        CHECK(h0[3].x == 9 && h0[4].x == 1 && h1[49].y == 8 && h1[0].z == 6);
        m.attachRegistered(1, nullptr);  // detach
        CHECK(m.numRegistered() == 1);
        bool threw = false;
        try { m.attachRegistered(7, r0); } catch (const Error&) { threw = true; }
        CHECK(threw);
        m.destroy();
        CHECK(m.size() == 0 && m.numRanges() == 0);
        cudaFree(r0);
        cudaFree(r1);
    }
    {  // VtRegisteredBuffer by itself
        VtRegisteredBuffer<float> r;
        float* ext = nullptr;
        CHECK(cudaMalloc(&ext, 64 * sizeof(float)) == cudaSuccess);
        r.registerBuffer(ext, 64);
        CHECK(r.size() == 64 && r.data() == ext && (float*)r == ext);
        r.destroy();  // does not free what it does not own
        CHECK(r.data() == nullptr && r.size() == 0);
        CHECK(cudaFree(ext) == cudaSuccess);
    }
    std::printf("vt_buffer_test: all checks passed\n");
    return 0;
}
