// Force-included before every reference translation unit (nvcc -include).
//  * transitive standard headers the MSVC build got for free (Timer.hpp / VtBuffer.hpp rely on them);
//  * `using namespace std` + `uint` as in the reference's precompiled environment;
//  * the reference sorts IN PLACE with cub::DeviceRadixSort::SortPairs(keys, keys, vals, vals) (SpatialHashGPU.cu
//    L145-156), which CUB 2.x documents as unsupported (its onesweep passes would race).  The macro below routes that one
//    call through an out-of-place wrapper (sort into scratch carved from the temp storage, copy back) so that the
//    reference source itself stays unmodified.
#pragma once
#include <cassert>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include <algorithm>
#include <cub/device/device_radix_sort.cuh>
#include <thrust/device_ptr.h>
#include <thrust/transform.h>
#include <thrust/sort.h>  // before the macro below: thrust itself calls cub::DeviceRadixSort

namespace cub {
struct RefOutOfPlaceRadixSort {
    template <class K, class V>
    static cudaError_t SortPairs(void* d_temp, size_t& bytes, const K* kin, K* kout, const V* vin, V* vout, int n,
                                 int begin_bit = 0, int end_bit = sizeof(K) * 8, cudaStream_t st = 0)
    {
        if ((const void*)kin != (const void*)kout)
            return DeviceRadixSort::SortPairs(d_temp, bytes, kin, kout, vin, vout, n, begin_bit, end_bit, st);
        size_t inner = 0;
        cudaError_t e = DeviceRadixSort::SortPairs(nullptr, inner, kin, kout, vin, vout, n, begin_bit, end_bit, st);
        if (e != cudaSuccess) return e;
        const size_t innerPad = (inner + 255) & ~size_t(255);
        const size_t keysPad = (sizeof(K) * (size_t)n + 255) & ~size_t(255);
        const size_t valsPad = (sizeof(V) * (size_t)n + 255) & ~size_t(255);
        if (d_temp == nullptr) { bytes = innerPad + keysPad + valsPad; return cudaSuccess; }
        char* base = (char*)d_temp;
        K* kalt = (K*)(base + innerPad);
        V* valt = (V*)(base + innerPad + keysPad);
        e = DeviceRadixSort::SortPairs(base, inner, kin, kalt, vin, valt, n, begin_bit, end_bit, st);
        if (e != cudaSuccess) return e;
        cudaMemcpyAsync(kout, kalt, sizeof(K) * (size_t)n, cudaMemcpyDeviceToDevice, st);
        return cudaMemcpyAsync(vout, valt, sizeof(V) * (size_t)n, cudaMemcpyDeviceToDevice, st);
    }
};
}  // namespace cub
#define DeviceRadixSort RefOutOfPlaceRadixSort
