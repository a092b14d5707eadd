// radix_sort.cuh -- stable LSD radix sort of (uint key, uint value) pairs, onesweep style.
//
// Replaces the reference's cub::DeviceRadixSort::SortPairs(keys, keys, vals, vals, n, 0, maxBit) call
// (SpatialHashGPU.cu L133-157), which sorts in place (undefined under CUB 2.x) and keeps its scratch in a
// function-static buffer.  Here: one histogram pass over the keys for all digit places, one 256-thread scan,
// then one "onesweep" pass per 8-bit digit (decoupled look-back across tiles, warp-match ranking so the sort
// is stable).  Result is bit-identical to any stable sort on bits [0,endBit).
#pragma once

#include <cuda_runtime.h>

#include "vt_buffer.hpp"

namespace velvet {

class RadixSorter {
public:
    // Number of 8-bit digit passes for keys with `endBit` significant bits.
    static int numPasses(int endBit) { return endBit <= 0 ? 0 : (endBit + 7) / 8; }

    // Sorts n pairs. Input is read from keysA/valsA; passes ping-pong between A and B.
    // Returns 0 when the sorted result is in (keysA, valsA), 1 when it is in (keysB, valsB)
    // (i.e. numPasses(endBit) odd).  Callers that need the result in A with an odd pass count should
    // produce their input in B and swap the arguments.
    int sort(unsigned* keysA, unsigned* valsA, unsigned* keysB, unsigned* valsB, unsigned n, int endBit,
             cudaStream_t stream);

    // Ensures scratch for up to n items (allocation happens outside CUDA-graph capture).
    void reserve(unsigned n);

    // kernels launched by the last sort() (for launch accounting)
    int lastLaunchCount() const { return m_lastLaunches; }

private:
    DeviceBuffer<unsigned> m_scratch;  // [hist 4*256][tileCounter 4][lookback 4*numTiles*256]
    unsigned m_reservedTiles = 0;
    int m_lastLaunches = 0;
};

}  // namespace velvet
