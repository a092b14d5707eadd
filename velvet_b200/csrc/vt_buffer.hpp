// vt_buffer.hpp -- VtBuffer<T> / VtMergedBuffer<T> (reference: VtBuffer.hpp L7-236, Common.cuh L66-78)
// plus DeviceBuffer<T>, the plain-cudaMalloc scratch used by the fused pipeline.
//
// VtBuffer keeps the reference's contract: a growable array in *managed* memory (so callers may index it
// on the host between frames), implicit conversion to T*, push_back / resize / reserve with 1.5x growth,
// destroy().  Differences: errors throw velvet::Error instead of exit(); bulk append is one memcpy; a
// `generation()` counter tells the solver when a pointer changed so it can re-capture its CUDA graph.
#pragma once

#include <cuda_runtime.h>

#include <cassert>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace velvet {

struct Error : std::runtime_error {
    int status;
    Error(int st, const std::string& msg) : std::runtime_error(msg), status(st) {}
};

inline void cuda_check(cudaError_t e, const char* what, const char* file, int line)
{
    if (e != cudaSuccess) {
        throw Error(-2, std::string(what) + " failed: " + cudaGetErrorString(e) + " (" + file + ":" + std::to_string(line) + ")");
    }
}
#define VT_CUDA(expr) ::velvet::cuda_check((expr), #expr, __FILE__, __LINE__)

template <class T>
class VtBuffer {
public:
    VtBuffer() = default;
    explicit VtBuffer(size_t size) { resize(size); }
    VtBuffer(const VtBuffer&) = delete;
    VtBuffer& operator=(const VtBuffer&) = delete;
    ~VtBuffer()
    {
        try { destroy(); } catch (...) {}
    }

    operator T*() const { return m_data; }
    T* data() const { return m_data; }
    size_t size() const { return m_count; }
    size_t capacity() const { return m_capacity; }
    unsigned generation() const { return m_generation; }

    T& operator[](size_t index)
    {
        assert(m_data && index < m_count);
        return m_data[index];
    }

    void push_back(const T& t)
    {
        reserve(m_count + 1);
        m_data[m_count++] = t;
    }
    void push_back(size_t newCount, const T& val)
    {
        reserve(m_count + newCount);
        for (size_t i = 0; i < newCount; i++) m_data[m_count++] = val;
    }
    void push_back(const std::vector<T>& data) { append(data.data(), data.size()); }
    void append(const T* src, size_t n)
    {
        if (!n) return;
        reserve(m_count + n);
        std::memcpy(m_data + m_count, src, n * sizeof(T));
        m_count += n;
    }

    // Appends `n` elements that the DEVICE is going to write (a generation kernel, an async copy): returns the first one.
    // For managed memory the pages of the new range are created on `device` right away, so that neither the writing kernel
    // nor the host faults on them (a host-side push_back of a million elements is a million first-touch stores into managed
    // pages that a prefetch then has to move: the whole cost of registering a cloth).
    T* extendOnDevice(size_t n, int device, cudaStream_t stream)
    {
        reserve(m_count + n);
        T* first = m_data + m_count;
        if (!m_deviceOnly && n) {
            cudaMemPrefetchAsync(first, n * sizeof(T), device, stream);
            (void)cudaGetLastError();  // best effort: without it the first kernel touching the range pays the faults
        }
        m_count += n;
        return first;
    }

    void reserve(size_t minCapacity)
    {
        if (minCapacity <= m_capacity) return;
        grow(minCapacity * 3 / 2);  // growth factor of the reference (push_back / append path)
    }
    // resize() sizes exactly: the big fixed-size arrays (64 neighbour slots per particle, hash tables) are resized once
    // and never pushed to, and 1.5x of 4.3 GB matters when 8 ranks register a 16.7M-particle cloth at the same time.
    void resize(size_t newCount)
    {
        if (newCount > m_capacity) grow(newCount);
        m_count = newCount;
    }
    void resize(size_t newCount, const T& val)
    {
        const size_t first = m_count;
        resize(newCount);
        for (size_t i = first; i < newCount; i++) m_data[i] = val;
    }
    // Device-only placement (plain cudaMalloc) for arrays that only kernels touch (hash tables, neighbour lists).  Must be
    // chosen while the buffer is empty; host indexing, push_back and append are then invalid.
    void setDeviceOnly()
    {
        if (m_data) throw Error(-4, "VtBuffer::setDeviceOnly on an allocated buffer");
        m_deviceOnly = true;
    }
    bool deviceOnly() const { return m_deviceOnly; }
    void setManaged()  // back to the reference's placement (host-indexable); only while empty
    {
        if (m_data) throw Error(-4, "VtBuffer::setManaged on an allocated buffer");
        m_deviceOnly = false;
    }

    void destroy()
    {
        if (m_data) {
            cudaDeviceSynchronize();
            cudaFree(m_data);
            m_generation++;
        }
        m_data = nullptr;
        m_count = m_capacity = 0;
    }

private:
    void grow(size_t newCapacity)
    {
        T* fresh = nullptr;
        const cudaError_t e = m_deviceOnly ? cudaMalloc((void**)&fresh, newCapacity * sizeof(T))
                                           : cudaMallocManaged((void**)&fresh, newCapacity * sizeof(T));
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            throw Error(-2, std::string(m_deviceOnly ? "cudaMalloc" : "cudaMallocManaged") + " of " + std::to_string(newCapacity * sizeof(T)) +
                                " bytes failed: " + cudaGetErrorString(e) + " (VtBuffer::grow)");
        }
        if (m_data) {
            // the GPU may still be reading the old block (async frames): drain before the host touches it
            VT_CUDA(cudaDeviceSynchronize());
            if (m_deviceOnly) VT_CUDA(cudaMemcpy(fresh, m_data, m_count * sizeof(T), cudaMemcpyDeviceToDevice));
            else std::memcpy(fresh, m_data, m_count * sizeof(T));
            VT_CUDA(cudaFree(m_data));
        }
        m_data = fresh;
        m_capacity = newCapacity;
        m_generation++;
    }

    size_t m_count = 0;
    size_t m_capacity = 0;
    T* m_data = nullptr;
    unsigned m_generation = 0;
    bool m_deviceOnly = false;
};

// VtRegisteredBuffer (reference: VtBuffer.hpp L122-186): a device array that somebody else owns -- in the reference a GL vertex
// buffer registered with CUDA and mapped once (registerBuffer(GLuint), L167-179).  Headless, the owner hands over the mapped
// device pointer itself (what cudaGraphicsResourceGetMappedPointer returned in the renderer, or any device allocation);
// with VELVET_GL_INTEROP defined (and GL + cuda_gl_interop.h available to the includer) the GLuint form is compiled as well.
template <class T>
class VtRegisteredBuffer {
public:
    VtRegisteredBuffer() = default;
    VtRegisteredBuffer(const VtRegisteredBuffer&) = delete;
    VtRegisteredBuffer& operator=(const VtRegisteredBuffer&) = delete;
    ~VtRegisteredBuffer() { destroy(); }

    T* data() const { return m_buffer; }
    operator T*() const { return m_buffer; }
    size_t size() const { return m_count; }

    void registerBuffer(T* mappedDevicePointer, size_t count)
    {
        destroy();
        m_buffer = mappedDevicePointer;
        m_count = count;
        m_numBytes = count * sizeof(T);
    }
#ifdef VELVET_GL_INTEROP
    void registerBuffer(GLuint vbo)  // VtBuffer.hpp L167-179, verbatim semantics: register, map once, read the pointer, unmap
    {
        destroy();
        VT_CUDA(cudaGraphicsGLRegisterBuffer(&m_cudaVboResource, vbo, cudaGraphicsRegisterFlagsNone));
        VT_CUDA(cudaGraphicsMapResources(1, &m_cudaVboResource, 0));
        VT_CUDA(cudaGraphicsResourceGetMappedPointer((void**)&m_buffer, &m_numBytes, m_cudaVboResource));
        m_count = m_numBytes / sizeof(T);
        VT_CUDA(cudaGraphicsUnmapResources(1, &m_cudaVboResource, 0));
    }
#endif
    void destroy()
    {
#ifdef VELVET_GL_INTEROP
        if (m_cudaVboResource) cudaGraphicsUnregisterResource(m_cudaVboResource);
        m_cudaVboResource = nullptr;
#endif
        m_buffer = nullptr;  // not ours to free
        m_count = m_numBytes = 0;
    }

private:
    size_t m_count = 0, m_numBytes = 0;
    T* m_buffer = nullptr;
#ifdef VELVET_GL_INTEROP
    struct cudaGraphicsResource* m_cudaVboResource = nullptr;
#endif
};

// Headless VtMergedBuffer: one managed array holding every cloth's range.  The reference mirrors each range
// into a GL VBO (registerNewBuffer(GLuint) / sync(), VtBuffer.hpp L202-229); headless callers register host
// data instead and read results back with cudaMemcpy (velvet_solver_download / readback_async).
template <class T>
class VtMergedBuffer {
public:
    VtMergedBuffer() = default;
    VtMergedBuffer(const VtMergedBuffer&) = delete;
    VtMergedBuffer& operator=(const VtMergedBuffer&) = delete;

    // Appends `count` elements (copied from host `src`, or zero when src == nullptr); returns the offset.
    size_t registerNewBuffer(const T* src, size_t count)
    {
        const size_t offset = m_vbuffer.size();
        m_offsets.push_back(offset);
        m_counts.push_back(count);
        m_vbuffer.reserve(offset + count);  // amortised growth when cloths are registered one by one
        m_vbuffer.resize(offset + count);
        if (src) std::memcpy(m_vbuffer.data() + offset, src, count * sizeof(T));
        else std::memset((void*)(m_vbuffer.data() + offset), 0, count * sizeof(T));
        return offset;
    }
    // Same bookkeeping, contents left to the device (see VtBuffer::extendOnDevice): returns the range's first element.
    T* registerNewBufferOnDevice(size_t count, int device, cudaStream_t stream)
    {
        m_offsets.push_back(m_vbuffer.size());
        m_counts.push_back(count);
        return m_vbuffer.extendOnDevice(count, device, stream);
    }
    // `numRanges` equal ranges of `count` elements each in one go (batched instances): returns the first range's first element.
    T* registerNewBuffersOnDevice(size_t count, size_t numRanges, int device, cudaStream_t stream)
    {
        for (size_t k = 0; k < numRanges; k++) {
            m_offsets.push_back(m_vbuffer.size() + k * count);
            m_counts.push_back(count);
        }
        return m_vbuffer.extendOnDevice(count * numRanges, device, stream);
    }
    size_t size() const { return m_vbuffer.size(); }
    size_t numRanges() const { return m_offsets.size(); }
    size_t rangeOffset(size_t i) const { return m_offsets[i]; }
    size_t rangeCount(size_t i) const { return m_counts[i]; }
    // Attaches a registered (externally owned) device array to range i: sync() then mirrors the range into it, which is what
    // the reference does for the renderer's VBOs (VtBuffer.hpp L202-229).  The array must hold rangeCount(i) elements.
    void attachRegistered(size_t i, T* mappedDevicePointer)
    {
        if (i >= m_offsets.size()) throw Error(-1, "VtMergedBuffer::attachRegistered: no such range");
        if (m_rbuffers.size() < m_offsets.size()) m_rbuffers.resize(m_offsets.size());
        if (!m_rbuffers[i]) m_rbuffers[i] = std::make_shared<VtRegisteredBuffer<T>>();
        if (mappedDevicePointer) m_rbuffers[i]->registerBuffer(mappedDevicePointer, m_counts[i]);
        else m_rbuffers[i]->destroy();
    }
    // copy from the merged array to the registered ones (L222-229), in stream order
    void sync(cudaStream_t stream = 0)
    {
        for (size_t i = 0; i < m_rbuffers.size(); i++)
            if (m_rbuffers[i] && m_rbuffers[i]->data())
                VT_CUDA(cudaMemcpyAsync(m_rbuffers[i]->data(), m_vbuffer.data() + m_offsets[i], m_counts[i] * sizeof(T),
                                        cudaMemcpyDeviceToDevice, stream));
    }
    size_t numRegistered() const
    {
        size_t n = 0;
        for (const auto& r : m_rbuffers) n += (r && r->data()) ? 1 : 0;
        return n;
    }
    void destroy()
    {
        m_vbuffer.destroy();
        m_offsets.clear();
        m_counts.clear();
        m_rbuffers.clear();
    }
    operator T*() const { return m_vbuffer.data(); }
    T* data() const { return m_vbuffer.data(); }
    T& operator[](size_t i) { return m_vbuffer[i]; }
    unsigned generation() const { return m_vbuffer.generation(); }

private:
    std::vector<size_t> m_offsets, m_counts;
    std::vector<std::shared_ptr<VtRegisteredBuffer<T>>> m_rbuffers;
    VtBuffer<T> m_vbuffer;
};

// Plain device scratch (never touched by the host).
template <class T>
class DeviceBuffer {
public:
    DeviceBuffer() = default;
    DeviceBuffer(const DeviceBuffer&) = delete;
    DeviceBuffer& operator=(const DeviceBuffer&) = delete;
    ~DeviceBuffer() { release(); }
    void release()
    {
        if (m_data) cudaFree(m_data);
        m_data = nullptr;
        m_count = 0;
    }
    // (Re)allocates when the size changes; contents are undefined afterwards.
    void allocate(size_t count)
    {
        if (count == m_count && m_data) return;
        release();
        if (count) VT_CUDA(cudaMalloc((void**)&m_data, count * sizeof(T)));
        m_count = count;
    }
    // Grow-only variant for scratch that is reused at changing sizes: never shrinks, grows by at least 1.5x.
    void reserve(size_t count)
    {
        if (count <= m_count && m_data) return;
        allocate(count > m_count + m_count / 2 ? count : m_count + m_count / 2);
    }
    void upload(const T* host, size_t count, cudaStream_t st = 0)
    {
        allocate(count);
        if (count) VT_CUDA(cudaMemcpyAsync(m_data, host, count * sizeof(T), cudaMemcpyHostToDevice, st));
    }
    void upload(const std::vector<T>& v, cudaStream_t st = 0) { upload(v.data(), v.size(), st); }
    T* data() const { return m_data; }
    operator T*() const { return m_data; }
    size_t size() const { return m_count; }
    size_t bytes() const { return m_count * sizeof(T); }

private:
    T* m_data = nullptr;
    size_t m_count = 0;
};

}  // namespace velvet
