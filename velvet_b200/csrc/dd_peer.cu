// dd_peer.cu -- kernels of the NVLink peer-memory exchange (see dd_peer.cuh for the protocol).
#include "dd_peer.cuh"

namespace velvet {
namespace ddpeer {
namespace {

constexpr int PB = 256;

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Called by every thread of every block after its peer stores.  The last block to arrive publishes the new epoch.
__device__ __forceinline__ void signal_when_grid_done(const PeerTable& T, Control* ctl)
{
    __threadfence_system();  // this thread's peer stores are performed before the block reports in
    __syncthreads();
    if (threadIdx.x != 0) return;
    const unsigned prev = atomicAdd(&ctl->blocksDone, 1u);
    if (prev != gridDim.x - 1) return;
    __threadfence_system();  // order the other blocks' stores (observed through the counter) before the flags
    ctl->blocksDone = 0;
    const unsigned e = ctl->epoch + 1;
    ctl->epoch = e;
    for (int q = 0; q < T.world; q++)
        if (q != T.rank) st_release_sys(T.flags[q] + T.rank, e);
}

__global__ void __launch_bounds__(PB) push_halo_kernel(const PeerTable T, Control* ctl, const int which, const float4* __restrict__ src,
                                                       const unsigned* __restrict__ sendIds, const unsigned char* __restrict__ sendPeer,
                                                       const unsigned sendTotal)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < sendTotal) {
        const unsigned id = sendIds[i];
        T.pred[which][sendPeer[i]][id] = src[id];
    }
    signal_when_grid_done(T, ctl);
}

__global__ void __launch_bounds__(PB) push_owned_kernel(const PeerTable T, Control* ctl, const int which, const float4* __restrict__ src,
                                                        const unsigned* __restrict__ ownedIds, const unsigned ownedCount)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ownedCount) {
        const unsigned id = ownedIds[i];
        const float4 v = src[id];
        for (int q = 0; q < T.world; q++)
            if (q != T.rank) T.pred[which][q][id] = v;
    }
    signal_when_grid_done(T, ctl);
}

__global__ void signal_kernel(const PeerTable T, Control* ctl) { signal_when_grid_done(T, ctl); }

__global__ void wait_kernel(const PeerTable T, Control* ctl, const unsigned* localFlags, const unsigned long long timeoutNs)
{
    const int q = threadIdx.x;
    if (q >= T.world || q == T.rank) return;
    if (*(volatile unsigned*)&ctl->error) return;  // a peer already went missing: do not stack timeouts
    const unsigned e = ctl->epoch;
    const unsigned long long t0 = global_timer_ns();
    unsigned spins = 0;
    while ((int)(ld_acquire_sys(localFlags + q) - e) < 0) {
        if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > timeoutNs) {
            *(volatile unsigned*)&ctl->error = 1u;
            break;
        }
    }
}

inline unsigned grid_of(unsigned n) { return n ? (n + PB - 1) / PB : 1u; }

// ---- strip protocol
// every thread of every block, after its peer stores: the last block publishes seq + 1 and completes the launch
__device__ __forceinline__ void strip_finish_push(const PeerTable& T, Control* ctl)
{
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x != 0) return;
    const unsigned prev = atomicAdd(&ctl->sendsDone, 1u);
    if (prev != gridDim.x - 1) return;
    __threadfence_system();
    ctl->sendsDone = 0;
    const unsigned v = ctl->seq + 1;
    strip_publish(T, v);
    ctl->seq = v;  // single-purpose kernel: nothing in this launch reads seq after its stores
}

__global__ void __launch_bounds__(PB) strip_push_range_kernel(const PeerTable T, Control* ctl, const int which, const float4* __restrict__ src,
                                                              const unsigned begin, const unsigned count)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) {
        const float4 v = src[begin + i];
        for (int q = 0; q < T.world; q++)
            if (q != T.rank) T.pred[which][q][begin + i] = v;
    }
    strip_finish_push(T, ctl);
}

__global__ void __launch_bounds__(PB) strip_push_rows_kernel(const StripArgs a, const float4* __restrict__ src, const unsigned side)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < side) {
        if (a.up >= 0) a.T.pred[a.which][a.up][(size_t)a.rowFirst * side + i] = src[(size_t)a.rowFirst * side + i];
        if (a.down >= 0) a.T.pred[a.which][a.down][(size_t)a.rowLast * side + i] = src[(size_t)a.rowLast * side + i];
    }
    strip_finish_push(a.T, a.ctl);
}

__global__ void strip_signal_kernel(const PeerTable T, Control* ctl) { strip_finish_push(T, ctl); }

__global__ void strip_wait_all_kernel(const PeerTable T, Control* ctl, const unsigned* localFlags, const unsigned long long timeoutNs)
{
    const int q = threadIdx.x;
    if (q >= T.world || q == T.rank) return;
    strip_wait_for(localFlags, q, ctl->seq, ctl, timeoutNs);
}

}  // namespace

void launch_push_halo(cudaStream_t st, const PeerTable* table, Control* ctl, int which, const float4* src, const unsigned* sendIds,
                      const unsigned char* sendPeer, unsigned sendTotal)
{
    push_halo_kernel<<<grid_of(sendTotal), PB, 0, st>>>(*table, ctl, which, src, sendIds, sendPeer, sendTotal);
}

void launch_push_owned(cudaStream_t st, const PeerTable* table, Control* ctl, int which, const float4* src, const unsigned* ownedIds,
                       unsigned ownedCount)
{
    push_owned_kernel<<<grid_of(ownedCount), PB, 0, st>>>(*table, ctl, which, src, ownedIds, ownedCount);
}

void launch_signal(cudaStream_t st, const PeerTable* table, Control* ctl) { signal_kernel<<<1, 32, 0, st>>>(*table, ctl); }

void launch_wait(cudaStream_t st, const PeerTable* table, Control* ctl, const unsigned* localFlags, unsigned long long timeoutNs)
{
    wait_kernel<<<1, 32, 0, st>>>(*table, ctl, localFlags, timeoutNs);
}

void launch_strip_push_range(cudaStream_t st, const PeerTable* table, Control* ctl, int which, const float4* src, unsigned begin,
                             unsigned count)
{
    strip_push_range_kernel<<<grid_of(count), PB, 0, st>>>(*table, ctl, which, src, begin, count);
}

void launch_strip_push_rows(cudaStream_t st, const StripArgs& a, const float4* src, unsigned side)
{
    strip_push_rows_kernel<<<grid_of(side), PB, 0, st>>>(a, src, side);
}

void launch_strip_signal(cudaStream_t st, const PeerTable* table, Control* ctl) { strip_signal_kernel<<<1, 32, 0, st>>>(*table, ctl); }

void launch_strip_wait_all(cudaStream_t st, const PeerTable* table, Control* ctl, const unsigned* localFlags, unsigned long long timeoutNs)
{
    strip_wait_all_kernel<<<1, 32, 0, st>>>(*table, ctl, localFlags, timeoutNs);
}

}  // namespace ddpeer
}  // namespace velvet
