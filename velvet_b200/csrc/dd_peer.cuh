// dd_peer.cuh -- NVLink peer-memory exchange of the domain-decomposed cloth (north_star mode 2).
//
// One process per GPU.  Every rank holds the full-size double-buffered `pred` arrays (float4 per particle, indexed by the
// global particle id), so a boundary particle travels as ONE 16-byte store into the peer's array at the same index:
// no pack / unpack, no host in the loop, no NCCL.  The peers' arrays are mapped with CUDA IPC (cudaIpcOpenMemHandle).
//
// Ordering between GPUs is a monotonically increasing epoch per rank:
//   push kernel : peer stores -> __threadfence_system -> last block bumps the local epoch and stores it (release, system
//                 scope) into flags[myRank] of every peer
//   wait kernel : spins (acquire, system scope) until flags[q] >= local epoch for every peer q, with a wall-clock timeout
// Every rank runs the same sequence of pushes, so the local epochs agree without any communication.  The kernels take all
// their state from device memory, so a whole frame (81+ kernels and 60+ exchanges) is ONE CUDA graph per rank.
#pragma once

#include <cuda_runtime.h>

namespace velvet {
namespace ddpeer {

constexpr int kMaxWorld = 16;

// Device-resident control block (one per solver).  `flags` is the IPC-exported array the peers write into.
struct Control {
    unsigned epoch;      // exchanges completed by this rank (bumped by the push / signal kernels)
    unsigned blocksDone; // last-block counter of the push kernels
    unsigned error;      // 1 = a wait timed out (peer missing); sticky until ddPeerReset
    unsigned pad;
    // ---- strip protocol (grid cloths, see below)
    unsigned seq;        // exchange launches completed by this rank: stable while a launch runs (its last block bumps it)
    unsigned sendsDone;  // blocks / boundary tiles of the running launch whose peer stores are performed
    unsigned exits;      // blocks that left the running launch
    unsigned pad2;
};

struct PeerTable {
    float4* pred[2][kMaxWorld];   // peers' predA / predB (own entry = local pointer)
    unsigned* flags[kMaxWorld];   // peers' flag arrays (own entry = local)
    int rank, world;
};

// other[ids[i]] of this rank -> the same index of peer sendPeer[i] (buffer `which` = 0 for predA, 1 for predB), then signal.
void launch_push_halo(cudaStream_t st, const PeerTable* table, Control* ctl, int which, const float4* src, const unsigned* sendIds,
                      const unsigned char* sendPeer, unsigned sendTotal);
// src[ownedIds[i]] -> every peer's buffer `which` at the same index (the per-substep all-gather), then signal.
void launch_push_owned(cudaStream_t st, const PeerTable* table, Control* ctl, int which, const float4* src, const unsigned* ownedIds,
                       unsigned ownedCount);
// no data: "I am done reading my buffers of the previous phase"
void launch_signal(cudaStream_t st, const PeerTable* table, Control* ctl);
void launch_wait(cudaStream_t st, const PeerTable* table, Control* ctl, const unsigned* localFlags, unsigned long long timeoutNs);

// ---- strip protocol: a grid cloth decomposed into strips of particle rows (contiguous index ranges).
// Exchange launch number L of a rank (every rank runs the same sequence) publishes the value L + 1 into flags[rank] of every
// peer once its peer stores are performed, and bumps Control::seq to L + 1 when its last block leaves.  A consumer reads
// seq at its start and waits until the flags of the ranks it depends on have reached it.  The Jacobi kernel takes part
// itself (fused_kernels.cu: iterate_grid_kernel processes its boundary tiles first, stores their outermost rows straight
// into the neighbours' arrays and publishes from the tile that finishes last), so an iteration needs no exchange launch.
constexpr int kStripFlagBase = 64;  // the strip protocol's flag words sit behind the epoch protocol's in the same exported array
struct StripArgs {
    PeerTable T;
    Control* ctl;
    const unsigned* localFlags;
    unsigned long long timeoutNs;
    int which;                 // output buffer of this launch (0 = predA, 1 = predB)
    int up, down;              // neighbour ranks (-1: none)
    unsigned tileRowBegin, tileRowEnd;  // owned tile rows of the cloth
    unsigned rowFirst, rowLast;         // first / last owned particle row
    int enabled;
    int gatherAll;             // this launch also stores EVERY owned result into every peer (the per-substep all-gather, fused)
};
// src[begin, begin + count) -> the same range of every peer's buffer `which`; publishes, bumps seq
void launch_strip_push_range(cudaStream_t st, const PeerTable* table, Control* ctl, int which, const float4* src, unsigned begin,
                             unsigned count);
// particle rows rowFirst -> rank `up`, rowLast -> rank `down` (side particles each); publishes, bumps seq
void launch_strip_push_rows(cudaStream_t st, const StripArgs& a, const float4* src, unsigned side);
void launch_strip_signal(cudaStream_t st, const PeerTable* table, Control* ctl);  // no data
void launch_strip_wait_all(cudaStream_t st, const PeerTable* table, Control* ctl, const unsigned* localFlags, unsigned long long timeoutNs);

#ifdef __CUDACC__
__device__ __forceinline__ void strip_st_release_sys(unsigned* p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned strip_ld_acquire_sys(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void strip_publish(const PeerTable& T, unsigned value)
{
    for (int q = 0; q < T.world; q++)
        if (q != T.rank) strip_st_release_sys(T.flags[q] + kStripFlagBase + T.rank, value);
}
// one thread: spin until rank q has published `value` (wall-clock timeout -> sticky error, like the wait kernel)
__device__ __forceinline__ void strip_wait_for(const unsigned* localFlags, int q, unsigned value, Control* ctl, unsigned long long timeoutNs)
{
    if (q < 0 || *(volatile unsigned*)&ctl->error) return;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    unsigned spins = 0;
    while ((int)(strip_ld_acquire_sys(localFlags + kStripFlagBase + q) - value) < 0) {
        if ((++spins & 1023u) == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t - t0 > timeoutNs) {
                *(volatile unsigned*)&ctl->error = 1u;
                break;
            }
        }
    }
}
#endif

}  // namespace ddpeer
}  // namespace velvet
