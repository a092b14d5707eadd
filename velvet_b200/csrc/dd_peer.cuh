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

}  // namespace ddpeer
}  // namespace velvet
