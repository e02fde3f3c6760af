// peer.cuh — device-side half of the leading-axis sharding (shard.cu; new — the reference is single
// device, Tensor/Tensor/Cuda/CudaUtils.fs:42-46).
//
// A sharded launch stores every output element into the local result buffer AND into the same offset of every
// peer's copy of it (plain stores to peer-mapped memory: NVLink / NVSwitch P2P, or CUDA-IPC mappings between
// processes). The last CTA of the grid to finish then publishes this rank's epoch into every rank's flag array
// (system-scope release stores). The WAIT half of the barrier is not a kernel: shard.cu enqueues stream memory
// operations (cuStreamWaitValue32 on the local flags) behind the launch, so the stream owns the complete replicated
// result once they have passed. One launch per device, no NCCL launch on the path, and no kernel ever spins — which
// keeps the scheme free of the deadlocks a spinning kernel invites (ranks sharing a device, lazily loaded kernels
// that need the context quiet, exhausted SM resources).
#pragma once

#include <stdint.h>

namespace dn {

constexpr int kMaxShardRanks = 8;

struct PeerSync {
    int32_t npeers;                        // other ranks that receive a copy of every output (0: plain launch)
    int32_t nflags;                        // ranks taking part in the exit barrier (0: no barrier)
    uint32_t epoch;                        // value this collective publishes
    int32_t pad;
    int64_t delta[kMaxShardRanks - 1];     // byte distance from a local window address to peer k's mapping of it
    uint32_t *flag_peer[kMaxShardRanks];   // &done[my rank] in every rank's window (own window included)
    uint32_t *flag_local;                  // done[0..nflags) of the local window
    uint32_t *counter;                     // CTAs of this launch that have finished (local, zero between launches)
};

__device__ __forceinline__ void st_release_sys_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Publish this rank's epoch into every rank's flag array; one thread.
__device__ __forceinline__ void peer_signal(const PeerSync &ps) {
    __threadfence_system();
    for (int k = 0; k < ps.nflags; ++k) st_release_sys_u32(ps.flag_peer[k], ps.epoch);
}

// Called by EVERY thread of EVERY CTA at the end of a kernel that may be a sharded launch.
__device__ __forceinline__ void peer_exit(const PeerSync &ps) {
    if (ps.nflags == 0) return;  // uniform
    __threadfence_system();      // this thread's peer stores
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) {
        const unsigned total = gridDim.x * gridDim.y * gridDim.z;
        if (atomicAdd(ps.counter, 1u) == total - 1) {
            *ps.counter = 0;  // the next launch of this rank is stream-ordered after this kernel
            peer_signal(ps);
        }
    }
}

}  // namespace dn
