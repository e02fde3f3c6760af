// peer.cuh — device-side half of the leading-axis sharding (shard.cu; new — the reference is single
// device, Tensor/Tensor/Cuda/CudaUtils.fs:42-46).
//
// A sharded launch stores every output element into the local result buffer AND into the same offset of every
// peer's copy of it (plain stores to peer-mapped memory: NVLink / NVSwitch P2P, or CUDA-IPC mappings between
// processes). The last CTA of the grid to finish then publishes this rank's epoch into every rank's flag array
// (one system-scope fence, then relaxed system-scope stores). The WAIT half of the barrier comes in two forms (shard.cu picks):
//   * one rank per process (torchrun): the signalling thread itself spins on the local flags, so the kernel retires
//     with the complete replicated result — one launch per device and nothing else on the path;
//   * one process driving several ranks: stream memory operations (cuStreamWaitValue32 on the local flags) enqueued
//     behind the launch. No kernel spins there, which keeps that model free of the deadlocks a spinning kernel
//     invites when the host thread it waits for is the one that launched it (ranks sharing a device, lazily loaded
//     kernels that need the context quiet).
// No NCCL launch in either form.
#pragma once

#include <stdint.h>

namespace dn {

constexpr int kMaxShardRanks = 8;

struct PeerSync {
    int32_t npeers;                        // other ranks that receive a copy of every output (0: plain launch)
    int32_t nflags;                        // ranks taking part in the exit barrier (0: no barrier)
    uint32_t epoch;                        // value this collective publishes
    int32_t wait_in_kernel;                // 1: the signalling thread also waits for every rank's flag (see shard.cu)
    int64_t delta[kMaxShardRanks - 1];     // byte distance from a local window address to peer k's mapping of it
    uint32_t *flag_peer[kMaxShardRanks];   // &done[my rank] in every rank's window (own window included)
    uint32_t *flag_local;                  // done[0..nflags) of the local window
    uint32_t *counter;                     // CTAs of this launch that have finished (local, zero between launches)
    uint32_t *error;                       // sticky: an in-kernel wait gave up (a rank never arrived)
};

__device__ __forceinline__ void st_relaxed_sys_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed_sys_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Publish this rank's epoch into every rank's flag array; one thread. With wait_in_kernel the same thread then
// waits until every rank has published into the LOCAL array: the kernel retires with the replicated result complete
// (a few microseconds sooner than stream memory operations notice the flags). Only used when the process drives a
// single rank, where nothing the host still has to do can be what the spinning thread waits for.
__device__ __forceinline__ void peer_signal(const PeerSync &ps) {
    // release pattern: ONE system-scope fence, then relaxed flag stores (a st.release per flag would pay the fence
    // once per rank); acquire pattern on the other side: relaxed polls, one fence when all flags are in
    __threadfence_system();
    for (int k = 0; k < ps.nflags; ++k) st_relaxed_sys_u32(ps.flag_peer[k], ps.epoch);
    if (!ps.wait_in_kernel) return;
    const uint64_t t0 = global_timer_ns();
    for (int k = 0; k < ps.nflags; ++k) {
        while ((int32_t)(ld_relaxed_sys_u32(ps.flag_local + k) - ps.epoch) < 0) {
            if (global_timer_ns() - t0 > 20000000000ull) {  // 20 s: a rank never arrived; do not hang the GPU
                *ps.error = 1;
                return;
            }
        }
    }
    __threadfence_system();
}

// Called by EVERY thread of EVERY CTA at the end of a kernel that may be a sharded launch. One system-scope fence
// per CTA, issued by the thread that then counts the CTA in: the barrier orders the other threads' (peer) stores
// before it and the fence is cumulative (the scheme of cooperative-groups grid synchronisation). A fence in every
// thread costs ~25 us on a 168 us kernel — each of ~600 k MEMBAR.SYS waits for the CTA's outstanding writes.
__device__ __forceinline__ void peer_exit(const PeerSync &ps) {
    if (ps.nflags == 0) return;  // uniform
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) {
        __threadfence_system();
        const unsigned total = gridDim.x * gridDim.y * gridDim.z;
        if (atomicAdd(ps.counter, 1u) == total - 1) {
            *ps.counter = 0;  // the next launch of this rank is stream-ordered after this kernel
            peer_signal(ps);
        }
    }
}

}  // namespace dn
