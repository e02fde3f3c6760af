// shard.cu — leading-axis sharding over the GPUs of one box behind the C ABI (include/dn_tensor.h, dn_shard_*).
//
// New: the reference runs on one device (Tensor/Tensor/Cuda/CudaUtils.fs:42-46); SURVEY.md §8e defines the
// partitioning this file implements. Design (see also peer.cuh):
//   * every rank owns a WINDOW (one cudaMalloc): [flags | partial-result scratch x2 | symmetric heap]. Inside one
//     process the windows are reached through peer access (NVLink / NVSwitch), between processes through CUDA IPC
//     mappings; either way a peer's copy of a local window address is `address + delta[peer]`.
//   * a sharded reduction is the ordinary reduction kernel launched with a PeerSync block: it stores every output
//     into all ranks' result buffers and its last CTA runs the exit barrier. One launch per rank, no NCCL.
//   * reductions over the sharded axis store per-rank partials into every rank's scratch (same mechanism) and
//     fold them locally in rank order with the ordinary operators, so every rank computes identical bits.
#include "reduce.cuh"

#include <cuda.h>

#include <mutex>
#include <new>

namespace dn {

namespace {

thread_local PeerSync t_pending = {};
thread_local bool t_has_pending = false;

constexpr int64_t kFlagBytes = 4096;             // done[8] | counter | counts[2][8]
constexpr int64_t kCounterOff = 256, kErrorOff = 320, kCountsOff = 512;
constexpr int64_t kScratchHalf = 8ll << 20;      // partials of reductions over the sharded axis (two halves)
constexpr int64_t kHeapOff = kFlagBytes + 2 * kScratchHalf;

// What still has to happen on a rank's stream after its signalling kernel: the wait half of the barrier and,
// for some collectives, a step that needs every rank's contribution.
struct PostJob {
    int kind = 0;  // 0: none, 1: fold the per-rank partials of a sharded-axis reduction, 2: read the counts back
    int red_kind = 0, op = 0, in_dt = 0, world = 0;
    dn_tensor t = {};
    char *half = nullptr;
    int64_t slot_bytes = 0, n = 0;
    int64_t *counts_dev = nullptr;
};

struct ShardRank {
    bool local = false;
    int device = -1;
    char *window = nullptr;     // this process' mapping of the rank's window
    bool ipc_mapped = false;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    int64_t heap_used = 0;
    uint32_t epoch = 0;         // collectives issued by this (local) rank
    uint32_t scratch_uses = 0;
    int64_t last_target = -1;   // window offset of the previous collective's target (outside brackets)
    int64_t prev_targets[8], cur_targets[8];  // targets of the previous / the current bracket
    int nprev = 0, ncur = 0;
    int64_t *host_counts = nullptr;  // pinned, kMaxShardRanks entries: read-back of dn_shard_count_true
    bool has_pending = false;        // inside dn_shard_group_start / _end: the deferred wait + post job
    PeerSync pending_sync = {};
    PostJob pending_job;
};

struct ShardGroup {
    int world = 0;
    int64_t heap_bytes = 0, window_bytes = 0;
    bool connected = false;
    int nlocal = 0;
    bool grouping = false;  // between dn_shard_group_start and dn_shard_group_end
    ShardRank r[kMaxShardRanks];
};

// Binds the calling thread to a rank's device and stream for the duration of one call.
struct RankScope {
    int prev_dev = 0;
    cudaStream_t prev_stream = nullptr;
    bool switched = false;
    explicit RankScope(const ShardRank &r) {
        cudaGetDevice(&prev_dev);
        if (prev_dev != r.device) {
            cudaSetDevice(r.device);
            switched = true;
        }
        prev_stream = current_stream();
        set_thread_stream(r.stream);
    }
    ~RankScope() {
        set_thread_stream(prev_stream);
        if (switched) cudaSetDevice(prev_dev);
    }
};

dn_status get_rank(void *group, int32_t rank, ShardGroup *&g, ShardRank *&r, const char *what) {
    g = static_cast<ShardGroup *>(group);
    if (!g || rank < 0 || rank >= g->world) return set_error(DN_ERR_INVALID_ARG, "%s: bad group or rank", what);
    r = &g->r[rank];
    if (!r->local) return set_error(DN_ERR_INVALID_ARG, "%s: rank %d is not driven by this process", what, rank);
    if (!g->connected) return set_error(DN_ERR_INVALID_ARG, "%s: dn_shard_group_connect has not been called", what);
    return DN_OK;
}

// The PeerSync block of the next collective of `rank`: `with_data` = outputs are replicated into the peers.
PeerSync make_sync(ShardGroup &g, ShardRank &me, int rank, bool with_data) {
    PeerSync ps = {};
    ps.nflags = g.world;
    ps.epoch = ++me.epoch;
    int n = 0;
    for (int k = 0; k < g.world; ++k) {
        ps.flag_peer[k] = reinterpret_cast<uint32_t *>(g.r[k].window) + rank;
        if (k != rank && with_data) ps.delta[n++] = g.r[k].window - me.window;
    }
    ps.npeers = n;
    ps.flag_local = reinterpret_cast<uint32_t *>(me.window);
    ps.counter = reinterpret_cast<uint32_t *>(me.window + kCounterOff);
    ps.error = reinterpret_cast<uint32_t *>(me.window + kErrorOff);
    // one rank per process: the wait half runs inside the signalling kernel (peer.cuh); several local ranks: as
    // stream memory operations, deferred to dn_shard_group_end inside a bracket
    ps.wait_in_kernel = (g.nlocal == 1 && !g.grouping) ? 1 : 0;
    return ps;
}

__global__ void peer_barrier_kernel(const __grid_constant__ PeerSync ps) {
    if (threadIdx.x == 0) peer_signal(ps);
}

// The wait half on its own, as a one-thread kernel (a process that drives a single rank closing a bracket).
__global__ void peer_wait_kernel(const __grid_constant__ PeerSync ps) {
    if (threadIdx.x != 0) return;
    const uint64_t t0 = global_timer_ns();
    for (int k = 0; k < ps.nflags; ++k) {
        while ((int32_t)(ld_relaxed_sys_u32(ps.flag_local + k) - ps.epoch) < 0) {
            if (global_timer_ns() - t0 > 20000000000ull) {
                *ps.error = 1;
                return;
            }
        }
    }
    __threadfence_system();
}

// The wait half of the barrier: the stream does not proceed until every rank's flag in the LOCAL window has
// reached `epoch` (cyclic >=). Stream memory operations (no SM is occupied while waiting) when the process drives
// several ranks; a one-thread kernel when it drives one (`spin`).
dn_status enqueue_wait(const PeerSync &ps, bool spin = false) {
    if (ps.wait_in_kernel) return DN_OK;  // the signalling kernel has already waited
    if (spin) {
        DN_LAUNCH(peer_wait_kernel, 1, 32, 0, ps);
        return launch_status("shard wait kernel");
    }
    CUstreamBatchMemOpParams ops[kMaxShardRanks];
    memset(ops, 0, sizeof ops);
    for (int k = 0; k < ps.nflags; ++k) {
        ops[k].waitValue.operation = CU_STREAM_MEM_OP_WAIT_VALUE_32;
        ops[k].waitValue.address = (CUdeviceptr)(uintptr_t)(ps.flag_local + k);
        ops[k].waitValue.value = ps.epoch;
        ops[k].waitValue.flags = CU_STREAM_WAIT_VALUE_GEQ;
    }
    // resolved through the runtime (no link-time dependency on libcuda: the library must load on a box without a driver)
    typedef CUresult (*BatchMemOpFn)(CUstream, unsigned int, CUstreamBatchMemOpParams *, unsigned int);
    static BatchMemOpFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamBatchMemOp", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
        return reinterpret_cast<BatchMemOpFn>(p);
    }();
    if (!fn) return set_error(DN_ERR_CUDA, "cuStreamBatchMemOp is not available from the driver");
    CUresult r = fn((CUstream)current_stream(), (unsigned)ps.nflags, ops, 0);
    if (r != CUDA_SUCCESS) return set_error(DN_ERR_CUDA, "cuStreamBatchMemOp (shard barrier wait) failed with code %d", (int)r);
    return DN_OK;
}

// Copies [begin, begin+nbytes) of the local window to the same place of every peer, then the exit barrier.
__global__ void __launch_bounds__(256) peer_push_kernel(const __grid_constant__ PeerSync ps, char *base, int64_t nbytes) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (int64_t)gridDim.x * blockDim.x;
    const uintptr_t a = reinterpret_cast<uintptr_t>(base);
    int64_t head = (int64_t)((16 - (a & 15)) & 15);
    if (head > nbytes) head = nbytes;
    const int64_t nvec = (nbytes - head) / 16, tail = head + nvec * 16;
    for (int k = 0; k < ps.npeers; ++k) {
        char *dst = base + ps.delta[k];
        for (int64_t i = tid; i < nvec; i += nthreads)
            reinterpret_cast<uint4 *>(dst + head)[i] = reinterpret_cast<const uint4 *>(base + head)[i];
        for (int64_t i = tid; i < head; i += nthreads) dst[i] = base[i];
        for (int64_t i = tail + tid; i < nbytes; i += nthreads) dst[i] = base[i];
    }
    peer_exit(ps);
}

// The SIGNAL half on its own (nothing to store, or a pure barrier).
dn_status launch_signal(const PeerSync &ps) {
    DN_LAUNCH(peer_barrier_kernel, 1, 32, 0, ps);
    return launch_status("shard signal kernel");
}

dn_status launch_push(const PeerSync &ps, char *base, int64_t nbytes) {
    if (nbytes <= 0 || ps.npeers == 0) return launch_signal(ps);
    int64_t ctas = (nbytes / 16 + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 4;
    if (ctas > cap) ctas = cap;
    if (ctas < 1) ctas = 1;
    DN_LAUNCH(peer_push_kernel, (unsigned)ctas, 256, 0, ps, base, nbytes);
    return launch_status("shard push kernel");
}

// Runs `body` (ordinary operator entry points) with `ps` pending; if no kernel consumed it (empty slab, nothing to
// store), the barrier is issued on its own so that the collective stays matched on every rank.
template <class F>
dn_status with_pending(const PeerSync &ps, F body) {
    t_pending = ps;
    t_has_pending = true;
    dn_status st = body();
    const bool unconsumed = t_has_pending;
    t_has_pending = false;
    if (st != DN_OK) return st;
    return unconsumed ? launch_signal(ps) : DN_OK;
}

dn_status run_post(ShardRank &me, const PostJob &job);  // below, next to the fold kernels

// The second half of every collective, on the rank's stream: hold the stream until every rank has signalled, then
// run the step that needs all contributions. Inside a dn_shard_group_start / _end bracket (ONE thread driving
// several ranks) this half is deferred to dn_shard_group_end, i.e. until every local rank has issued its signalling
// kernel: a stream that is already waiting for a peer whose kernel the same thread has yet to launch would turn any
// blocking call on the way there (a lazily loaded kernel, a pool that has to grow) into a deadlock — the reason
// ncclGroupStart / ncclGroupEnd exist.
dn_status complete(ShardGroup &g, ShardRank &me, const PeerSync &ps, const PostJob &job) {
    if (g.grouping) {
        // several collectives per rank may share a bracket: flags only grow, so waiting for the LAST epoch covers the
        // earlier ones. A collective with a post step needs its wait first and therefore has to be the last one.
        if (me.has_pending && me.pending_job.kind != 0)
            return set_error(DN_ERR_INVALID_ARG, "inside a bracket a reduction over the sharded axis / count exchange "
                                                 "must be the rank's last collective");
        me.has_pending = true;
        me.pending_sync = ps;
        me.pending_job = job;
        return DN_OK;
    }
    dn_status st = enqueue_wait(ps);
    return st != DN_OK ? st : run_post(me, job);
}

bool in_heap(const ShardGroup &g, const ShardRank &r, const char *p, int64_t nbytes) {
    return p >= r.window + kHeapOff && p + nbytes <= r.window + g.window_bytes;
}

// A target may be reused only after another collective in between (peers may still be consuming it): when the
// same target comes twice in a row, an entry barrier restores the order. Every rank sees the same window offsets,
// so every rank takes the same decision.
dn_status guard_target(ShardGroup &g, ShardRank &me, int rank, const char *target) {
    const int64_t off = target - me.window;
    if (g.grouping) {
        // a peer may be one bracket ahead: this bracket's targets must differ from each other and from the previous
        // bracket's (they are free again from the bracket after that, or after a dn_shard_barrier)
        for (int i = 0; i < me.ncur; ++i)
            if (me.cur_targets[i] == off)
                return set_error(DN_ERR_INVALID_ARG, "the same target twice inside one bracket");
        for (int i = 0; i < me.nprev; ++i)
            if (me.prev_targets[i] == off)
                return set_error(DN_ERR_INVALID_ARG, "a target of the previous bracket is reused: alternate two sets of "
                                                     "targets or issue dn_shard_barrier on every rank in between");
        if (me.ncur < 8) me.cur_targets[me.ncur++] = off;
        return DN_OK;
    }
    bool recent = off == me.last_target;
    for (int i = 0; i < me.nprev; ++i) recent = recent || me.prev_targets[i] == off;
    if (recent) {
        const PeerSync ps = make_sync(g, me, rank, false);
        dn_status st = launch_signal(ps);
        if (st == DN_OK) st = enqueue_wait(ps);
        if (st != DN_OK) return st;
    }
    me.last_target = off;
    me.nprev = 0;
    return DN_OK;
}

// The rows [row_begin, row_begin + rows) of t_full as a view.
dn_status slab_view(dn_tensor &v, const dn_tensor *t_full, int64_t row_begin, int64_t rows, const char *what) {
    if (!tensor_valid(t_full) || t_full->ndims < 1)
        return set_error(DN_ERR_INVALID_ARG, "%s: the full result must have at least one dimension", what);
    if (row_begin < 0 || rows < 0 || row_begin + rows > t_full->shape[0])
        return set_error(DN_ERR_SHAPE_MISMATCH, "%s: rows [%lld, %lld) outside the full result's %lld rows", what,
                         (long long)row_begin, (long long)(row_begin + rows), (long long)t_full->shape[0]);
    v = *t_full;
    v.offset += row_begin * t_full->stride[0];
    v.shape[0] = rows;
    return DN_OK;
}

int64_t span_bytes(const dn_tensor *t) {  // extent of a non-negative-stride view from its first element
    int64_t last = 0;
    for (int d = 0; d < t->ndims; ++d) {
        if (t->shape[d] == 0) return 0;
        last += (t->shape[d] - 1) * (t->stride[d] < 0 ? -t->stride[d] : t->stride[d]);
    }
    return (last + 1) * dtype_size(t->dtype);
}

dn_status check_full_target(const ShardGroup &g, const ShardRank &me, const dn_tensor *t_full, const char *what) {
    if (!tensor_valid(t_full)) return set_error(DN_ERR_INVALID_ARG, "%s: bad target", what);
    for (int d = 0; d < t_full->ndims; ++d)
        if (t_full->stride[d] < 0)
            return set_error(DN_ERR_INVALID_ARG, "%s: the full result must not be a reversed view", what);
    if (!in_heap(g, me, data_ptr(t_full), span_bytes(t_full)))
        return set_error(DN_ERR_INVALID_ARG, "%s: the full result must live in the rank's symmetric heap "
                                             "(dn_shard_heap_alloc)", what);
    return DN_OK;
}

bool is_c_contiguous(const dn_tensor *t) {
    int64_t expect = 1;
    for (int d = t->ndims - 1; d >= 0; --d) {
        if (t->shape[d] != 1 && t->stride[d] != expect) return false;
        expect *= t->shape[d];
    }
    return true;
}

dn_tensor make_contig(void *base, int64_t byte_off, int dtype, int ndims, const int64_t *shape) {
    dn_tensor t = {};
    t.base = static_cast<char *>(base) + byte_off;
    t.dtype = dtype;
    t.ndims = ndims;
    int64_t st = 1;
    for (int d = ndims - 1; d >= 0; --d) {
        t.shape[d] = shape[d];
        t.stride[d] = st;
        st *= shape[d];
    }
    return t;
}

}  // namespace

bool peer_take(PeerSync &out) {
    if (!t_has_pending) return false;
    out = t_pending;
    t_has_pending = false;
    return true;
}

}  // namespace dn

using namespace dn;

extern "C" {

dn_status dn_shard_slab(int64_t nrows, int32_t rank, int32_t world, int64_t *begin, int64_t *count) {
    if (!begin || !count || world < 1 || rank < 0 || rank >= world || nrows < 0)
        return set_error(DN_ERR_INVALID_ARG, "dn_shard_slab: bad argument");
    const int64_t base = nrows / world, rem = nrows % world;
    *begin = rank * base + (rank < rem ? rank : rem);
    *count = base + (rank < rem ? 1 : 0);
    return DN_OK;
}

dn_status dn_shard_group_create(int32_t world, int32_t nlocal, const int32_t *local_ranks,
                                const int32_t *local_devices, int64_t heap_bytes, void **group) {
    if (!group || world < 1 || world > kMaxShardRanks || nlocal < 1 || nlocal > world || !local_ranks ||
        !local_devices || heap_bytes < 0)
        return set_error(DN_ERR_INVALID_ARG, "dn_shard_group_create: bad argument (1 <= world <= %d)", kMaxShardRanks);
    ShardGroup *g = new (std::nothrow) ShardGroup();
    if (!g) return set_error(DN_ERR_OUT_OF_MEMORY, "dn_shard_group_create: out of host memory");
    g->world = world;
    g->nlocal = nlocal;
    g->heap_bytes = (heap_bytes + 255) / 256 * 256;
    g->window_bytes = kHeapOff + g->heap_bytes;
    int prev = 0;
    cudaGetDevice(&prev);
    dn_status st = DN_OK;
    for (int i = 0; i < nlocal && st == DN_OK; ++i) {
        const int rk = local_ranks[i];
        if (rk < 0 || rk >= world || g->r[rk].local) {
            st = set_error(DN_ERR_INVALID_ARG, "dn_shard_group_create: bad or duplicate local rank %d", rk);
            break;
        }
        ShardRank &r = g->r[rk];
        st = dn_init(local_devices[i]);
        if (st != DN_OK) break;
        r.local = true;
        r.device = local_devices[i];
        cudaError_t e = cudaMalloc((void **)&r.window, (size_t)g->window_bytes);
        if (e == cudaSuccess) e = cudaMemset(r.window, 0, (size_t)kFlagBytes);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&r.own_stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaMallocHost((void **)&r.host_counts, sizeof(int64_t) * kMaxShardRanks);
        if (e != cudaSuccess) {
            st = cuda_error(e, "dn_shard_group_create");
            break;
        }
        r.stream = r.own_stream;
    }
    cudaSetDevice(prev);
    if (st != DN_OK) {
        dn_shard_group_destroy(g);
        return st;
    }
    *group = g;
    return DN_OK;
}

dn_status dn_shard_group_handle(void *group, int32_t rank, void *handle) {
    ShardGroup *g = static_cast<ShardGroup *>(group);
    if (!g || !handle || rank < 0 || rank >= g->world || !g->r[rank].local)
        return set_error(DN_ERR_INVALID_ARG, "dn_shard_group_handle: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == DN_SHARD_HANDLE_BYTES, "IPC handle size");
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(g->r[rank].device);
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, g->r[rank].window);
    cudaSetDevice(prev);
    if (e != cudaSuccess) return cuda_error(e, "cudaIpcGetMemHandle");
    memcpy(handle, &h, sizeof h);
    return DN_OK;
}

dn_status dn_shard_group_connect(void *group, const void *handles) {
    ShardGroup *g = static_cast<ShardGroup *>(group);
    if (!g) return set_error(DN_ERR_INVALID_ARG, "dn_shard_group_connect: null group");
    if (g->connected) return DN_OK;
    int prev = 0;
    cudaGetDevice(&prev);
    int first_local = -1;
    for (int k = 0; k < g->world; ++k)
        if (g->r[k].local && first_local < 0) first_local = k;
    dn_status st = DN_OK;
    // windows of ranks driven by other processes: CUDA IPC mappings (opened on the first local rank's device)
    for (int k = 0; k < g->world && st == DN_OK; ++k) {
        ShardRank &r = g->r[k];
        if (r.local) continue;
        if (!handles) {
            st = set_error(DN_ERR_INVALID_ARG, "dn_shard_group_connect: rank %d is not local and no handles were given", k);
            break;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char *>(handles) + (size_t)k * DN_SHARD_HANDLE_BYTES, sizeof h);
        cudaSetDevice(g->r[first_local].device);
        cudaError_t e = cudaIpcOpenMemHandle((void **)&r.window, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            st = cuda_error(e, "cudaIpcOpenMemHandle (peer window)");
            break;
        }
        r.ipc_mapped = true;
    }
    // local ranks on different devices: direct peer access
    for (int a = 0; a < g->world && st == DN_OK; ++a) {
        if (!g->r[a].local) continue;
        for (int b = 0; b < g->world && st == DN_OK; ++b) {
            if (!g->r[b].local || g->r[a].device == g->r[b].device) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, g->r[a].device, g->r[b].device);
            if (!can) {
                st = set_error(DN_ERR_UNSUPPORTED, "devices %d and %d cannot access each other's memory", g->r[a].device,
                               g->r[b].device);
                break;
            }
            cudaSetDevice(g->r[a].device);
            cudaError_t e = cudaDeviceEnablePeerAccess(g->r[b].device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) {
                cudaGetLastError();
            } else if (e != cudaSuccess) {
                st = cuda_error(e, "cudaDeviceEnablePeerAccess");
            }
        }
    }
    cudaSetDevice(prev);
    if (st == DN_OK) g->connected = true;
    return st;
}

dn_status dn_shard_group_destroy(void *group) {
    ShardGroup *g = static_cast<ShardGroup *>(group);
    if (!g) return DN_OK;
    int prev = 0;
    cudaGetDevice(&prev);
    for (int k = 0; k < g->world; ++k) {
        ShardRank &r = g->r[k];
        if (r.local) {
            cudaSetDevice(r.device);
            if (r.own_stream) {
                cudaStreamSynchronize(r.own_stream);
                cudaStreamDestroy(r.own_stream);
            }
            if (r.window) cudaFree(r.window);
            if (r.host_counts) cudaFreeHost(r.host_counts);
        } else if (r.ipc_mapped && r.window) {
            cudaIpcCloseMemHandle(r.window);
        }
    }
    cudaSetDevice(prev);
    delete g;
    return DN_OK;
}

dn_status dn_shard_set_stream(void *group, int32_t rank, void *stream) {
    ShardGroup *g = static_cast<ShardGroup *>(group);
    if (!g || rank < 0 || rank >= g->world || !g->r[rank].local)
        return set_error(DN_ERR_INVALID_ARG, "dn_shard_set_stream: bad argument");
    g->r[rank].stream = stream ? static_cast<cudaStream_t>(stream) : g->r[rank].own_stream;
    return DN_OK;
}

dn_status dn_shard_sync(void *group, int32_t rank) {
    ShardGroup *g;
    ShardRank *me;
    dn_status st = get_rank(group, rank, g, me, "dn_shard_sync");
    if (st != DN_OK) return st;
    RankScope scope(*me);
    uint32_t err = 0;
    DN_CUDA_TRY(cudaMemcpyAsync(&err, me->window + kErrorOff, sizeof err, cudaMemcpyDeviceToHost, me->stream));
    DN_CUDA_TRY(cudaStreamSynchronize(me->stream));
    if (err) {
        cudaMemsetAsync(me->window + kErrorOff, 0, sizeof err, me->stream);
        return set_error(DN_ERR_CUDA, "shard barrier timed out on rank %d: a rank did not issue the matching collective", rank);
    }
    return DN_OK;
}

dn_status dn_shard_heap_alloc(void *group, int32_t rank, int64_t nbytes, void **ptr) {
    ShardGroup *g = static_cast<ShardGroup *>(group);
    if (!g || !ptr || rank < 0 || rank >= g->world || !g->r[rank].local || nbytes < 0)
        return set_error(DN_ERR_INVALID_ARG, "dn_shard_heap_alloc: bad argument");
    ShardRank &r = g->r[rank];
    const int64_t need = (nbytes + 255) / 256 * 256;
    if (r.heap_used + need > g->heap_bytes)
        return set_error(DN_ERR_OUT_OF_MEMORY, "symmetric heap of rank %d exhausted (%lld of %lld bytes in use)", rank,
                         (long long)r.heap_used, (long long)g->heap_bytes);
    *ptr = r.window + kHeapOff + r.heap_used;
    r.heap_used += need;
    return DN_OK;
}

dn_status dn_shard_heap_reset(void *group, int32_t rank) {
    ShardGroup *g = static_cast<ShardGroup *>(group);
    if (!g || rank < 0 || rank >= g->world || !g->r[rank].local)
        return set_error(DN_ERR_INVALID_ARG, "dn_shard_heap_reset: bad argument");
    g->r[rank].heap_used = 0;
    g->r[rank].last_target = -1;
    return DN_OK;
}

dn_status dn_shard_barrier(void *group, int32_t rank) {
    ShardGroup *g;
    ShardRank *me;
    dn_status st = get_rank(group, rank, g, me, "dn_shard_barrier");
    if (st != DN_OK) return st;
    RankScope scope(*me);
    me->last_target = -1;
    me->nprev = 0;
    const PeerSync ps = make_sync(*g, *me, rank, false);
    if ((st = launch_signal(ps)) != DN_OK) return st;
    return complete(*g, *me, ps, PostJob());
}

// One thread driving several ranks brackets each collective: start; the call on every local rank; end.
dn_status dn_shard_group_start(void *group) {
    ShardGroup *g = static_cast<ShardGroup *>(group);
    if (!g || !g->connected) return set_error(DN_ERR_INVALID_ARG, "dn_shard_group_start: bad group");
    if (g->grouping) return set_error(DN_ERR_INVALID_ARG, "dn_shard_group_start: brackets do not nest");
    g->grouping = true;
    return DN_OK;
}

dn_status dn_shard_group_end(void *group) {
    ShardGroup *g = static_cast<ShardGroup *>(group);
    if (!g || !g->grouping) return set_error(DN_ERR_INVALID_ARG, "dn_shard_group_end without dn_shard_group_start");
    g->grouping = false;
    dn_status first = DN_OK;
    for (int k = 0; k < g->world; ++k) {
        ShardRank &r = g->r[k];
        if (!r.local) continue;
        for (int i = 0; i < r.ncur; ++i) r.prev_targets[i] = r.cur_targets[i];
        r.nprev = r.ncur;
        r.ncur = 0;
        r.last_target = -1;
        if (!r.has_pending) continue;
        r.has_pending = false;
        RankScope scope(r);
        dn_status st = enqueue_wait(r.pending_sync, g->nlocal == 1);
        if (st == DN_OK) st = run_post(r, r.pending_job);
        if (st != DN_OK && first == DN_OK) first = st;
    }
    return first;
}

// ---- reductions over a non-sharded axis: one launch per rank, direct peer stores ------------------------------
static dn_status shard_rows_op(void *group, int32_t rank, const dn_tensor *t_full, int64_t row_begin,
                               const dn_tensor *a_local, const char *what,
                               dn_status (*run)(const dn_tensor *t, const dn_tensor *a, const void *ctx), const void *ctx) {
    ShardGroup *g;
    ShardRank *me;
    dn_status st = get_rank(group, rank, g, me, what);
    if (st != DN_OK) return st;
    if (!tensor_valid(a_local) || a_local->ndims < 2)
        return set_error(DN_ERR_INVALID_ARG, "%s: the slab must have the sharded dim 0 and a reduced last axis", what);
    if ((st = check_full_target(*g, *me, t_full, what)) != DN_OK) return st;
    dn_tensor mine;
    if ((st = slab_view(mine, t_full, row_begin, a_local->shape[0], what)) != DN_OK) return st;
    RankScope scope(*me);
    if ((st = guard_target(*g, *me, rank, data_ptr(t_full))) != DN_OK) return st;
    const PeerSync ps = make_sync(*g, *me, rank, true);
    if ((st = with_pending(ps, [&] { return run(&mine, a_local, ctx); })) != DN_OK) return st;
    return complete(*g, *me, ps, PostJob());
}

dn_status dn_shard_reduce_last_axis(void *group, int32_t rank, int32_t op, const dn_tensor *t_full,
                                    int64_t row_begin, const dn_tensor *a_local) {
    return shard_rows_op(group, rank, t_full, row_begin, a_local, "dn_shard_reduce_last_axis",
                         [](const dn_tensor *t, const dn_tensor *a, const void *c) {
                             return dn_reduce_last_axis(*static_cast<const int32_t *>(c), t, a);
                         }, &op);
}

dn_status dn_shard_arg_reduce_last_axis(void *group, int32_t rank, int32_t op, const dn_tensor *t_full,
                                        int64_t row_begin, const dn_tensor *a_local) {
    return shard_rows_op(group, rank, t_full, row_begin, a_local, "dn_shard_arg_reduce_last_axis",
                         [](const dn_tensor *t, const dn_tensor *a, const void *c) {
                             return dn_arg_reduce_last_axis(*static_cast<const int32_t *>(c), t, a);
                         }, &op);
}

dn_status dn_shard_find_last_axis(void *group, int32_t rank, const void *value, const dn_tensor *t_full,
                                  int64_t row_begin, const dn_tensor *a_local) {
    return shard_rows_op(group, rank, t_full, row_begin, a_local, "dn_shard_find_last_axis",
                         [](const dn_tensor *t, const dn_tensor *a, const void *c) { return dn_find_last_axis(c, t, a); },
                         value);
}

// Fused Min/Max + ArgMin/ArgMax (one pass over the source).
static dn_status minmax_arg_local(int32_t op, const dn_tensor *tv, const dn_tensor *ti, const dn_tensor *a) {
    if (op != DN_ARG_MIN && op != DN_ARG_MAX) return set_error(DN_ERR_INVALID_ARG, "minmax+arg: bad op %d", op);
    RedPlan plan;
    dn_status st = red_make_plan(plan, tv, a, "minmax+arg");
    if (st != DN_OK) return st;
    if (!tensor_valid(ti) || ti->dtype != DN_I64 || !same_shape(tv, ti))
        return set_error(DN_ERR_INVALID_ARG, "minmax+arg: the index target must be int64 with the value target's shape");
    if (tv->dtype != a->dtype || (a->dtype != DN_F32 && a->dtype != DN_F64))
        return set_error(DN_ERR_UNSUPPORTED, "minmax+arg: float32 / float64 sources only");
    for (int d = 0; d < tv->ndims; ++d)
        if (tv->shape[d] > 1 && tv->stride[d] != ti->stride[d])
            return set_error(DN_ERR_INVALID_ARG, "minmax+arg: value and index targets must have equal element strides");
    plan.dst2 = data_ptr(ti);
    if (a->dtype == DN_F32)
        return op == DN_ARG_MAX ? red_run(plan, MinMaxArgOp<float, true>()) : red_run(plan, MinMaxArgOp<float, false>());
    return op == DN_ARG_MAX ? red_run(plan, MinMaxArgOp<double, true>()) : red_run(plan, MinMaxArgOp<double, false>());
}

dn_status dn_shard_minmax_arg_last_axis(void *group, int32_t rank, int32_t op, const dn_tensor *t_val_full,
                                        const dn_tensor *t_idx_full, int64_t row_begin, const dn_tensor *a_local) {
    if (!group) {
        if (row_begin != 0) return set_error(DN_ERR_INVALID_ARG, "minmax+arg: row_begin must be 0 without a group");
        return minmax_arg_local(op, t_val_full, t_idx_full, a_local);
    }
    const char *what = "dn_shard_minmax_arg_last_axis";
    ShardGroup *g;
    ShardRank *me;
    dn_status st = get_rank(group, rank, g, me, what);
    if (st != DN_OK) return st;
    if (!tensor_valid(a_local) || a_local->ndims < 2)
        return set_error(DN_ERR_INVALID_ARG, "%s: the slab must have the sharded dim 0 and a reduced last axis", what);
    if ((st = check_full_target(*g, *me, t_val_full, what)) != DN_OK) return st;
    if ((st = check_full_target(*g, *me, t_idx_full, what)) != DN_OK) return st;
    dn_tensor mv, mi;
    if ((st = slab_view(mv, t_val_full, row_begin, a_local->shape[0], what)) != DN_OK) return st;
    if ((st = slab_view(mi, t_idx_full, row_begin, a_local->shape[0], what)) != DN_OK) return st;
    RankScope scope(*me);
    if ((st = guard_target(*g, *me, rank, data_ptr(t_val_full))) != DN_OK) return st;
    const PeerSync ps = make_sync(*g, *me, rank, true);
    if ((st = with_pending(ps, [&] { return minmax_arg_local(op, &mv, &mi, a_local); })) != DN_OK) return st;
    return complete(*g, *me, ps, PostJob());
}

dn_status dn_shard_all_gather_rows(void *group, int32_t rank, const dn_tensor *t_full, int64_t row_begin,
                                   int64_t nrows) {
    const char *what = "dn_shard_all_gather_rows";
    ShardGroup *g;
    ShardRank *me;
    dn_status st = get_rank(group, rank, g, me, what);
    if (st != DN_OK) return st;
    if ((st = check_full_target(*g, *me, t_full, what)) != DN_OK) return st;
    if (!is_c_contiguous(t_full)) return set_error(DN_ERR_INVALID_ARG, "%s: the full result must be C-contiguous", what);
    dn_tensor mine;
    if ((st = slab_view(mine, t_full, row_begin, nrows, what)) != DN_OK) return st;
    RankScope scope(*me);
    me->last_target = data_ptr(t_full) - me->window;
    const PeerSync ps = make_sync(*g, *me, rank, true);
    if ((st = launch_push(ps, data_ptr(&mine), num_elements(&mine) * dtype_size(mine.dtype))) != DN_OK) return st;
    return complete(*g, *me, ps, PostJob());
}

}  // extern "C"

// ---- reductions over the sharded axis: partials into every rank's scratch, local ordered fold -----------------
namespace dn {
namespace {

struct Strided {  // a target view addressed by the row-major linear index of its elements
    int32_t ndims;
    int64_t shape[DN_MAX_DIMS], stride[DN_MAX_DIMS];  // elements
    __device__ int64_t offset(int64_t o) const {
        int64_t off = 0;
        for (int d = ndims - 1; d >= 0; --d) {
            const int64_t q = o / shape[d];
            off += (o - q * shape[d]) * stride[d];
            o = q;
        }
        return off;
    }
};

Strided strided_of(const dn_tensor *t) {
    Strided s = {};
    s.ndims = t->ndims;
    for (int d = 0; d < t->ndims; ++d) {
        s.shape[d] = t->shape[d];
        s.stride[d] = t->stride[d];
    }
    return s;
}

// The state-emitting form of an operator: the reduction stores its State instead of the finalized value.
template <class Op>
struct StateOf : Op {
    using Out = typename Op::State;
    __device__ static Out finalize(typename Op::State s) { return s; }
};

enum FoldMode { kFoldState = 0, kFoldValue = 1, kFoldArg = 2, kFoldFind = 3 };

// out[o] = finalize(combine over ranks r = 0..W-1, in rank order, of slot r's partial for output o).
//   kFoldState: slots hold Op::State (float Min/Max);   kFoldValue: slots hold Op::Out, State == conversion of it;
//   kFoldArg:   slots hold ArgOp::State with GLOBAL indices, NotFound partials never win;
//   kFoldFind:  slots hold int64 global indices or NotFound; the lowest found index wins.
template <class Op, int MODE>
__global__ void shard_fold_kernel(char *out, Strided t, const char *slots, int64_t slot_bytes, int W, int64_t n) {
    using State = typename Op::State;
    using Out = typename Op::Out;
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n) return;
    State acc = Op::identity();
    bool any = false;
    for (int r = 0; r < W; ++r) {
        const char *slot = slots + r * slot_bytes;
        State s;
        if constexpr (MODE == kFoldValue) s = static_cast<State>(reinterpret_cast<const Out *>(slot)[o]);
        else s = reinterpret_cast<const State *>(slot)[o];
        if constexpr (MODE == kFoldArg) {
            if (s.idx == (int64_t)DN_NOT_FOUND) continue;
        }
        if constexpr (MODE == kFoldFind) {
            if (s == (int64_t)DN_NOT_FOUND) continue;
        }
        acc = any ? Op::combine(acc, s) : s;
        any = true;
    }
    reinterpret_cast<Out *>(out)[t.offset(o)] = Op::finalize(acc);
}

// Find's partials are plain indices: State == int64 with INT64_MAX as "none"; reuse FindOp's combine / finalize.
template <class Op, int MODE>
dn_status launch_fold(const dn_tensor *t, const char *slots, int64_t slot_bytes, int W, int64_t n) {
    if (n == 0) return DN_OK;
    DN_LAUNCH((shard_fold_kernel<Op, MODE>), (unsigned)((n + 255) / 256), 256, 0, data_ptr(t), strided_of(t), slots,
              slot_bytes, W, n);
    return launch_status("shard fold kernel");
}

// (value at the arg position, GLOBAL index) pairs from the local ArgMin/ArgMax result.
template <class T>
__global__ void shard_arg_pairs_kernel(const __grid_constant__ RedParams p, const int64_t *idx, char *states, int64_t base) {
    using State = typename ArgOp<T, true>::State;
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= p.nrows) return;
    int64_t soff, toff;
    red_offsets(p.outer, (uint32_t)r, soff, toff);  // toff: bytes into an int64 array laid out like the slot
    const int64_t o = toff / 8;
    const int64_t i = idx[o];
    State s;
    s.idx = i;
    s.val = T();
    if (i != (int64_t)DN_NOT_FOUND) {
        s.val = *reinterpret_cast<const T *>(p.src + soff + i * p.lstride);
        s.idx = i + base;
    }
    reinterpret_cast<State *>(states)[o] = s;
}

__global__ void shard_shift_found_kernel(int64_t *idx, int64_t n, int64_t base) {
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o < n && idx[o] != (int64_t)DN_NOT_FOUND) idx[o] += base;
}

__global__ void shard_shift_column_kernel(int64_t *col, int64_t rows, int64_t row_stride, int64_t base) {
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o < rows) col[o * row_stride] += base;
}

template <class T>
dn_status fold_value(int op, const dn_tensor *t, const char *slots, int64_t sb, int W, int64_t n) {
    switch (op) {
    case DN_SUM: return launch_fold<SumOp<T>, kFoldValue>(t, slots, sb, W, n);
    case DN_PRODUCT: return launch_fold<ProductOp<T>, kFoldValue>(t, slots, sb, W, n);
    case DN_MIN:
        if constexpr (kIsFloat<T>) return launch_fold<MinMaxFloatOp<T, false>, kFoldState>(t, slots, sb, W, n);
        else return launch_fold<MinMaxIntOp<T, false>, kFoldValue>(t, slots, sb, W, n);
    default:
        if constexpr (kIsFloat<T>) return launch_fold<MinMaxFloatOp<T, true>, kFoldState>(t, slots, sb, W, n);
        else return launch_fold<MinMaxIntOp<T, true>, kFoldValue>(t, slots, sb, W, n);
    }
}

// Float Min/Max partial as the operator's STATE (value after the last NaN, saw-NaN / non-empty flags): the exact
// monoid of the host's order-dependent fold (ScalarOps.fs:620-628), so the cross-rank fold reproduces it bit for bit.
template <class T>
dn_status minmax_state_partial(int op, char *slot, const dn_tensor *slot_like, const dn_tensor *a_local) {
    RedPlan plan;
    dn_status st = red_make_plan(plan, slot_like, a_local, "sharded Min/Max");
    if (st != DN_OK) return st;
    using S = typename MinMaxFloatOp<T, true>::State;
    plan.dst = slot;
    plan.out_size = (int)sizeof(S);
    if (op == DN_MAX) return red_run(plan, StateOf<MinMaxFloatOp<T, true>>());
    return red_run(plan, StateOf<MinMaxFloatOp<T, false>>());
}

}  // namespace
}  // namespace dn

extern "C" {

dn_status dn_shard_reduce_sharded_axis(void *group, int32_t rank, int32_t kind, int32_t op, const void *value,
                                       const dn_tensor *t, int64_t axis_begin, const dn_tensor *a_local) {
    const char *what = "dn_shard_reduce_sharded_axis";
    ShardGroup *g;
    ShardRank *me;
    dn_status st = get_rank(group, rank, g, me, what);
    if (st != DN_OK) return st;
    if (!tensor_valid(t) || !tensor_valid(a_local) || a_local->ndims < 1 || t->ndims != a_local->ndims - 1)
        return set_error(DN_ERR_INVALID_ARG, "%s: source rank must be target rank + 1", what);
    for (int d = 0; d < t->ndims; ++d)
        if (t->shape[d] != a_local->shape[d]) return set_error(DN_ERR_SHAPE_MISMATCH, "%s: target shape", what);
    if (kind < 0 || kind > 2 || (kind == 2 && !value)) return set_error(DN_ERR_INVALID_ARG, "%s: bad kind", what);
    if (kind == 0 && (op < 0 || op >= DN_REDUCE_OP_COUNT)) return set_error(DN_ERR_INVALID_ARG, "%s: bad op", what);
    if (kind == 1 && op != DN_ARG_MIN && op != DN_ARG_MAX) return set_error(DN_ERR_INVALID_ARG, "%s: bad op", what);
    const int W = g->world;
    const int64_t n = num_elements(t);
    const int in_dt = a_local->dtype;
    const bool is_float = in_dt == DN_F32 || in_dt == DN_F64;
    const bool minmax_float = kind == 0 && (op == DN_MIN || op == DN_MAX) && is_float;
    if (kind == 0) {
        const int want = op == DN_COUNT_TRUE ? DN_I64 : in_dt;
        if (t->dtype != want || ((op == DN_ALL || op == DN_ANY || op == DN_COUNT_TRUE) != (in_dt == DN_BOOL)))
            return set_error(DN_ERR_INVALID_ARG, "%s: source / target types do not fit the operator", what);
    } else if (t->dtype != DN_I64 || (kind == 1 && in_dt == DN_BOOL)) {
        return set_error(DN_ERR_INVALID_ARG, "%s: the target must be int64 (and the source not bool for Arg*)", what);
    }
    // one slot per rank in this use's scratch half; element = the partial (Out, or State for float Min/Max / Arg)
    int64_t elem = dtype_size(t->dtype);
    if (minmax_float) elem = in_dt == DN_F32 ? 8 : 16;
    if (kind == 1) elem = 16;  // ArgOp<T>::State = {T val; int64 idx}
    const int64_t slot_bytes = (n * elem + 255) / 256 * 256;
    const int64_t tmp_bytes = kind == 1 ? (n * 8 + 255) / 256 * 256 : 0;  // local ArgMin/ArgMax indices
    if (W * slot_bytes + tmp_bytes > kScratchHalf)
        return set_error(DN_ERR_UNSUPPORTED, "%s: %lld outputs x %d ranks exceed the partial-result scratch", what,
                         (long long)n, W);
    RankScope scope(*me);
    char *half = me->window + kFlagBytes + (me->scratch_uses++ & 1) * kScratchHalf;
    char *slot = half + rank * slot_bytes;
    // a C-contiguous tensor of the target's shape over the slot (value-typed partials are written through it)
    dn_tensor slot_t = make_contig(slot, 0, t->dtype, t->ndims, t->shape);

    // 1. local partial into this rank's slot
    if (kind == 0 && minmax_float) {
        dn_tensor like = make_contig(slot, 0, DN_I64, t->ndims, t->shape);  // shape / layout only
        like.dtype = in_dt;
        st = in_dt == DN_F32 ? minmax_state_partial<float>(op, slot, &like, a_local)
                             : minmax_state_partial<double>(op, slot, &like, a_local);
    } else if (kind == 0) {
        st = dn_reduce_last_axis(op, &slot_t, a_local);
    } else if (kind == 1) {
        dn_tensor idx_t = make_contig(half + W * slot_bytes, 0, DN_I64, t->ndims, t->shape);
        st = dn_arg_reduce_last_axis(op, &idx_t, a_local);
        if (st == DN_OK && n > 0) {
            RedPlan plan;
            st = red_make_plan(plan, &idx_t, a_local, what);
            if (st == DN_OK && plan.nrows > 0) {
                RedParams p = {};
                red_fill_outer(p.outer, plan);
                p.src = plan.src;
                p.nrows = (uint32_t)plan.nrows;
                p.lstride = plan.lstride_elems * plan.in_size;
                const unsigned grid = (unsigned)((plan.nrows + 255) / 256);
                const int64_t *idx = reinterpret_cast<const int64_t *>(data_ptr(&idx_t));
                DN_SWITCH_DTYPE(in_dt, {
                    if constexpr (!kIsBool<T>) DN_LAUNCH((shard_arg_pairs_kernel<T>), grid, 256, 0, p, idx, slot, axis_begin);
                });
                st = launch_status("shard arg pairs kernel");
            }
        }
    } else {
        st = dn_find_last_axis(value, &slot_t, a_local);
        if (st == DN_OK && n > 0 && axis_begin != 0) {
            DN_LAUNCH(shard_shift_found_kernel, (unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<int64_t *>(slot), n,
                      axis_begin);
            st = launch_status("shard find shift kernel");
        }
    }
    if (st != DN_OK) return st;
    // 2. replicate the slot into every rank's scratch and signal; 3. (after the wait) fold the W partials locally
    me->last_target = -1;
    const PeerSync ps = make_sync(*g, *me, rank, true);
    if ((st = launch_push(ps, slot, n * elem)) != DN_OK) return st;
    PostJob job;
    job.kind = 1;
    job.red_kind = kind;
    job.op = op;
    job.in_dt = in_dt;
    job.world = W;
    job.t = *t;
    job.half = half;
    job.slot_bytes = slot_bytes;
    job.n = n;
    return complete(*g, *me, ps, job);
}

}  // extern "C"

namespace dn {
namespace {

// Runs on the rank's stream (RankScope active) after the wait.
dn_status run_post(ShardRank &me, const PostJob &job) {
    if (job.kind == 2) {
        DN_CUDA_TRY(cudaMemcpyAsync(me.host_counts, job.counts_dev, sizeof(int64_t) * job.world, cudaMemcpyDeviceToHost,
                                    current_stream()));
        return DN_OK;
    }
    if (job.kind != 1) return DN_OK;
    // local fold of the W partials in rank order (identical bits on every rank)
    const dn_tensor *t = &job.t;
    const char *half = job.half;
    const int64_t slot_bytes = job.slot_bytes, n = job.n;
    const int W = job.world, op = job.op;
    if (job.red_kind == 0) {
        if (op == DN_COUNT_TRUE) return launch_fold<SumOp<int64_t>, kFoldValue>(t, half, slot_bytes, W, n);
        if (op == DN_ALL) return launch_fold<AllAnyOp<true>, kFoldValue>(t, half, slot_bytes, W, n);
        if (op == DN_ANY) return launch_fold<AllAnyOp<false>, kFoldValue>(t, half, slot_bytes, W, n);
        DN_SWITCH_DTYPE(job.in_dt, {
            if constexpr (!kIsBool<T>) return fold_value<T>(op, t, half, slot_bytes, W, n);
        });
        return DN_OK;
    }
    if (job.red_kind == 1) {
        DN_SWITCH_DTYPE(job.in_dt, {
            if constexpr (!kIsBool<T>) {
                if (op == DN_ARG_MAX) return launch_fold<ArgOp<T, true>, kFoldArg>(t, half, slot_bytes, W, n);
                return launch_fold<ArgOp<T, false>, kFoldArg>(t, half, slot_bytes, W, n);
            }
        });
        return DN_OK;
    }
    return launch_fold<FindOp<int64_t>, kFoldFind>(t, half, slot_bytes, W, n);
}

}  // namespace
}  // namespace dn

extern "C" {

// ---- ordered compaction over the shards (row-major order of the full tensor == rank order of the slabs) ----------
dn_status dn_shard_count_true_begin(void *group, int32_t rank, const dn_tensor *mask_local) {
    const char *what = "dn_shard_count_true";
    ShardGroup *g;
    ShardRank *me;
    dn_status st = get_rank(group, rank, g, me, what);
    if (st != DN_OK) return st;
    RankScope scope(*me);
    // counts[2][8] in the window, alternating per use (a fast peer may already be publishing its NEXT count)
    int64_t *slot = reinterpret_cast<int64_t *>(me->window + kCountsOff) + (me->scratch_uses++ & 1) * kMaxShardRanks;
    if ((st = count_true_async(mask_local, reinterpret_cast<unsigned long long *>(slot + rank))) != DN_OK) return st;
    me->last_target = -1;
    const PeerSync ps = make_sync(*g, *me, rank, true);
    if ((st = launch_push(ps, reinterpret_cast<char *>(slot + rank), sizeof(int64_t))) != DN_OK) return st;
    PostJob job;
    job.kind = 2;
    job.world = g->world;
    job.counts_dev = slot;
    return complete(*g, *me, ps, job);
}

dn_status dn_shard_count_true_end(void *group, int32_t rank, int64_t *counts) {
    ShardGroup *g;
    ShardRank *me;
    dn_status st = get_rank(group, rank, g, me, "dn_shard_count_true");
    if (st != DN_OK) return st;
    if (!counts) return set_error(DN_ERR_INVALID_ARG, "dn_shard_count_true: null counts");
    if ((st = dn_shard_sync(group, rank)) != DN_OK) return st;
    memcpy(counts, me->host_counts, sizeof(int64_t) * g->world);
    return DN_OK;
}

dn_status dn_shard_count_true(void *group, int32_t rank, const dn_tensor *mask_local, int64_t *counts) {
    dn_status st = dn_shard_count_true_begin(group, rank, mask_local);
    return st != DN_OK ? st : dn_shard_count_true_end(group, rank, counts);
}

dn_status dn_shard_true_indices(void *group, int32_t rank, const dn_tensor *t_full, int64_t row_offset,
                                int64_t nrows_local, const dn_tensor *mask_local, int64_t dim0_begin) {
    const char *what = "dn_shard_true_indices";
    ShardGroup *g;
    ShardRank *me;
    dn_status st = get_rank(group, rank, g, me, what);
    if (st != DN_OK) return st;
    if ((st = check_full_target(*g, *me, t_full, what)) != DN_OK) return st;
    if (t_full->ndims != 2 || t_full->dtype != DN_I64 || !is_c_contiguous(t_full))
        return set_error(DN_ERR_INVALID_ARG, "%s: the full result must be a C-contiguous int64 [nTrue, nDims]", what);
    dn_tensor mine;
    if ((st = slab_view(mine, t_full, row_offset, nrows_local, what)) != DN_OK) return st;
    RankScope scope(*me);
    if ((st = guard_target(*g, *me, rank, data_ptr(t_full))) != DN_OK) return st;
    if (nrows_local > 0) {
        if ((st = dn_true_indices(&mine, mask_local)) != DN_OK) return st;
        if (dim0_begin != 0) {
            DN_LAUNCH(shard_shift_column_kernel, (unsigned)((nrows_local + 255) / 256), 256, 0,
                      reinterpret_cast<int64_t *>(data_ptr(&mine)), nrows_local, mine.stride[0], dim0_begin);
            if ((st = launch_status("shard column shift kernel")) != DN_OK) return st;
        }
    }
    const PeerSync ps = make_sync(*g, *me, rank, true);
    if ((st = launch_push(ps, data_ptr(&mine), nrows_local * t_full->shape[1] * 8)) != DN_OK) return st;
    return complete(*g, *me, ps, PostJob());
}

dn_status dn_shard_masked_get(void *group, int32_t rank, const dn_tensor *t_full, int64_t elem_offset,
                              int64_t nelems_local, const dn_tensor *a_local, const dn_tensor *mask_local) {
    const char *what = "dn_shard_masked_get";
    ShardGroup *g;
    ShardRank *me;
    dn_status st = get_rank(group, rank, g, me, what);
    if (st != DN_OK) return st;
    if ((st = check_full_target(*g, *me, t_full, what)) != DN_OK) return st;
    if (!tensor_valid(a_local) || !tensor_valid(mask_local) || t_full->ndims != 1 || !is_c_contiguous(t_full) ||
        t_full->dtype != a_local->dtype || !same_shape(a_local, mask_local))
        return set_error(DN_ERR_INVALID_ARG, "%s: expected a C-contiguous 1-D result of the source's type and a mask of "
                                             "the source's shape", what);
    dn_tensor mine;
    if ((st = slab_view(mine, t_full, elem_offset, nelems_local, what)) != DN_OK) return st;
    RankScope scope(*me);
    if ((st = guard_target(*g, *me, rank, data_ptr(t_full))) != DN_OK) return st;
    if (nelems_local > 0) {
        // the flattened walk: both views as 1-D when they flatten without a copy, else through contiguous copies
        const int64_t ne = num_elements(a_local);
        void *tmp_a = nullptr, *tmp_m = nullptr;
        dn_tensor a1 = make_contig(nullptr, 0, a_local->dtype, 1, &ne), m1 = make_contig(nullptr, 0, DN_BOOL, 1, &ne);
        if (is_c_contiguous(a_local)) {
            a1.base = a_local->base;
            a1.offset = a_local->offset;
        } else {
            if ((st = scratch_alloc((size_t)ne * dtype_size(a_local->dtype), &tmp_a)) != DN_OK) return st;
            dn_tensor c = make_contig(tmp_a, 0, a_local->dtype, a_local->ndims, a_local->shape);
            st = dn_copy(&c, a_local);
            a1.base = tmp_a;
        }
        if (st == DN_OK) {
            if (is_c_contiguous(mask_local)) {
                m1.base = mask_local->base;
                m1.offset = mask_local->offset;
            } else {
                st = scratch_alloc((size_t)ne, &tmp_m);
                if (st == DN_OK) {
                    dn_tensor c = make_contig(tmp_m, 0, DN_BOOL, mask_local->ndims, mask_local->shape);
                    st = dn_copy(&c, mask_local);
                    m1.base = tmp_m;
                }
            }
        }
        if (st == DN_OK) {
            const dn_tensor *masks[1] = {&m1};
            st = dn_masked_get(&mine, &a1, masks, 1);
        }
        scratch_free(tmp_a);
        scratch_free(tmp_m);
        if (st != DN_OK) return st;
    }
    const PeerSync ps = make_sync(*g, *me, rank, true);
    if ((st = launch_push(ps, data_ptr(&mine), nelems_local * dtype_size(t_full->dtype))) != DN_OK) return st;
    return complete(*g, *me, ps, PostJob());
}

}  // extern "C"
