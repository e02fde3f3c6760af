// index.cu — Gather, Scatter, CountTrue, MaskedGet, MaskedSet, TrueIndices (Tensor/Tensor/TensorBackend.fs:119-123).
//
// Gather / Scatter replace CudaBackend.fs:362-381, CudaKernels.fs:302-345 and Kernels/GatherScatter.cuh:26-114.
// MaskedGet / MaskedSet / TrueIndices / CountTrue do not exist in the reference's CUDA backend at all
// (CudaBackend.fs:489-492 raise NotSupportedException); their semantics are the host's (ScalarOps.fs:667-707):
// walk the tensor in LOGICAL row-major order of the view, whatever its strides.
//
// Ordered compaction: one single-pass primitive shared by TrueIndices, MaskedGet and MaskedSet. Every CTA draws a
// tile of 8192 or 32768 consecutive logical positions from a ticket counter (thread = 16 consecutive positions
// per chunk, one 128-bit load when they are contiguous in memory, all loads of the tile issued back to back), ranks
// its true elements with a register popcount + shuffle scan, obtains the number of true elements in all earlier
// tiles by decoupled look-back over a 64-bit {status, count} word per tile (the mask is read from HBM exactly once,
// no count pass, no scan kernel), stages the selected positions in shared memory so that consecutive threads own
// consecutive ranks, and hands (rank, logical position) to a sink: coordinates (TrueIndices), source -> dense
// target (MaskedGet), dense values -> target (MaskedSet), or a plain index list (per-dimension masks, which select
// a cartesian product: ScalarOps.fs:672-681). Sinks are two-phase (load, store) so that four independent loads are
// in flight per thread; MaskedGet's source loads are issued before the look-back returns.
#include <cstdlib>

#include "ew_ops.cuh"

using namespace dn;

namespace {

constexpr int kIdxThreads = 256;

// ---------------------------------------------------------------------------------------------------------------
// Gather / Scatter
// ---------------------------------------------------------------------------------------------------------------
struct GSParams {
    char *it_ptr;       // tensor that is walked: target (gather) / source (scatter)
    char *other_ptr;    // tensor that is indexed: source (gather) / target (scatter)
    int32_t nd_it, nd_other;
    uint32_t n;
    uint32_t it_shape[DN_MAX_DIMS];     // innermost-first
    FastDiv it_div[DN_MAX_DIMS];
    int64_t it_stride[DN_MAX_DIMS];     // bytes, innermost-first
    int64_t other_shape[DN_MAX_DIMS];   // descriptor order
    int64_t other_stride[DN_MAX_DIMS];  // bytes, descriptor order
    const char *idx_ptr[DN_MAX_DIMS];   // per indexed dim (descriptor order), nullptr = None
    int64_t idx_stride[DN_MAX_DIMS][DN_MAX_DIMS];  // bytes, [indexed dim][walked dim innermost-first]
    int *err;
};

// Decomposes walked position `f` into the byte offset into the walked tensor and the byte offset into the indexed
// tensor; returns false if an index is out of range. NDI / NDO: compile-time ranks of the walked / indexed tensor
// (0 = take them from the parameter block); the common small ranks get loop-free index math.
template <int NDI, int NDO>
__device__ __forceinline__ bool gs_addresses(const GSParams &p, uint32_t f, int64_t &it_off, int64_t &other_off) {
    const int nd_it = NDI > 0 ? NDI : p.nd_it;
    const int nd_other = NDO > 0 ? NDO : p.nd_other;
    constexpr int kMaxI = NDI > 0 ? NDI : DN_MAX_DIMS;
    constexpr int kMaxO = NDO > 0 ? NDO : DN_MAX_DIMS;
    uint32_t pos[kMaxI];
    uint32_t rem = f;
    it_off = 0;
#pragma unroll
    for (int k = 0; k < kMaxI; ++k) {
        if (k >= nd_it) break;
        uint32_t q, x;
        if (k == nd_it - 1) { x = rem; q = 0; }
        else { q = p.it_div[k].div(rem); x = rem - q * p.it_shape[k]; }
        pos[k] = x;
        it_off += (int64_t)x * p.it_stride[k];
        rem = q;
    }
    other_off = 0;
    bool ok = true;
#pragma unroll
    for (int d = 0; d < kMaxO; ++d) {
        if (d >= nd_other) break;
        int64_t ix;
        if (p.idx_ptr[d]) {
            int64_t io = 0;
#pragma unroll
            for (int k = 0; k < kMaxI; ++k) {
                if (k >= nd_it) break;
                io += (int64_t)pos[k] * p.idx_stride[d][k];
            }
            ix = *reinterpret_cast<const int64_t *>(p.idx_ptr[d] + io);
        } else {
            ix = 0;  // None: identity on descriptor dim d (walked dim nd_it-1-d, innermost-first)
#pragma unroll
            for (int k = 0; k < kMaxI; ++k)
                if (k == nd_it - 1 - d) ix = pos[k];
        }
        ok = ok && ix >= 0 && ix < p.other_shape[d];
        other_off += ix * p.other_stride[d];
    }
    return ok;
}

template <class B, int NDI, int NDO>
__global__ void __launch_bounds__(kIdxThreads) gather_kernel(const __grid_constant__ GSParams p) {
    constexpr int U = 4;
    for (uint64_t base = (uint64_t)blockIdx.x * (kIdxThreads * U); base < p.n; base += (uint64_t)gridDim.x * (kIdxThreads * U)) {
        int64_t to[U], so[U];
        int state[U];  // 0 = inactive, 1 = in range, 2 = index out of range
        B v[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const uint64_t f = base + (uint32_t)j * kIdxThreads + threadIdx.x;
            state[j] = 0;
            if (f < p.n) state[j] = gs_addresses<NDI, NDO>(p, (uint32_t)f, to[j], so[j]) ? 1 : 2;
        }
#pragma unroll
        for (int j = 0; j < U; ++j)
            if (state[j] == 1) v[j] = *reinterpret_cast<const B *>(p.other_ptr + so[j]);
#pragma unroll
        for (int j = 0; j < U; ++j) {
            if (state[j] == 1) *reinterpret_cast<B *>(p.it_ptr + to[j]) = v[j];
            else if (state[j] == 2) atomicExch(p.err, 1);
        }
    }
}

// atomic add for the element types that have a native red/atom: f32, f64, i32, u32, u64; int64 goes through the
// unsigned 64-bit add (two's complement wrap-around is the same operation). GatherScatter.cuh:89 relies on
// atomicAdd overloads and therefore cannot instantiate int64 at all. 8- and 16-bit targets never reach this
// function: dn_scatter accumulates them in a 32-bit scratch tensor and narrows afterwards (no sub-word CAS, no
// access outside the target's own bytes).
template <class T> __device__ __forceinline__ void atomic_add_any(T *addr, T v) {
    if constexpr (sizeof(T) == 8 && std::is_integral<T>::value) {
        atomicAdd(reinterpret_cast<unsigned long long *>(addr), (unsigned long long)v);
    } else {
        atomicAdd(addr, v);
    }
}

// Lanes of a warp that add into the SAME target element are combined first (match.any on the target offset, then a
// segmented shuffle reduction) and only the group's first lane issues the atomic: index distributions with hot
// spots no longer serialise on one L2 atomic unit (2^26 adds onto 4096 cells: 20 ms without).
// Warps without duplicates pay one match + one vote.
template <class T>
__device__ __forceinline__ void warp_aggregated_add(char *base, int64_t off, T v, bool active) {
    const int lane = threadIdx.x & 31;
    const unsigned long long key = active ? (unsigned long long)off : ~(unsigned long long)lane;  // inactive: unique
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    if (__all_sync(0xffffffffu, (peers & (peers - 1)) == 0)) {  // no two lanes share a target
        if (active) atomic_add_any(reinterpret_cast<T *>(base + off), v);
        return;
    }
    const int leader = __ffs(peers) - 1;
    unsigned rest = peers & ~(1u << leader);  // the leader accumulates everybody else's value
    T sum = v;
    while (__any_sync(0xffffffffu, rest != 0)) {
        const int src = rest ? __ffs(rest) - 1 : lane;
        T o;
        if constexpr (sizeof(T) == 8) {
            const unsigned long long bits = __shfl_sync(0xffffffffu, *reinterpret_cast<const unsigned long long *>(&v), src);
            o = *reinterpret_cast<const T *>(&bits);
        } else {
            const unsigned bits = __shfl_sync(0xffffffffu, *reinterpret_cast<const unsigned *>(&v), src);
            o = *reinterpret_cast<const T *>(&bits);
        }
        if (rest) {
            if constexpr (std::is_integral<T>::value) sum = (T)((UnsignedT<T>)sum + (UnsignedT<T>)o);
            else sum = sum + o;
            rest &= rest - 1;
        }
    }
    if (active && lane == leader) atomic_add_any(reinterpret_cast<T *>(base + off), sum);
}

template <class TS, class TA, int NDI, int NDO>
__global__ void __launch_bounds__(kIdxThreads) scatter_kernel(const __grid_constant__ GSParams p) {
    using T = TA;
    constexpr int U = 4;
    // whole warps stay in the loop together (the aggregation uses full-warp votes): the bound is per warp
    for (uint64_t base = (uint64_t)blockIdx.x * (kIdxThreads * U); base < p.n; base += (uint64_t)gridDim.x * (kIdxThreads * U)) {
        int64_t so[U], to[U];
        int state[U];
        T v[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const uint64_t f = base + (uint32_t)j * kIdxThreads + threadIdx.x;
            state[j] = 0;
            to[j] = 0;
            v[j] = T(0);
            if (f < p.n) {
                state[j] = gs_addresses<NDI, NDO>(p, (uint32_t)f, so[j], to[j]) ? 1 : 2;
                v[j] = (TA)*reinterpret_cast<const TS *>(p.it_ptr + so[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < U; ++j) {
            warp_aggregated_add<T>(p.other_ptr, to[j], v[j], state[j] == 1);
            if (state[j] == 2) atomicExch(p.err, 1);
        }
    }
}

// Rank dispatch: (walked rank, indexed rank) in {1,2,3}^2 are compiled with loop-free index math.
#define DN_GS_RANK_DISPATCH(NDI_RT, NDO_RT, ...)                                  \
    do {                                                                          \
        const int _ri = (NDI_RT), _ro = (NDO_RT);                                 \
        if (_ri == 1 && _ro == 1) { constexpr int NDI = 1, NDO = 1; __VA_ARGS__; } \
        else if (_ri == 2 && _ro == 2) { constexpr int NDI = 2, NDO = 2; __VA_ARGS__; } \
        else if (_ri == 2 && _ro == 1) { constexpr int NDI = 2, NDO = 1; __VA_ARGS__; } \
        else if (_ri == 1 && _ro == 2) { constexpr int NDI = 1, NDO = 2; __VA_ARGS__; } \
        else if (_ri == 3 && _ro == 3) { constexpr int NDI = 3, NDO = 3; __VA_ARGS__; } \
        else { constexpr int NDI = 0, NDO = 0; __VA_ARGS__; }                     \
    } while (0)

// ---------------------------------------------------------------------------------------------------------------
// Scatter into a LARGE dense target: partition by target bin, then accumulate in shared memory.
//
// One global atomic per element is bound by the L2 atomic units (2^26 random int64 adds: 2.9 ms, 14 % of the HBM
// rate — 2.5 ms even into a target that lives in L2). Instead the (target index, value) pairs are PARTITIONED by
// target bin (a bin = the 64 KiB of the target one CTA can hold in shared memory), then every bin is summed with
// shared-memory atomics and written out once, coalesced — which is also the zero-fill (CudaBackend.fs:379).
// The partition runs in TWO levels (<= 64 coarse bins, then <= 128 fine bins inside each): every CTA then has at
// most ~128 output runs of thousands of elements open, which L2 merges into whole lines. (A single 8192-way pass
// keeps CTAs x bins = 2.4 M runs of ~27 elements open, far more lines than L2 holds: measured 5.6 ms.)
//   level 1   hist (reads the indices) -> offsets -> partition: (index u32, value) pairs grouped by coarse bin
//   level 2   hist (reads the u32 indices) -> offsets -> partition: (index-in-bin u16, value) grouped by fine bin
//   accumulate: one CTA per fine bin: zero 64 KiB of shared memory, add the bin's pairs, store the bin
// Integer sums are exact in any order; floating-point sums are order-dependent exactly as with global atomics.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kBinThreads = 512;
constexpr int kBinSmemBytes = 64 * 1024;
constexpr int kMaxBins = 8192;
constexpr uint32_t kSpreadBins = 64;  // histogram passes with at most this many bins keep one counter column per lane

struct BinParams {
    GSParams gs;
    uint32_t n;                       // elements
    uint32_t nbins, bin_log, chunk;   // bin of an element = lin >> bin_log; positions per CTA
    uint32_t elem_log;                // log2(sizeof accumulator element)
    uint64_t nt;                      // target elements
    uint32_t *cta_hist;               // [gridDim.x][nbins]: per-CTA counts of the histogram pass
    uint32_t *bin_start;              // [nbins + 1]
    const uint32_t *in_lin;           // level 2 input: pairs grouped by coarse bin
    const char *in_vals;
    uint32_t *out_lin;                // level 1 output
    uint16_t *out_low;                // level 2 output: index inside the fine bin
    char *out_vals;
    uint32_t *cursor;                 // [nbins]: next free slot of every bin of this level (starts as bin_start)
    const uint32_t *coarse_start;     // level 2: [ncoarse + 1] starts of the coarse bins in the level-1 output
    uint32_t ncoarse;
};

// Linear target element index of walked position f, or false when an index is out of range.
template <int NDI, int NDO>
__device__ __forceinline__ bool bin_target(const BinParams &p, uint32_t f, int64_t &src_off, uint32_t &lin) {
    int64_t to;
    const bool ok = gs_addresses<NDI, NDO>(p.gs, f, src_off, to);
    lin = (uint32_t)((uint64_t)to >> p.elem_log);
    return ok;
}

// PAIRS = false: level 1, positions are walked through the index tensors; PAIRS = true: level 2, positions are pairs.
template <bool PAIRS, int NDI, int NDO>
__global__ void __launch_bounds__(kBinThreads) scatter_hist_kernel(const __grid_constant__ BinParams p) {
    extern __shared__ uint32_t sh_hist[];
    // Few bins (level 1: <= 64 coarse bins): 512 threads adding into a handful of counters collide on the same address
    // all the time, and shared-memory atomics serialise those. Every lane gets its own column of counters instead
    // (counter = bin * 32 + lane: distinct addresses AND distinct banks inside a warp), summed at the end.
    const bool spread = p.nbins <= kSpreadBins;
    const uint32_t ncounters = spread ? p.nbins * 32 : p.nbins;
    const uint32_t lane = threadIdx.x & 31;
    for (uint32_t b = threadIdx.x; b < ncounters; b += kBinThreads) sh_hist[b] = 0;
    __syncthreads();
    const uint64_t begin = (uint64_t)blockIdx.x * p.chunk;
    uint64_t end = begin + p.chunk;
    if (end > p.n) end = p.n;
    constexpr int U = 8;  // loads of eight positions in flight per thread (four: 2.1 TB/s over the index tensor)
    for (uint64_t base = begin; base < end; base += (uint64_t)kBinThreads * U) {
        uint32_t lin[U];
        int state[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const uint64_t f = base + (uint32_t)j * kBinThreads + threadIdx.x;
            state[j] = 0;
            if (f < end) {
                if constexpr (PAIRS) {
                    lin[j] = p.in_lin[f];
                    state[j] = 1;
                } else {
                    int64_t so;
                    state[j] = bin_target<NDI, NDO>(p, (uint32_t)f, so, lin[j]) ? 1 : 2;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < U; ++j) {
            // an element with an out-of-range index raises the error flag and travels on as "add 0 to element 0",
            // so that both levels see the same number of elements
            if (state[j] == 2) { atomicExch(p.gs.err, 1); lin[j] = 0; }
            if (state[j] != 0) {
                const uint32_t bin = lin[j] >> p.bin_log;
                atomicAdd(&sh_hist[spread ? bin * 32 + lane : bin], 1u);
            }
        }
    }
    __syncthreads();
    uint32_t *out = p.cta_hist + (uint64_t)blockIdx.x * p.nbins;
    if (spread) {
        for (uint32_t b = threadIdx.x; b < p.nbins; b += kBinThreads) {
            uint32_t sum = 0;
#pragma unroll
            for (uint32_t l = 0; l < 32; ++l) sum += sh_hist[b * 32 + ((l + b) & 31)];
            out[b] = sum;
        }
    } else {
        for (uint32_t b = threadIdx.x; b < p.nbins; b += kBinThreads) out[b] = sh_hist[b];
    }
}

// bin_count[b] = sum over the CTAs' histograms (one thread per bin, four independent loads in flight).
__global__ void scatter_offsets_kernel(const uint32_t *cta_hist, uint32_t *bin_start, uint32_t nbins, uint32_t nctas) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbins) return;
    uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    uint32_t c = 0;
    for (; c + 4 <= nctas; c += 4) {
        s0 += cta_hist[(uint64_t)c * nbins + b];
        s1 += cta_hist[(uint64_t)(c + 1) * nbins + b];
        s2 += cta_hist[(uint64_t)(c + 2) * nbins + b];
        s3 += cta_hist[(uint64_t)(c + 3) * nbins + b];
    }
    for (; c < nctas; ++c) s0 += cta_hist[(uint64_t)c * nbins + b];
    bin_start[b] = s0 + s1 + s2 + s3;  // counts for now; scatter_scan_kernel turns them into starts
}

// Exclusive scan of the (<= 8192) bin counts in one CTA of 1024 threads; bin_start[nbins] = total.
__global__ void __launch_bounds__(1024) scatter_scan_kernel(uint32_t *bin_start, uint32_t nbins) {
    __shared__ uint32_t part[1024];
    constexpr int PER = kMaxBins / 1024;
    uint32_t v[PER], sum = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        const uint32_t b = threadIdx.x * PER + j;
        v[j] = b < nbins ? bin_start[b] : 0;
        sum += v[j];
    }
    part[threadIdx.x] = sum;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        const uint32_t add = threadIdx.x >= (unsigned)off ? part[threadIdx.x - off] : 0;
        __syncthreads();
        part[threadIdx.x] += add;
        __syncthreads();
    }
    uint32_t run = part[threadIdx.x] - sum;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        const uint32_t b = threadIdx.x * PER + j;
        if (b < nbins) bin_start[b] = run;
        run += v[j];
    }
    if (threadIdx.x == 1023) bin_start[nbins] = part[1023];
}

// Staged partition of one level. A tile of kTile elements is ranked by bin INSIDE shared memory and copied out bin
// by bin, so that consecutive threads write consecutive addresses of each bin's run (whole sectors), instead of every
// lane of a store hitting a different run (one 8-byte sector write per lane: measured 2.2 ms per level). Space in the
// output is reserved per (tile, bin) with one global atomic on the bin's cursor.
//   PAIRS = false (level 1): tiles walk the source through the index tensors; the digit is lin >> bin_log.
//   PAIRS = true  (level 2): tiles walk the level-1 output INSIDE one coarse bin (they never straddle two), the digit is
//                            the fine bin within that coarse bin (kFinePerCoarse of them).
//   FINAL: this level's bins are the fine bins: the index inside the bin (u16) is written; otherwise the full index.
constexpr int kTile = kBinThreads * 8;
constexpr int kFinePerCoarseLog = 7, kFinePerCoarse = 1 << kFinePerCoarseLog;
constexpr int kMaxRadix = 128;

template <class TA>
struct PartitionSmem {
    uint32_t hist[kMaxRadix], scan[kMaxRadix], fill[kMaxRadix], gbase[kMaxRadix];
    uint32_t tile_prefix[65];   // level 2: tiles before coarse bin b
    uint32_t key[kTile];
    uint8_t digit[kTile];
    TA val[kTile];
};

template <bool PAIRS, bool FINAL, class TS, class TA, int NDI, int NDO>
__global__ void __launch_bounds__(kBinThreads, 2) scatter_partition_kernel(const __grid_constant__ BinParams p) {
    extern __shared__ __align__(16) unsigned char sh_part_raw[];
    PartitionSmem<TA> &sm = *reinterpret_cast<PartitionSmem<TA> *>(sh_part_raw);
    const int tid = threadIdx.x;
    uint32_t ntiles;
    if constexpr (PAIRS) {
        // tiles per coarse bin from the level-1 bin starts (coarse_start has ncoarse + 1 entries)
        if (tid == 0) {
            uint32_t run = 0;
            for (uint32_t b = 0; b < p.ncoarse; ++b) {
                sm.tile_prefix[b] = run;
                run += (p.coarse_start[b + 1] - p.coarse_start[b] + kTile - 1) / kTile;
            }
            sm.tile_prefix[p.ncoarse] = run;
        }
        __syncthreads();
        ntiles = sm.tile_prefix[p.ncoarse];
    } else {
        ntiles = (p.n + kTile - 1) / kTile;
    }
    const uint32_t radix = PAIRS ? (uint32_t)kFinePerCoarse : p.nbins;
    const TA *in_vals = reinterpret_cast<const TA *>(p.in_vals);
    TA *out_vals = reinterpret_cast<TA *>(p.out_vals);
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        uint32_t begin, end, bin0 = 0;
        if constexpr (PAIRS) {
            uint32_t cb = 0;  // coarse bin of this tile: last b with tile_prefix[b] <= tile (<= 64 entries)
            for (uint32_t step = 32; step >= 1; step >>= 1)
                if (cb + step < p.ncoarse && sm.tile_prefix[cb + step] <= tile) cb += step;
            begin = p.coarse_start[cb] + (tile - sm.tile_prefix[cb]) * kTile;
            end = begin + kTile < p.coarse_start[cb + 1] ? begin + kTile : p.coarse_start[cb + 1];
            bin0 = cb << kFinePerCoarseLog;
        } else {
            begin = tile * kTile;
            end = begin + kTile < p.n ? begin + kTile : p.n;
        }
        if (tid < kMaxRadix) { sm.hist[tid] = 0; sm.fill[tid] = 0; }
        __syncthreads();
        // 1. load the tile, count the digits
        uint32_t lin[8];
        TA v[8];
        int dg[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t f = begin + (uint32_t)k * kBinThreads + tid;
            dg[k] = -1;
            if (f < end) {
                if constexpr (PAIRS) {
                    lin[k] = p.in_lin[f];
                    v[k] = in_vals[f];
                    dg[k] = (int)((lin[k] >> p.bin_log) - bin0);
                } else {
                    int64_t so;
                    if (bin_target<NDI, NDO>(p, f, so, lin[k])) {
                        v[k] = (TA)*reinterpret_cast<const TS *>(p.gs.it_ptr + so);
                    } else {  // out-of-range index: raise the flag, travel on as "add 0 to element 0"
                        atomicExch(p.gs.err, 1);
                        lin[k] = 0;
                        v[k] = TA(0);
                    }
                    dg[k] = (int)(lin[k] >> p.bin_log);
                }
                atomicAdd(&sm.hist[dg[k]], 1u);
            }
        }
        __syncthreads();
        // 2. exclusive scan of the digit counts (warp 0, 4 digits per lane) and space reservation per (tile, bin)
        if (tid < 32) {
            uint32_t c[4], sum = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) { c[j] = (uint32_t)(tid * 4 + j) < radix ? sm.hist[tid * 4 + j] : 0; sum += c[j]; }
            uint32_t incl = sum;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, incl, off);
                if (tid >= off) incl += o;
            }
            uint32_t run = incl - sum;
#pragma unroll
            for (int j = 0; j < 4; ++j) { sm.scan[tid * 4 + j] = run; run += c[j]; }
        }
        if ((uint32_t)tid < radix && sm.hist[tid] > 0) sm.gbase[tid] = atomicAdd(&p.cursor[bin0 + tid], sm.hist[tid]);
        __syncthreads();
        // 3. rank inside the tile, stage bin by bin
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (dg[k] >= 0) {
                const uint32_t pos = sm.scan[dg[k]] + atomicAdd(&sm.fill[dg[k]], 1u);
                sm.key[pos] = lin[k];
                sm.digit[pos] = (uint8_t)dg[k];
                sm.val[pos] = v[k];
            }
        __syncthreads();
        // 4. copy out: consecutive threads -> consecutive slots of a bin's run
        const uint32_t count = end - begin;
        const uint32_t mask = (1u << p.bin_log) - 1;
        for (uint32_t i = tid; i < count; i += kBinThreads) {
            const uint32_t d = sm.digit[i];
            const uint32_t gpos = sm.gbase[d] + (i - sm.scan[d]);
            if constexpr (FINAL) p.out_low[gpos] = (uint16_t)(sm.key[i] & mask);
            else p.out_lin[gpos] = sm.key[i];
            out_vals[gpos] = sm.val[i];
        }
        __syncthreads();
    }
}

// Shared-memory add. 64-bit integers: two 32-bit adds, the carry of the low word (known from the value the low add
// returns) rides on the high add — additions commute, so the final words are right whatever the interleaving; the
// native 64-bit form compiles to a compare-and-swap loop (ATOMS.CAST.SPIN.64).
template <class T> __device__ __forceinline__ void smem_add(T *addr, T v) {
    if constexpr (sizeof(T) == 8 && std::is_integral<T>::value) {
        uint32_t *w = reinterpret_cast<uint32_t *>(addr);
        const uint64_t u = (uint64_t)v;
        const uint32_t lo = (uint32_t)u, hi = (uint32_t)(u >> 32);
        const uint32_t old = atomicAdd(w, lo);
        const uint32_t carry = (uint32_t)(old + lo < old);
        if (hi + carry != 0 || (hi == 0xffffffffu && carry)) atomicAdd(w + 1, hi + carry);
    } else {
        atomicAdd(addr, v);
    }
}

template <class TA>
__global__ void __launch_bounds__(kBinThreads) scatter_accumulate_kernel(const __grid_constant__ BinParams p) {
    extern __shared__ __align__(16) unsigned char sh_raw[];
    TA *acc = reinterpret_cast<TA *>(sh_raw);
    const uint32_t bin_elems = 1u << p.bin_log;
    const TA *vals = reinterpret_cast<const TA *>(p.out_vals);
    TA *target = reinterpret_cast<TA *>(p.gs.other_ptr);
    for (uint32_t b = blockIdx.x; b < p.nbins; b += gridDim.x) {
        for (uint32_t i = threadIdx.x; i < bin_elems; i += kBinThreads) acc[i] = TA(0);
        __syncthreads();
        const uint32_t lo = p.bin_start[b], hi = p.bin_start[b + 1];
        constexpr int U = 4;
        for (uint32_t base = lo; base < hi; base += kBinThreads * U) {
            uint16_t l[U];
            TA v[U];
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const uint32_t i = base + j * kBinThreads + threadIdx.x;
                if (i < hi) { l[j] = p.out_low[i]; v[j] = vals[i]; }
            }
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const uint32_t i = base + j * kBinThreads + threadIdx.x;
                if (i < hi) smem_add(&acc[l[j]], v[j]);
            }
        }
        __syncthreads();
        const uint64_t first = (uint64_t)b << p.bin_log;
        const uint64_t left = p.nt - first;
        const uint32_t cnt = left < bin_elems ? (uint32_t)left : bin_elems;
        for (uint32_t i = threadIdx.x; i < cnt; i += kBinThreads) target[first + i] = acc[i];
        __syncthreads();
    }
}

bool dense_row_major(const dn_tensor *t) {
    int64_t expect = 1;
    for (int d = t->ndims - 1; d >= 0; --d) {
        if (t->shape[d] != 1 && t->stride[d] != expect) return false;
        expect *= t->shape[d];
    }
    return true;
}

// hist -> offsets -> scan -> (cursors = bin starts) -> staged partition for one level.
template <bool PAIRS, bool FINAL, class TS, class TA>
dn_status scatter_partition_level(BinParams &p, int grid) {
    const size_t hist_smem = (size_t)(p.nbins <= kSpreadBins ? p.nbins * 32 : p.nbins) * 4;
    DN_GS_RANK_DISPATCH(p.gs.nd_it, p.gs.nd_other,
                        DN_LAUNCH((scatter_hist_kernel<PAIRS, NDI, NDO>), grid, kBinThreads, hist_smem, p));
    DN_LAUNCH(scatter_offsets_kernel, (unsigned)((p.nbins + 255) / 256), 256, 0, p.cta_hist, p.bin_start, p.nbins, (uint32_t)grid);
    DN_LAUNCH(scatter_scan_kernel, 1, 1024, 0, p.bin_start, p.nbins);
    DN_CUDA_TRY(cudaMemcpyAsync(p.cursor, p.bin_start, (size_t)p.nbins * 4, cudaMemcpyDeviceToDevice, current_stream()));
    const int part_smem = (int)sizeof(PartitionSmem<TA>);
    // (the attribute is per kernel instantiation and device: set before every launch, it is a host-side table write)
    DN_GS_RANK_DISPATCH(p.gs.nd_it, p.gs.nd_other,
                        cudaFuncSetAttribute(scatter_partition_kernel<PAIRS, FINAL, TS, TA, NDI, NDO>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, part_smem));
    const int pgrid = sm_count() * 2;
    DN_GS_RANK_DISPATCH(p.gs.nd_it, p.gs.nd_other,
                        DN_LAUNCH((scatter_partition_kernel<PAIRS, FINAL, TS, TA, NDI, NDO>), pgrid, kBinThreads, part_smem, p));
    return launch_status("scatter partition kernels");
}

// Returns DN_OK with *done = false when the problem does not qualify (the caller runs the atomic kernel).
template <class TS, class TA>
dn_status scatter_binned(const GSParams &gs, const dn_tensor *acc, int64_t nt, bool *done) {
    *done = false;
    const int esz = (int)sizeof(TA);
    const uint32_t fine_log = esz == 8 ? 13 : 14;   // 64 KiB of accumulators per CTA
    const int64_t nfine = (nt + (1ll << fine_log) - 1) >> fine_log;
    // Taken for large problems (2^26 random int64 adds: 1.8 ms against 2.9 ms for the atomic kernel, hot spots 2.2 ms
    // against 5.8 ms). DN_SCATTER_BINNED (test hook): 1 = at any size (the parity tests cover the path on their small
    // cases), 0 = never.
    static const int mode = [] { const char *e = getenv("DN_SCATTER_BINNED"); return e ? atoi(e) : 2; }();
    if (mode <= 0 || nfine > kMaxBins || !dense_row_major(acc) || gs.n == 0 || nt == 0) return DN_OK;
    if (mode != 1 && (gs.n < (1u << 22) || nt < (1ll << 16))) return DN_OK;
    // coarse bins: kFinePerCoarse fine bins each (at most 64 of them); a target of at most kMaxRadix fine bins needs
    // one level only
    const bool two_level = nfine > kMaxRadix;
    const uint32_t coarse_log = fine_log + kFinePerCoarseLog;
    const int64_t ncoarse = (nt + (1ll << coarse_log) - 1) >> coarse_log;
    BinParams p;
    p.gs = gs;
    p.n = gs.n;
    p.elem_log = esz == 8 ? 3 : 2;
    p.nt = (uint64_t)nt;
    const int nctas = sm_count() * 2;
    uint32_t chunk = (uint32_t)(((uint64_t)gs.n + nctas - 1) / nctas);
    chunk = (chunk + kBinThreads * 8 - 1) / (kBinThreads * 8) * (kBinThreads * 8);
    p.chunk = chunk;
    const int grid = (int)(((uint64_t)gs.n + chunk - 1) / chunk);
    void *s_hist = nullptr, *s_start = nullptr, *s_start1 = nullptr, *s_cursor = nullptr, *s_lin = nullptr, *s_vals1 = nullptr,
         *s_low = nullptr, *s_vals2 = nullptr;
    dn_status st = scratch_alloc((size_t)grid * nfine * 4, &s_hist);
    if (st == DN_OK) st = scratch_alloc((size_t)(nfine + 1) * 4, &s_start);
    if (st == DN_OK) st = scratch_alloc((size_t)(ncoarse + 1) * 4, &s_start1);
    if (st == DN_OK) st = scratch_alloc((size_t)(nfine + 1) * 4, &s_cursor);
    if (st == DN_OK && two_level) st = scratch_alloc((size_t)gs.n * 4, &s_lin);
    if (st == DN_OK && two_level) st = scratch_alloc((size_t)gs.n * esz, &s_vals1);
    if (st == DN_OK) st = scratch_alloc((size_t)gs.n * 2, &s_low);
    if (st == DN_OK) st = scratch_alloc((size_t)gs.n * esz, &s_vals2);
    if (st == DN_OK) {
        p.cta_hist = static_cast<uint32_t *>(s_hist);
        p.cursor = static_cast<uint32_t *>(s_cursor);
        p.in_lin = nullptr;
        p.in_vals = nullptr;
        p.out_lin = static_cast<uint32_t *>(s_lin);
        p.out_low = static_cast<uint16_t *>(s_low);
        p.coarse_start = nullptr;
        p.ncoarse = 0;
        static std::atomic<bool> configured[64];
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
            cudaFuncSetAttribute(scatter_accumulate_kernel<TA>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBinSmemBytes);
            if (dev >= 0 && dev < 64) configured[dev].store(true, std::memory_order_release);
        }
        if (two_level) {
            p.nbins = (uint32_t)ncoarse;
            p.bin_log = coarse_log;
            p.bin_start = static_cast<uint32_t *>(s_start1);
            p.out_vals = static_cast<char *>(s_vals1);
            st = scatter_partition_level<false, false, TS, TA>(p, grid);
            p.in_lin = p.out_lin;
            p.in_vals = p.out_vals;
            p.coarse_start = static_cast<const uint32_t *>(s_start1);
            p.ncoarse = (uint32_t)ncoarse;
        }
        if (st == DN_OK) {
            p.nbins = (uint32_t)nfine;
            p.bin_log = fine_log;
            p.bin_start = static_cast<uint32_t *>(s_start);
            p.out_vals = static_cast<char *>(s_vals2);
            st = two_level ? scatter_partition_level<true, true, TS, TA>(p, grid)
                           : scatter_partition_level<false, true, TS, TA>(p, grid);
        }
        if (st == DN_OK) {
            const int acc_grid = nfine < (int64_t)sm_count() * 3 ? (int)nfine : sm_count() * 3;
            DN_LAUNCH((scatter_accumulate_kernel<TA>), acc_grid, kBinThreads, kBinSmemBytes, p);
            st = launch_status("scatter accumulate kernel");
        }
        *done = st == DN_OK;
    }
    scratch_free(s_hist);
    scratch_free(s_start);
    scratch_free(s_start1);
    scratch_free(s_cursor);
    scratch_free(s_lin);
    scratch_free(s_vals1);
    scratch_free(s_low);
    scratch_free(s_vals2);
    return st;
}

dn_status gs_fill(GSParams &p, const dn_tensor *walked, const dn_tensor *other, const dn_tensor *const *idxs,
                  const char *what) {
    const int64_t n = num_elements(walked);
    if (n >= ((int64_t)1 << 31)) return set_error(DN_ERR_UNSUPPORTED, "%s: more than 2^31-1 elements", what);
    p.n = (uint32_t)n;
    p.it_ptr = data_ptr(walked);
    p.other_ptr = data_ptr(other);
    p.nd_it = walked->ndims;
    p.nd_other = other->ndims;
    const int wsz = dtype_size(walked->dtype), osz = dtype_size(other->dtype);
    for (int k = 0; k < DN_MAX_DIMS; ++k) {
        const int d = walked->ndims - 1 - k;
        p.it_shape[k] = d >= 0 ? (uint32_t)walked->shape[d] : 1;
        p.it_div[k].init(p.it_shape[k]);
        p.it_stride[k] = d >= 0 ? walked->stride[d] * wsz : 0;
    }
    for (int d = 0; d < DN_MAX_DIMS; ++d) {
        const bool on = d < other->ndims;
        p.other_shape[d] = on ? other->shape[d] : 1;
        p.other_stride[d] = on ? other->stride[d] * osz : 0;
        p.idx_ptr[d] = (on && idxs[d]) ? data_ptr(idxs[d]) : nullptr;
        for (int k = 0; k < DN_MAX_DIMS; ++k) {
            const int wd = walked->ndims - 1 - k;
            p.idx_stride[d][k] = (on && idxs[d] && wd >= 0) ? idxs[d]->stride[wd] * 8 : 0;
        }
    }
    p.err = index_error_flag();
    if (!p.err) return set_error(DN_ERR_NO_DEVICE, "%s: device not initialised (call dn_init)", what);
    return DN_OK;
}

dn_status check_index_error(const char *what) {
    if (!check_errors_enabled()) return DN_OK;
    int h = 0;
    DN_CUDA_TRY(cudaMemcpyAsync(&h, index_error_flag(), sizeof(int), cudaMemcpyDeviceToHost, current_stream()));
    DN_CUDA_TRY(cudaStreamSynchronize(current_stream()));
    if (h) {
        DN_CUDA_TRY(cudaMemsetAsync(index_error_flag(), 0, sizeof(int), current_stream()));
        return set_error(DN_ERR_INDEX_OUT_OF_RANGE, "invalid index during gather or scatter (%s)", what);
    }
    return DN_OK;
}

dn_status validate_gs(const dn_tensor *walked, const dn_tensor *other, const dn_tensor *const *idxs, int nidxs,
                      const char *what) {
    if (!tensor_valid(walked) || !tensor_valid(other) || !idxs)
        return set_error(DN_ERR_INVALID_ARG, "%s: bad argument", what);
    if (walked->dtype != other->dtype) return set_error(DN_ERR_INVALID_ARG, "%s: source and target types differ", what);
    if (nidxs != other->ndims)
        return set_error(DN_ERR_INVALID_ARG, "%s: one index tensor (or None) per indexed dimension is required", what);
    for (int d = 0; d < nidxs; ++d) {
        if (idxs[d]) {
            if (!tensor_valid(idxs[d]) || idxs[d]->dtype != DN_I64 || !same_shape(idxs[d], walked))
                return set_error(DN_ERR_INVALID_ARG, "%s: index tensors must be int64 and have the walked tensor's shape", what);
        } else if (d >= walked->ndims) {
            return set_error(DN_ERR_INVALID_ARG, "%s: index dimensions beyond the walked tensor's rank must not be None", what);
        }
    }
    return DN_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Ordered compaction
// ---------------------------------------------------------------------------------------------------------------
constexpr int kItems = 16;                          // consecutive logical positions per thread
constexpr int kCompactThreads = 256;
constexpr int kTileElems = kCompactThreads * kItems * 2;  // 8192 per tile (two chunks per thread)

struct BoolView {           // a bool tensor walked in logical row-major order
    const char *ptr;
    int32_t nd;
    uint32_t n;
    uint32_t shape[DN_MAX_DIMS];   // innermost-first
    FastDiv div[DN_MAX_DIMS];
    int64_t stride[DN_MAX_DIMS];   // bytes
};

// Bit i of the result = byte i of the 16-byte vector is non-zero. Per word: 0x01 per non-zero byte, then the
// carry-free multiply 0x01020408 moves byte i's flag to bit 24+i.
__device__ __forceinline__ uint32_t bool16_to_bits(const uint4 &w) {
    const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
    uint32_t bits = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t x = ws[i];
        const uint32_t nz = ((((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) >> 7) & 0x01010101u;
        bits |= ((nz * 0x01020408u) >> 24) << (4 * i);
    }
    return bits;
}

// Loads the (up to 16) mask bytes of logical positions [f0, f0+16) into bits of a 16-bit word.
__device__ __forceinline__ uint32_t load_mask_bits(const BoolView &m, uint32_t f0) {
    if (f0 >= m.n) return 0;
    uint32_t pos[DN_MAX_DIMS];
    uint32_t rem = f0;
    int64_t off = 0;
#pragma unroll
    for (int k = 0; k < DN_MAX_DIMS; ++k) {
        if (k >= m.nd) break;
        uint32_t q, x;
        if (k == m.nd - 1) { x = rem; q = 0; }
        else { q = m.div[k].div(rem); x = rem - q * m.shape[k]; }
        pos[k] = x;
        off += (int64_t)x * m.stride[k];
        rem = q;
    }
    const uint32_t cnt = (m.n - f0 < (uint32_t)kItems) ? m.n - f0 : (uint32_t)kItems;
    const char *a = m.ptr + off;
    uint32_t bits = 0;
    if (cnt == kItems && m.stride[0] == 1 && pos[0] + kItems <= m.shape[0] && (reinterpret_cast<uintptr_t>(a) & 15) == 0) {
        return bool16_to_bits(*reinterpret_cast<const uint4 *>(a));
    }
    // generic: odometer walk
    for (uint32_t j = 0; j < cnt; ++j) {
        if (*reinterpret_cast<const uint8_t *>(m.ptr + off)) bits |= 1u << j;
        // advance one logical position
#pragma unroll
        for (int k = 0; k < DN_MAX_DIMS; ++k) {
            if (k >= m.nd) break;
            off += m.stride[k];
            if (++pos[k] < m.shape[k] || k == m.nd - 1) break;
            off -= (int64_t)m.shape[k] * m.stride[k];
            pos[k] = 0;
        }
    }
    return bits;
}

// dn_count_true: several 128-bit loads in flight per thread, one block reduction at the very end.
__global__ void __launch_bounds__(kIdxThreads) mask_count_kernel(const __grid_constant__ BoolView m, unsigned long long *total) {
    unsigned long long acc = 0;
    const uint64_t nchunks = ((uint64_t)m.n + kItems - 1) / kItems;
    for (uint64_t c = (uint64_t)blockIdx.x * kIdxThreads + threadIdx.x; c < nchunks; c += (uint64_t)gridDim.x * kIdxThreads * 4) {
        uint32_t b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint64_t cc = c + (uint64_t)u * gridDim.x * kIdxThreads;
            b[u] = cc < nchunks ? load_mask_bits(m, (uint32_t)(cc * kItems)) : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) acc += __popc(b[u]);
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    __shared__ unsigned long long wsum[kIdxThreads / 32];
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
#pragma unroll
        for (int w = 0; w < kIdxThreads / 32; ++w) t += wsum[w];
        if (t) atomicAdd(total, t);
    }
}

// Sinks: `load(rank, f)` fetches what the element needs, `store(rank, f, v)` writes it; the kernel only calls
// them for ranks below `cap` (rows / values available on the dense side) ------------------------------------------
struct CoordSink {
    using Value = int;
    static constexpr bool kLoadNeedsRank = false;
    char *t;
    int64_t ts0, ts1;  // bytes
    int64_t cap;       // rows available in t
    int32_t nd;
    uint32_t shape[DN_MAX_DIMS];  // innermost-first (the ORIGINAL dims of the view)
    FastDiv div[DN_MAX_DIMS];
    __device__ __forceinline__ Value load(int64_t, uint32_t) const { return 0; }
    __device__ __forceinline__ void store(int64_t rank, uint32_t f, Value) const {
        uint32_t rem = f;
#pragma unroll
        for (int k = 0; k < DN_MAX_DIMS; ++k) {
            if (k >= nd) break;
            uint32_t q, x;
            if (k == nd - 1) { x = rem; q = 0; }
            else { q = div[k].div(rem); x = rem - q * shape[k]; }
            *reinterpret_cast<int64_t *>(t + rank * ts0 + (int64_t)(nd - 1 - k) * ts1) = (int64_t)x;
            rem = q;
        }
    }
};

// TrueIndices of a matrix mask into a dense, 16-byte aligned [nTrue, 2] target (the common case): one division and
// one 128-bit store per row.
struct CoordSink2D {
    using Value = int;
    static constexpr bool kLoadNeedsRank = false;
    longlong2 *t;
    int64_t cap;
    uint32_t ncols;
    FastDiv div;
    __device__ __forceinline__ Value load(int64_t, uint32_t) const { return 0; }
    __device__ __forceinline__ void store(int64_t rank, uint32_t f, Value) const {
        const uint32_t q = div.div(f);
        __stcs(t + rank, make_longlong2((long long)q, (long long)(f - q * ncols)));
    }
};

struct IndexListSink {  // sel[rank] = f
    using Value = int;
    static constexpr bool kLoadNeedsRank = false;
    int64_t *sel;
    int64_t cap;
    __device__ __forceinline__ Value load(int64_t, uint32_t) const { return 0; }
    __device__ __forceinline__ void store(int64_t rank, uint32_t f, Value) const { sel[rank] = (int64_t)f; }
};

template <class B>
struct GetSink {  // MaskedGet 1-D: t[rank] = a[f]
    using Value = B;
    static constexpr bool kLoadNeedsRank = false;  // the source element is known before the rank is
    char *t;
    const char *a;
    int64_t ts, as;  // bytes
    int64_t cap;
    __device__ __forceinline__ Value load(int64_t, uint32_t f) const {
        return *reinterpret_cast<const B *>(a + (int64_t)f * as);
    }
    __device__ __forceinline__ void store(int64_t rank, uint32_t, Value v) const {
        *reinterpret_cast<B *>(t + rank * ts) = v;
    }
};

template <class B>
struct SetSink {  // MaskedSet 1-D: t[f] = a[rank]
    using Value = B;
    static constexpr bool kLoadNeedsRank = true;
    char *t;
    const char *a;
    int64_t ts, as;
    int64_t cap;  // values available in a
    __device__ __forceinline__ Value load(int64_t rank, uint32_t) const {
        return *reinterpret_cast<const B *>(a + rank * as);
    }
    __device__ __forceinline__ void store(int64_t, uint32_t f, Value v) const {
        *reinterpret_cast<B *>(t + (int64_t)f * ts) = v;
    }
};

// Tile status word of the decoupled look-back: bits 63..62 = status, bits 61..0 = count.
constexpr unsigned long long kStAggregate = 1ull << 62;  // count = true elements of this tile only
constexpr unsigned long long kStPrefix = 2ull << 62;     // count = true elements of tiles 0..this (inclusive)
constexpr unsigned long long kStValueMask = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long ld_state(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_state(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Exclusive prefix (number of true elements in all earlier tiles) of `tile` by decoupled look-back; called by ONE
// full warp after the tile's aggregate was published. Publishes the tile's inclusive prefix. The chain "prefix
// published -> visible to the tiles behind" advances one window per L2 round trip, so the window is 32 * kLook
// predecessors wide.
__device__ __forceinline__ unsigned long long warp_look_back(unsigned long long *state, uint32_t tile, uint32_t tile_total,
                                                             int lane) {
    constexpr int kLook = 4;
    unsigned long long excl = 0;
    int64_t look = (int64_t)tile - 1;  // nearest predecessor not yet accounted for
    while (look >= 0) {
        unsigned long long v[kLook];
        int stop;         // first k whose state ends this lane's walk (a prefix, or not yet published)
        bool stop_ready;  // ... and that state is a prefix
        uint32_t stops;
        do {
#pragma unroll
            for (int k = 0; k < kLook; ++k) {
                const int64_t idx = look - (lane * kLook + k);
                v[k] = idx >= 0 ? ld_state(state + idx) : kStPrefix;
            }
            stop = kLook;
            stop_ready = false;
#pragma unroll
            for (int k = kLook - 1; k >= 0; --k)
                if ((v[k] >> 62) != 1) {
                    stop = k;
                    stop_ready = (v[k] >> 62) == 2;
                }
            stops = __ballot_sync(0xffffffffu, stop < kLook);
            // the nearest stopping state must be a prefix; if it is unpublished, poll again
        } while (stops && !__shfl_sync(0xffffffffu, stop_ready, __ffs(stops) - 1));
        const int first = stops ? __ffs(stops) - 1 : 32;
        unsigned long long contrib = 0;
#pragma unroll
        for (int k = 0; k < kLook; ++k)
            if (lane < first || (lane == first && k <= stop)) contrib += v[k] & kStValueMask;
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, s);
        excl += contrib;
        if (stops) break;
        look -= 32 * kLook;
    }
    if (lane == 0 && tile != 0) st_state(state + tile, kStPrefix | (excl + tile_total));
    return excl;
}

// One CTA per tile; a tile is ROUNDS x 8192 consecutive logical positions (2*ROUNDS 128-bit mask loads per
// thread, all issued up front). state[0..ntiles) and *ticket must be zero on entry. Tiles are handed out in ticket
// order, so every tile a CTA waits on during look-back is owned by a CTA that is already running: no deadlock
// whatever order the hardware dispatches CTAs in. Phases per tile: mask loads -> ranks inside the tile -> publish
// the tile's count -> per round: stage the selected positions (16-bit offsets) in shared memory, issue the first
// group of sink loads, [first round: look-back by warp 0, its latency overlaps the loads in flight], stores.
// ROUNDS = 4 amortises ticket, look-back and their barriers over 32768 positions; ROUNDS = 1 keeps small inputs
// spread over the machine.
constexpr int kRoundElems = kTileElems;  // positions staged at a time

template <class Sink, bool FAST, int ROUNDS>
__global__ void __launch_bounds__(kCompactThreads, ROUNDS == 1 ? 6 : 4) compact_kernel(const __grid_constant__ BoolView m, unsigned long long *state,
                                                                    uint32_t *ticket, const Sink sink) {
    constexpr int kWarps = kCompactThreads / 32;
    constexpr int NCH = 2 * ROUNDS;                     // 4096-position chunks per tile
    constexpr int kChunkElems = kCompactThreads * kItems;
    constexpr int U = 4;
    __shared__ uint32_t warp_sums[NCH][kWarps];
    __shared__ uint32_t s_tile;
    __shared__ unsigned long long s_base;
    __shared__ uint16_t staged[kRoundElems];  // round offsets of the round's true elements, in logical order
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t tile0 = (uint64_t)tile * (kRoundElems * ROUNDS);
    uint32_t bits[NCH];
    if (FAST && tile0 + (uint64_t)kChunkElems * NCH <= m.n) {
        // interior tile of a contiguous mask: all 128-bit loads are issued back to back, then converted
        uint4 w[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c)
            w[c] = __ldcs(reinterpret_cast<const uint4 *>(m.ptr + tile0 + c * kChunkElems + threadIdx.x * kItems));
#pragma unroll
        for (int c = 0; c < NCH; ++c) bits[c] = bool16_to_bits(w[c]);
    } else {
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const uint64_t f0 = tile0 + c * kChunkElems + threadIdx.x * kItems;
            if (FAST && f0 + kItems <= m.n) bits[c] = bool16_to_bits(__ldcs(reinterpret_cast<const uint4 *>(m.ptr + f0)));
            else bits[c] = f0 < m.n ? load_mask_bits(m, (uint32_t)f0) : 0u;
        }
    }
    uint32_t excl_in_warp[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const uint32_t cnt = __popc(bits[c]);
        uint32_t incl = cnt;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, incl, s);
            if (lane >= s) incl += o;
        }
        if (lane == 31) warp_sums[c][warp] = incl;
        excl_in_warp[c] = incl - cnt;
    }
    __syncthreads();
    uint32_t tile_total = 0;
    uint32_t local[NCH];      // rank of this thread's first element of chunk c inside its round
    uint32_t round_total[ROUNDS];
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        uint32_t acc = 0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = 2 * r + h;
            uint32_t before = 0, total = 0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) {
                const uint32_t ws = warp_sums[c][w];
                if (w < warp) before += ws;
                total += ws;
            }
            local[c] = acc + before + excl_in_warp[c];
            acc += total;
        }
        round_total[r] = acc;
        tile_total += acc;
    }
    if (threadIdx.x == 0) st_state(state + tile, (tile == 0 ? kStPrefix : kStAggregate) | tile_total);

    auto look_back = [&]() {
        if (warp == 0) {
            const unsigned long long excl = warp_look_back(state, tile, tile_total, lane);
            if (lane == 0) s_base = excl;
        }
        __syncthreads();
        return (int64_t)s_base;
    };

    const uint32_t staged_base = (uint32_t)__cvta_generic_to_shared(staged);
    int64_t base = 0;
    bool have_base = false;
    if constexpr (Sink::kLoadNeedsRank) {
        base = look_back();
        have_base = true;
    }
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        const uint32_t round0 = (uint32_t)tile0 + (uint32_t)r * kRoundElems;
        if (r > 0) __syncthreads();  // the previous round's staged offsets are no longer read
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t b = bits[2 * r + h];
            const uint32_t off0 = h * kChunkElems + threadIdx.x * kItems;
            // predicated, no divergent loop; the slot is a 32-bit shared-window address that is bumped by two — three
            // instructions per position (test, store, add). Indexing `staged[at++]` made the compiler rebuild the
            // generic address for every store (S2R + MOV + LEA): ~8 instructions per position, 55 % of all the
            // kernel issued on a sparse mask (profiles/r02c_compact_trueidx_sparse: 15 instructions per position).
            uint32_t slot = staged_base + local[2 * r + h] * 2;
#pragma unroll
            for (int j = 0; j < kItems; ++j)
                if ((b >> j) & 1u) {
                    asm volatile("st.shared.u16 [%0], %1;" ::"r"(slot), "h"((uint16_t)(off0 + j)) : "memory");
                    slot += 2;
                }
        }
        __syncthreads();
        // Consecutive threads own consecutive ranks: dense-side accesses are fully coalesced; U loads in flight.
        // Index math inside the round is 32-bit; the target capacity is folded into the round's limit once.
        const uint32_t rt = round_total[r];
        for (uint32_t i0 = 0; i0 < rt || (i0 == 0 && !have_base); i0 += kCompactThreads * U) {
            uint32_t f[U];
            typename Sink::Value v[U];
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const uint32_t i = i0 + j * kCompactThreads + threadIdx.x;
                f[j] = i < rt ? round0 + staged[i] : 0xffffffffu;
            }
            if constexpr (!Sink::kLoadNeedsRank) {
#pragma unroll
                for (int j = 0; j < U; ++j)
                    if (f[j] != 0xffffffffu) v[j] = sink.load(0, f[j]);
                if (!have_base) {  // uniform; overlaps the loads issued above
                    base = look_back();
                    have_base = true;
                }
            }
            const int64_t room = sink.cap - base;  // ranks of this round that fit the target
            const uint32_t lim = room <= 0 ? 0u : (room < (int64_t)rt ? (uint32_t)room : rt);
            const int64_t rank0 = base + threadIdx.x;
            if constexpr (Sink::kLoadNeedsRank) {
#pragma unroll
                for (int j = 0; j < U; ++j)
                    if (i0 + j * kCompactThreads + threadIdx.x < lim) v[j] = sink.load(rank0 + (i0 + j * kCompactThreads), f[j]);
            }
#pragma unroll
            for (int j = 0; j < U; ++j)
                if (i0 + j * kCompactThreads + threadIdx.x < lim) sink.store(rank0 + (i0 + j * kCompactThreads), f[j], v[j]);
        }
        base += rt;
    }
}

dn_status make_bool_view(BoolView &v, const dn_tensor *a, const char *what) {
    const int64_t n = num_elements(a);
    if (n >= ((int64_t)1 << 31)) return set_error(DN_ERR_UNSUPPORTED, "%s: more than 2^31-1 elements", what);
    v.ptr = data_ptr(a);
    v.nd = a->ndims;
    v.n = (uint32_t)n;
    for (int k = 0; k < DN_MAX_DIMS; ++k) {
        const int d = a->ndims - 1 - k;
        v.shape[k] = d >= 0 ? (uint32_t)a->shape[d] : 1;
        v.div[k].init(v.shape[k]);
        v.stride[k] = d >= 0 ? a->stride[d] : 0;
    }
    if (v.nd == 0) {  // rank 0: one element
        v.nd = 1;
        v.shape[0] = 1;
        v.div[0].init(1);
        v.stride[0] = 0;
    }
    return DN_OK;
}

// Merges adjacent dims that are contiguous in row-major order (and drops size-1 dims); the logical walk order is
// unchanged, the index math gets shorter and contiguous masks become 1-D (one 128-bit load per 16 positions).
void merge_bool_view(BoolView &v) {
    uint32_t shape[DN_MAX_DIMS];
    int64_t stride[DN_MAX_DIMS];
    int m = 0;
    for (int k = 0; k < v.nd; ++k) {  // innermost-first
        if (v.shape[k] == 1) continue;
        if (m > 0 && v.stride[k] == stride[m - 1] * (int64_t)shape[m - 1] &&
            (uint64_t)shape[m - 1] * v.shape[k] < (1ull << 31)) {
            shape[m - 1] *= v.shape[k];
            continue;
        }
        shape[m] = v.shape[k];
        stride[m] = v.stride[k];
        ++m;
    }
    if (m == 0) { shape[0] = 1; stride[0] = 0; m = 1; }
    v.nd = m;
    for (int k = 0; k < DN_MAX_DIMS; ++k) {
        v.shape[k] = k < m ? shape[k] : 1;
        v.stride[k] = k < m ? stride[k] : 0;
        v.div[k].init(v.shape[k]);
    }
}

int tiles_grid(uint32_t ntiles, int per_sm) {
    const int64_t cap = (int64_t)sm_count() * per_sm;
    return (int)(ntiles < cap ? (ntiles ? ntiles : 1) : cap);
}

// Single-pass ordered compaction of one bool view into `sink`.
template <class Sink>
dn_status run_compaction(const BoolView &m, const Sink &sink) {
    if (m.n == 0) return DN_OK;
    // big inputs: 32768 positions per CTA; small ones: 8192, so that they still spread over the machine
    // (DN_COMPACT_ROUNDS = 1 | 4 forces either configuration: test hook so that both are covered at small sizes)
    static const int forced = [] { const char *e = getenv("DN_COMPACT_ROUNDS"); return e ? atoi(e) : 0; }();
    const int rounds = forced == 1 || forced == 4 ? forced
                                                  : ((uint64_t)m.n >= (uint64_t)sm_count() * 12 * kTileElems * 4 ? 4 : 1);
    const uint32_t ntiles = (uint32_t)(((uint64_t)m.n + (uint64_t)kTileElems * rounds - 1) / ((uint64_t)kTileElems * rounds));
    void *scratch = nullptr;
    const size_t nbytes = ((size_t)ntiles + 1) * sizeof(unsigned long long);
    dn_status st = scratch_alloc(nbytes, &scratch);
    if (st != DN_OK) return st;
    cudaError_t e = cudaMemsetAsync(scratch, 0, nbytes, current_stream());
    if (e != cudaSuccess) {
        scratch_free(scratch);
        return cuda_error(e, "compaction scratch");
    }
    unsigned long long *state = reinterpret_cast<unsigned long long *>(scratch);
    uint32_t *ticket = reinterpret_cast<uint32_t *>(state + ntiles);
    const bool fast = m.nd == 1 && m.stride[0] == 1 && (reinterpret_cast<uintptr_t>(m.ptr) & 15) == 0;

    if (fast && rounds == 4) DN_LAUNCH((compact_kernel<Sink, true, 4>), ntiles, kCompactThreads, 0, m, state, ticket, sink);
    else if (fast) DN_LAUNCH((compact_kernel<Sink, true, 1>), ntiles, kCompactThreads, 0, m, state, ticket, sink);
    else if (rounds == 4) DN_LAUNCH((compact_kernel<Sink, false, 4>), ntiles, kCompactThreads, 0, m, state, ticket, sink);
    else DN_LAUNCH((compact_kernel<Sink, false, 1>), ntiles, kCompactThreads, 0, m, state, ticket, sink);
    scratch_free(scratch);
    return launch_status("compaction kernel");
}

// Separable gather / scatter through per-dimension index lists (general MaskedGet / MaskedSet).
struct SepParams {
    char *dense;          // the dense side: MaskedGet target / MaskedSet values
    char *full;           // the full side: MaskedGet source / MaskedSet target
    int32_t nd;
    uint32_t n;           // elements of the dense side
    uint32_t dshape[DN_MAX_DIMS];   // dense shape, innermost-first
    FastDiv ddiv[DN_MAX_DIMS];
    int64_t dstride[DN_MAX_DIMS];   // bytes
    int64_t fstride[DN_MAX_DIMS];   // bytes
    const int64_t *sel[DN_MAX_DIMS];  // per dim index list or nullptr (identity)
};

template <class B, bool IsGet>
__global__ void __launch_bounds__(kIdxThreads) separable_kernel(const __grid_constant__ SepParams p) {
    for (uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; f < p.n; f += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t rem = (uint32_t)f;
        int64_t doff = 0, foff = 0;
        bool selected = true;   // false: a dense position beyond the number of selected elements (list entry -1)
#pragma unroll
        for (int k = 0; k < DN_MAX_DIMS; ++k) {
            if (k >= p.nd) break;
            uint32_t q, x;
            if (k == p.nd - 1) { x = rem; q = 0; }
            else { q = p.ddiv[k].div(rem); x = rem - q * p.dshape[k]; }
            doff += (int64_t)x * p.dstride[k];
            const int64_t fx = p.sel[k] ? p.sel[k][x] : (int64_t)x;
            selected = selected && fx >= 0;
            foff += fx * p.fstride[k];
            rem = q;
        }
        if (!selected) continue;
        if (IsGet) *reinterpret_cast<B *>(p.dense + doff) = *reinterpret_cast<const B *>(p.full + foff);
        else *reinterpret_cast<B *>(p.full + foff) = *reinterpret_cast<const B *>(p.dense + doff);
    }
}

dn_status check_masks(const dn_tensor *full, const dn_tensor *dense, const dn_tensor *const *masks, int nmasks,
                      const char *what) {
    if (!tensor_valid(full) || !tensor_valid(dense) || !masks)
        return set_error(DN_ERR_INVALID_ARG, "%s: bad argument", what);
    if (full->dtype != dense->dtype || full->ndims != dense->ndims)
        return set_error(DN_ERR_INVALID_ARG, "%s: source and target must have the same type and rank", what);
    if (nmasks != full->ndims)
        return set_error(DN_ERR_INVALID_ARG, "%s: one mask (or NoMask) per dimension is required", what);
    for (int d = 0; d < nmasks; ++d) {
        if (masks[d]) {
            if (!tensor_valid(masks[d]) || masks[d]->dtype != DN_BOOL || masks[d]->ndims != 1 ||
                masks[d]->shape[0] != full->shape[d])
                return set_error(DN_ERR_INVALID_ARG, "%s: mask %d must be a 1-D bool tensor of length %lld", what, d,
                                 (long long)full->shape[d]);
        } else if (dense->shape[d] != full->shape[d]) {
            return set_error(DN_ERR_SHAPE_MISMATCH, "%s: unmasked dimension %d must have equal sizes", what, d);
        }
    }
    return DN_OK;
}

template <bool IsGet>
dn_status masked_general(const dn_tensor *dense, const dn_tensor *full, const dn_tensor *const *masks, const char *what) {
    const int nd = full->ndims;
    const int64_t n = num_elements(dense);
    if (n == 0) return DN_OK;
    if (n >= ((int64_t)1 << 31)) return set_error(DN_ERR_UNSUPPORTED, "%s: more than 2^31-1 elements", what);
    // index lists for the masked dims, in one scratch block
    int64_t total = 0;
    for (int d = 0; d < nd; ++d)
        if (masks[d]) total += dense->shape[d];
    void *scratch = nullptr;
    dn_status st = scratch_alloc((size_t)(total + 1) * sizeof(int64_t), &scratch);
    if (st != DN_OK) return st;
    // the lists start as -1: a dense side larger than the number of selected elements leaves its surplus positions
    // untouched (the host backend walks the selected elements only, ScalarOps.fs:667-707)
    const cudaError_t me = cudaMemsetAsync(scratch, 0xFF, (size_t)(total + 1) * sizeof(int64_t), current_stream());
    if (me != cudaSuccess) {
        scratch_free(scratch);
        return cuda_error(me, what);
    }
    SepParams p;
    int64_t *cursor = reinterpret_cast<int64_t *>(scratch);
    const int sz = dtype_size(full->dtype);
    p.nd = nd;
    p.n = (uint32_t)n;
    p.dense = data_ptr(dense);
    p.full = data_ptr(full);
    for (int k = 0; k < DN_MAX_DIMS; ++k) {
        const int d = nd - 1 - k;
        p.dshape[k] = d >= 0 ? (uint32_t)dense->shape[d] : 1;
        p.ddiv[k].init(p.dshape[k]);
        p.dstride[k] = d >= 0 ? dense->stride[d] * sz : 0;
        p.fstride[k] = d >= 0 ? full->stride[d] * sz : 0;
        p.sel[k] = nullptr;
        if (d >= 0 && masks[d]) {
            BoolView mv;
            st = make_bool_view(mv, masks[d], what);
            if (st == DN_OK) st = run_compaction(mv, IndexListSink{cursor, dense->shape[d]});
            if (st != DN_OK) {
                scratch_free(scratch);
                return st;
            }
            p.sel[k] = cursor;
            cursor += dense->shape[d];
        }
    }
    const int grid = ew_grid_for(n, kIdxThreads);
    DN_SWITCH_SIZE(sz, { DN_LAUNCH((separable_kernel<B, IsGet>), grid, kIdxThreads, 0, p); });
    scratch_free(scratch);
    return launch_status(what);
}

}  // namespace

namespace dn {
// countTrue of a bool view into a device counter, stream-ordered (no read-back). Used by dn_count_true and by the
// sharded compaction (shard.cu), which publishes the count to the peers without a host round trip.
dn_status count_true_async(const dn_tensor *a, unsigned long long *dev_total) {
    if (!tensor_valid(a) || a->dtype != DN_BOOL) return set_error(DN_ERR_INVALID_ARG, "countTrue: bad argument");
    DN_CUDA_TRY(cudaMemsetAsync(dev_total, 0, sizeof(unsigned long long), current_stream()));
    BoolView m;
    dn_status st = make_bool_view(m, a, "countTrue");
    if (st != DN_OK) return st;
    if (m.n == 0) return DN_OK;
    merge_bool_view(m);
    const uint32_t nblocks = (uint32_t)(((uint64_t)m.n + kIdxThreads * kItems - 1) / (kIdxThreads * kItems));
    DN_LAUNCH(mask_count_kernel, tiles_grid(nblocks, 16), kIdxThreads, 0, m, dev_total);
    return launch_status("countTrue kernel");
}
}  // namespace dn

extern "C" {

dn_status dn_gather(const dn_tensor *t, const dn_tensor *const *idxs, int32_t nidxs, const dn_tensor *a) {
    dn_status st = validate_gs(t, a, idxs, nidxs, "Gather");
    if (st != DN_OK) return st;
    if (num_elements(t) == 0) return DN_OK;
    GSParams p;
    st = gs_fill(p, t, a, idxs, "Gather");
    if (st != DN_OK) return st;
    const int grid = ew_grid_for(p.n, kIdxThreads * 4);
    DN_SWITCH_SIZE(dtype_size(t->dtype), {
        DN_GS_RANK_DISPATCH(p.nd_it, p.nd_other, DN_LAUNCH((gather_kernel<B, NDI, NDO>), grid, kIdxThreads, 0, p));
    });
    st = launch_status("gather kernel");
    if (st != DN_OK) return st;
    return check_index_error("Gather");
}

dn_status dn_scatter(const dn_tensor *t, const dn_tensor *const *idxs, int32_t nidxs, const dn_tensor *a) {
    dn_status st = validate_gs(a, t, idxs, nidxs, "Scatter");
    if (st != DN_OK) return st;
    if (t->dtype == DN_BOOL) return set_error(DN_ERR_UNSUPPORTED, "Scatter is not defined for type bool (no addition)");
    const int esize = dtype_size(t->dtype);
    const int64_t nt = num_elements(t);
    // 8/16-bit element types have no native atomic add: accumulate in a dense int32 scratch tensor of the target's
    // shape (sums agree modulo 2^8 / 2^16, i.e. the host's wrapping arithmetic) and narrow into the target.
    dn_tensor acc = *t;
    void *scratch = nullptr;
    if (esize < 4) {
        st = scratch_alloc((size_t)(nt > 0 ? nt : 1) * 4, &scratch);
        if (st != DN_OK) return st;
        acc.base = scratch;
        acc.offset = 0;
        acc.dtype = DN_I32;
        int64_t stride = 1;
        for (int d = t->ndims - 1; d >= 0; --d) {
            acc.stride[d] = stride;
            stride *= t->shape[d];
        }
    }
    // Large dense targets: partition by target bin and accumulate in shared memory (the bins are written whole, which
    // is the zero-fill). Everything else: zero-fill (CudaBackend.fs:379), then one (warp-aggregated) atomic per element.
    GSParams p;
    bool binned = false;
    const bool have_src = num_elements(a) > 0;
    if (have_src) {
        st = gs_fill(p, a, &acc, idxs, "Scatter");
        if (st == DN_OK) {
            switch (t->dtype) {
            case DN_F32: st = scatter_binned<float, float>(p, &acc, nt, &binned); break;
            case DN_F64: st = scatter_binned<double, double>(p, &acc, nt, &binned); break;
            case DN_I8: st = scatter_binned<int8_t, int32_t>(p, &acc, nt, &binned); break;
            case DN_U8: st = scatter_binned<uint8_t, int32_t>(p, &acc, nt, &binned); break;
            case DN_I16: st = scatter_binned<int16_t, int32_t>(p, &acc, nt, &binned); break;
            case DN_U16: st = scatter_binned<uint16_t, int32_t>(p, &acc, nt, &binned); break;
            case DN_I32: st = scatter_binned<int32_t, int32_t>(p, &acc, nt, &binned); break;
            case DN_U32: st = scatter_binned<uint32_t, uint32_t>(p, &acc, nt, &binned); break;
            case DN_I64: st = scatter_binned<int64_t, int64_t>(p, &acc, nt, &binned); break;
            case DN_U64: st = scatter_binned<uint64_t, uint64_t>(p, &acc, nt, &binned); break;
            default: break;
            }
        }
    }
    uint64_t zero = 0;
    if (st == DN_OK && !binned) st = dn_fill_const(&acc, &zero);
    if (st == DN_OK && have_src && !binned) {
        {
            const int grid = ew_grid_for(p.n, kIdxThreads * 4);
            switch (t->dtype) {
            case DN_F32: DN_GS_RANK_DISPATCH(p.nd_it, p.nd_other, DN_LAUNCH((scatter_kernel<float, float, NDI, NDO>), grid, kIdxThreads, 0, p)); break;
            case DN_F64: DN_GS_RANK_DISPATCH(p.nd_it, p.nd_other, DN_LAUNCH((scatter_kernel<double, double, NDI, NDO>), grid, kIdxThreads, 0, p)); break;
            case DN_I8: DN_GS_RANK_DISPATCH(p.nd_it, p.nd_other, DN_LAUNCH((scatter_kernel<int8_t, int32_t, NDI, NDO>), grid, kIdxThreads, 0, p)); break;
            case DN_U8: DN_GS_RANK_DISPATCH(p.nd_it, p.nd_other, DN_LAUNCH((scatter_kernel<uint8_t, int32_t, NDI, NDO>), grid, kIdxThreads, 0, p)); break;
            case DN_I16: DN_GS_RANK_DISPATCH(p.nd_it, p.nd_other, DN_LAUNCH((scatter_kernel<int16_t, int32_t, NDI, NDO>), grid, kIdxThreads, 0, p)); break;
            case DN_U16: DN_GS_RANK_DISPATCH(p.nd_it, p.nd_other, DN_LAUNCH((scatter_kernel<uint16_t, int32_t, NDI, NDO>), grid, kIdxThreads, 0, p)); break;
            case DN_I32: DN_GS_RANK_DISPATCH(p.nd_it, p.nd_other, DN_LAUNCH((scatter_kernel<int32_t, int32_t, NDI, NDO>), grid, kIdxThreads, 0, p)); break;
            case DN_U32: DN_GS_RANK_DISPATCH(p.nd_it, p.nd_other, DN_LAUNCH((scatter_kernel<uint32_t, uint32_t, NDI, NDO>), grid, kIdxThreads, 0, p)); break;
            case DN_I64: DN_GS_RANK_DISPATCH(p.nd_it, p.nd_other, DN_LAUNCH((scatter_kernel<int64_t, int64_t, NDI, NDO>), grid, kIdxThreads, 0, p)); break;
            case DN_U64: DN_GS_RANK_DISPATCH(p.nd_it, p.nd_other, DN_LAUNCH((scatter_kernel<uint64_t, uint64_t, NDI, NDO>), grid, kIdxThreads, 0, p)); break;
            default: st = set_error(DN_ERR_INVALID_ARG, "bad dtype"); break;
            }
            if (st == DN_OK) st = launch_status("scatter kernel");
        }
    }
    if (st == DN_OK && esize < 4) st = dn_convert(t, &acc);  // unchecked narrowing = wrap-around
    scratch_free(scratch);
    if (st != DN_OK) return st;
    return check_index_error("Scatter");
}

dn_status dn_count_true(const dn_tensor *a, int64_t *count) {
    if (!tensor_valid(a) || !count || a->dtype != DN_BOOL) return set_error(DN_ERR_INVALID_ARG, "countTrue: bad argument");
    *count = 0;
    void *scratch = nullptr;
    dn_status st = scratch_alloc(sizeof(unsigned long long), &scratch);
    if (st != DN_OK) return st;
    st = count_true_async(a, static_cast<unsigned long long *>(scratch));
    unsigned long long h = 0;
    cudaError_t e = cudaSuccess;
    if (st == DN_OK) {
        e = cudaMemcpyAsync(&h, scratch, sizeof h, cudaMemcpyDeviceToHost, current_stream());
        if (e == cudaSuccess) e = cudaStreamSynchronize(current_stream());
    }
    scratch_free(scratch);
    if (st != DN_OK) return st;
    if (e != cudaSuccess) return cuda_error(e, "countTrue");
    *count = (int64_t)h;
    return DN_OK;
}

dn_status dn_true_indices(const dn_tensor *t, const dn_tensor *a) {
    if (!tensor_valid(t) || !tensor_valid(a)) return set_error(DN_ERR_INVALID_ARG, "TrueIndices: bad argument");
    if (t->dtype != DN_I64 || a->dtype != DN_BOOL || t->ndims != 2 || t->shape[1] != a->ndims)
        return set_error(DN_ERR_INVALID_ARG, "TrueIndices: target must be int64 of shape [nTrue, %d]", a->ndims);
    if (t->shape[0] == 0 || a->ndims == 0) return DN_OK;
    BoolView m;
    dn_status st = make_bool_view(m, a, "TrueIndices");
    if (st != DN_OK) return st;
    const bool dense = t->stride[1] == 1 && t->stride[0] == t->shape[1] && (reinterpret_cast<uintptr_t>(data_ptr(t)) & 15) == 0;
    if (dense && m.nd == 2) {
        const uint32_t ncols = m.shape[0];
        const FastDiv div = m.div[0];
        merge_bool_view(m);
        CoordSink2D sink2{reinterpret_cast<longlong2 *>(data_ptr(t)), t->shape[0], ncols, div};
        return run_compaction(m, sink2);
    }
    if (dense && m.nd == 1) {
        IndexListSink sink1{reinterpret_cast<int64_t *>(data_ptr(t)), t->shape[0]};
        return run_compaction(m, sink1);
    }
    CoordSink sink;
    sink.t = data_ptr(t);
    sink.ts0 = t->stride[0] * 8;
    sink.ts1 = t->stride[1] * 8;
    sink.cap = t->shape[0];
    sink.nd = m.nd;
    for (int k = 0; k < DN_MAX_DIMS; ++k) {  // coordinates are reported in the ORIGINAL dims
        sink.shape[k] = m.shape[k];
        sink.div[k] = m.div[k];
    }
    merge_bool_view(m);  // reading may use the merged view
    return run_compaction(m, sink);
}

dn_status dn_masked_get(const dn_tensor *t, const dn_tensor *a, const dn_tensor *const *masks, int32_t nmasks) {
    dn_status st = check_masks(a, t, masks, nmasks, "MaskedGet");
    if (st != DN_OK) return st;
    if (num_elements(t) == 0) return DN_OK;
    if (a->ndims == 1 && masks[0]) {  // one mask over a flattened tensor: fused compaction
        BoolView m;
        st = make_bool_view(m, masks[0], "MaskedGet");
        if (st != DN_OK) return st;
        const int sz = dtype_size(a->dtype);
        DN_SWITCH_SIZE(sz, {
            GetSink<B> sink{data_ptr(t), data_ptr(a), t->stride[0] * sz, a->stride[0] * sz, t->shape[0]};
            return run_compaction(m, sink);
        });
    }
    return masked_general<true>(t, a, masks, "MaskedGet");
}

dn_status dn_masked_set(const dn_tensor *t, const dn_tensor *const *masks, int32_t nmasks, const dn_tensor *a) {
    dn_status st = check_masks(t, a, masks, nmasks, "MaskedSet");
    if (st != DN_OK) return st;
    if (num_elements(a) == 0) return DN_OK;
    if (t->ndims == 1 && masks[0]) {
        BoolView m;
        st = make_bool_view(m, masks[0], "MaskedSet");
        if (st != DN_OK) return st;
        const int sz = dtype_size(a->dtype);
        DN_SWITCH_SIZE(sz, {
            SetSink<B> sink{data_ptr(t), data_ptr(a), t->stride[0] * sz, a->stride[0] * sz, a->shape[0]};
            return run_compaction(m, sink);
        });
    }
    return masked_general<false>(a, t, masks, "MaskedSet");
}

}  // extern "C"
