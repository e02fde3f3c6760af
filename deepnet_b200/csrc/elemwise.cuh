// elemwise.cuh — kernel templates and launcher for every element-wise operator of ITensorBackend
// (Tensor/Tensor/TensorBackend.fs:67-118). Replaces Tensor/Tensor/Cuda/Kernels/Elemwise.cuh:13-308 and the launch
// scheme of Work.cuh:25-110 (one thread per N-d position, int64 index math per element, scalar 4-byte accesses).
//
// Design (B200, HBM-bound):
//   * The host canonicalises the operand layouts (ew_plan.cu): drops size-1 dims, flips negative target strides,
//     sorts dims by target stride, merges dims that are contiguous in EVERY operand. What is left is usually 1-D.
//   * ew_kernel<F, VEC>: one template for all ranks. Work items are VEC consecutive elements of the innermost
//     dim; every thread keeps U independent items in flight (U*VEC*size >= 64 bytes of loads per operand).
//     VEC > 1 needs every operand's innermost stride to be 1 (vector access) or 0 (splat) and 16-byte alignment;
//     anything else runs VEC = 1, still coalesced along the target's fastest dim. Index math is 32-bit with
//     multiply-shift division, per work item, not per element.
//   * ew_tiled_kernel<F>: when a source's unit-stride dim differs from the target's (transposed views), a
//     32x32 tile of that source is staged through shared memory so both sides are accessed with full sectors.
//   * Aliasing: target and sources may be the same memory (f3.FillMultiply f3 e, Tensor.Sample/Program.fs:186),
//     so no __restrict__ / ld.global.nc; every thread reads all its inputs before it writes.
#pragma once

#include "common.cuh"

namespace dn {

constexpr int kEwMaxOps = 4;  // target + up to 3 sources
constexpr int kEwThreads = 256;

// Host-side operand description in canonical form (dims innermost-first).
struct EwOperand {
    char *ptr = nullptr;      // base + offset, already in bytes; for index operands: the starting index
    int esize = 0;            // element size in bytes; index operands use 1
    bool is_index = false;    // virtual operand whose "value" is its own linear position (FillIncrementing)
    int64_t stride[DN_MAX_DIMS] = {0};  // in elements
};

struct EwPlan {
    int nops = 0;
    int ndims = 0;
    int64_t shape[DN_MAX_DIMS] = {0};
    EwOperand op[kEwMaxOps];
    int64_t n = 0;  // total elements; 0 = nothing to do
};

// Builds the canonical plan. `srcs[k] == nullptr` with `index_dim >= 0` inserts the virtual index operand.
dn_status ew_make_plan(EwPlan &plan, const dn_tensor *t, const dn_tensor *const *srcs, int nsrc,
                       int index_operand = -1, int index_dim = -1);

template <int NOPS>
struct EwParams {
    char *ptr[NOPS];
    int64_t stride[NOPS][DN_MAX_DIMS];  // BYTES per step of dim d (dim 0: per work item)
    uint32_t shape[DN_MAX_DIMS];
    FastDiv div[DN_MAX_DIMS];
    int32_t ndims;
    uint32_t n;           // work items in this launch
    uint32_t splat_mask;  // bit k: operand k has innermost stride 0 (VEC mode: load one element and splat)
};

struct IndexT {};  // tag type of the virtual index operand; its loaded value is an int64_t position

template <class T> struct LoadedType { using type = T; };
template <> struct LoadedType<IndexT> { using type = int64_t; };
template <class T> constexpr int ew_sizeof() { return std::is_same<T, IndexT>::value ? 1 : (int)sizeof(T); }

// ---- packs -------------------------------------------------------------------------------------------------
template <class T, int N>
struct alignas((N * sizeof(T) >= 16) ? 16 : (N * sizeof(T))) Pack {
    T v[N];
};

template <class P>
__device__ __forceinline__ P load_pack(const char *addr) {
    P out;
    if constexpr (sizeof(P) % 16 == 0) {
        uint4 *dst = reinterpret_cast<uint4 *>(&out);
#pragma unroll
        for (int i = 0; i < (int)(sizeof(P) / 16); ++i) dst[i] = __ldcs(reinterpret_cast<const uint4 *>(addr) + i);
    } else if constexpr (sizeof(P) == 8) {
        *reinterpret_cast<uint2 *>(&out) = __ldcs(reinterpret_cast<const uint2 *>(addr));
    } else if constexpr (sizeof(P) == 4) {
        *reinterpret_cast<unsigned *>(&out) = __ldcs(reinterpret_cast<const unsigned *>(addr));
    } else if constexpr (sizeof(P) == 2) {
        *reinterpret_cast<unsigned short *>(&out) = __ldcs(reinterpret_cast<const unsigned short *>(addr));
    } else {
        static_assert(sizeof(P) == 1, "unsupported pack size");
        *reinterpret_cast<unsigned char *>(&out) = __ldcs(reinterpret_cast<const unsigned char *>(addr));
    }
    return out;
}

template <class P>
__device__ __forceinline__ void store_pack(char *addr, const P &val) {
    if constexpr (sizeof(P) % 16 == 0) {
        const uint4 *src = reinterpret_cast<const uint4 *>(&val);
#pragma unroll
        for (int i = 0; i < (int)(sizeof(P) / 16); ++i) __stcs(reinterpret_cast<uint4 *>(addr) + i, src[i]);
    } else if constexpr (sizeof(P) == 8) {
        __stcs(reinterpret_cast<uint2 *>(addr), *reinterpret_cast<const uint2 *>(&val));
    } else if constexpr (sizeof(P) == 4) {
        __stcs(reinterpret_cast<unsigned *>(addr), *reinterpret_cast<const unsigned *>(&val));
    } else if constexpr (sizeof(P) == 2) {
        __stcs(reinterpret_cast<unsigned short *>(addr), *reinterpret_cast<const unsigned short *>(&val));
    } else {
        __stcs(reinterpret_cast<unsigned char *>(addr), *reinterpret_cast<const unsigned char *>(&val));
    }
}

// Loads the VEC inputs of one work item for operand type T (vector, splat or index).
template <class T, int VEC>
struct InPack {
    using L = typename LoadedType<T>::type;
    Pack<L, VEC> p;
    __device__ __forceinline__ void load(const char *addr, bool splat) {
        if constexpr (std::is_same<T, IndexT>::value) {
            const int64_t base = (int64_t)(intptr_t)addr;
#pragma unroll
            for (int j = 0; j < VEC; ++j) p.v[j] = base + (splat ? 0 : j);
        } else if constexpr (VEC == 1) {
            p = load_pack<Pack<T, 1>>(addr);
        } else {
            if (splat) {
                Pack<T, 1> s = load_pack<Pack<T, 1>>(addr);
#pragma unroll
                for (int j = 0; j < VEC; ++j) p.v[j] = s.v[0];
            } else {
                p = load_pack<Pack<T, VEC>>(addr);
            }
        }
    }
};

// Functor signature helper: every functor derives from EwSig<Out, In0[, In1[, In2]]>.
template <class T> constexpr int ew_vec_size() {
    if constexpr (std::is_void<T>::value || std::is_same<T, IndexT>::value) return 0;
    else return (int)sizeof(T);
}
constexpr int ew_cmax(int a, int b) { return a > b ? a : b; }
constexpr int ew_cmin_nz(int a, int b) { return a == 0 ? b : (b == 0 ? a : (a < b ? a : b)); }

template <class OutT, class A = void, class B = void, class C = void>
struct EwSig {
    using Out = OutT;
    using In0 = A;
    using In1 = B;
    using In2 = C;
    static constexpr int NSRC = (std::is_void<A>::value ? 0 : 1) + (std::is_void<B>::value ? 0 : 1) +
                                (std::is_void<C>::value ? 0 : 1);
    static constexpr int MaxSize =
        ew_cmax(ew_cmax((int)sizeof(OutT), ew_vec_size<A>()), ew_cmax(ew_vec_size<B>(), ew_vec_size<C>()));
    static constexpr int MinSize =
        ew_cmin_nz(ew_cmin_nz((int)sizeof(OutT), ew_vec_size<A>()), ew_cmin_nz(ew_vec_size<B>(), ew_vec_size<C>()));
    // elements per vector work item: 16-byte accesses on the narrowest operand, at most 64 bytes on the widest
    static constexpr int Vec = (16 / MinSize) < (64 / MaxSize) ? (16 / MinSize) : (64 / MaxSize);
    static constexpr bool Tiled = true;  // instantiate the shared-memory transpose kernel for this functor
};

template <int VEC>
struct InPack<void, VEC> {
    __device__ __forceinline__ void load(const char *, bool) {}
};

template <class F, int K> struct InType;
template <class F> struct InType<F, 0> { using type = typename F::In0; };
template <class F> struct InType<F, 1> { using type = typename F::In1; };
template <class F> struct InType<F, 2> { using type = typename F::In2; };
template <class T> struct SmemType { using type = typename LoadedType<T>::type; };
template <> struct SmemType<void> { using type = char; };

template <int NOPS>
__device__ __forceinline__ void ew_offsets(const EwParams<NOPS> &p, uint32_t idx, int64_t (&off)[NOPS]) {
#pragma unroll
    for (int k = 0; k < NOPS; ++k) off[k] = 0;
    uint32_t rem = idx;
#pragma unroll
    for (int d = 0; d < DN_MAX_DIMS; ++d) {
        if (d == p.ndims - 1) {
#pragma unroll
            for (int k = 0; k < NOPS; ++k) off[k] += (int64_t)rem * p.stride[k][d];
            break;
        }
        const uint32_t q = p.div[d].div(rem);
        const uint32_t r = rem - q * p.shape[d];
#pragma unroll
        for (int k = 0; k < NOPS; ++k) off[k] += (int64_t)r * p.stride[k][d];
        rem = q;
    }
}

// F: struct { using Out; using In0[, In1, In2]; static constexpr int NSRC; __device__ Out operator()(...) const; }
template <class F, int VEC, int U>
__global__ void __launch_bounds__(kEwThreads) ew_kernel(const __grid_constant__ EwParams<F::NSRC + 1> p, const F f) {
    constexpr int NOPS = F::NSRC + 1;
    using Out = typename F::Out;
    using P0 = InPack<typename F::In0, VEC>;
    using P1 = InPack<typename F::In1, VEC>;
    using P2 = InPack<typename F::In2, VEC>;

    const uint32_t tile = kEwThreads * U;
    for (uint64_t base = (uint64_t)blockIdx.x * tile; base < p.n; base += (uint64_t)gridDim.x * tile) {
        P0 a[U];
        P1 b[U];
        P2 c[U];
        int64_t toff[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const uint64_t idx = base + (uint32_t)j * kEwThreads + threadIdx.x;
            if (idx < p.n) {
                int64_t off[NOPS];
                ew_offsets<NOPS>(p, (uint32_t)idx, off);
                toff[j] = off[0];
                if constexpr (F::NSRC > 0) a[j].load(p.ptr[1] + off[1], (p.splat_mask >> 1) & 1);
                if constexpr (F::NSRC > 1) b[j].load(p.ptr[2] + off[2], (p.splat_mask >> 2) & 1);
                if constexpr (F::NSRC > 2) c[j].load(p.ptr[3] + off[3], (p.splat_mask >> 3) & 1);
            }
        }
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const uint64_t idx = base + (uint32_t)j * kEwThreads + threadIdx.x;
            if (idx < p.n) {
                Pack<Out, VEC> r;
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    if constexpr (F::NSRC == 0) r.v[e] = f();
                    else if constexpr (F::NSRC == 1) r.v[e] = f(a[j].p.v[e]);
                    else if constexpr (F::NSRC == 2) r.v[e] = f(a[j].p.v[e], b[j].p.v[e]);
                    else r.v[e] = f(a[j].p.v[e], b[j].p.v[e], c[j].p.v[e]);
                }
                store_pack(p.ptr[0] + toff[j], r);
            }
        }
    }
}

// ---- tiled kernel for transposed sources ---------------------------------------------------------------------
// Dims: dT = canonical dim 0 (target stride 1), dS = the dim in which the "S-type" sources have stride 1.
// A CTA (32 x 8 threads) handles one 32(dT) x 32(dS) tile; S-type sources are read with threadIdx.x along dS
// into padded shared memory and consumed with threadIdx.x along dT; T-type sources and the target are accessed
// directly with threadIdx.x along dT. Batch dims (all others) are decomposed once per CTA.
constexpr int kTile = 32;
constexpr int kTileRows = 8;

template <int NOPS>
struct EwTiledParams {
    char *ptr[NOPS];
    int64_t strideT[NOPS];  // bytes per step along dT
    int64_t strideS[NOPS];  // bytes per step along dS
    int64_t strideB[NOPS][DN_MAX_DIMS];  // bytes per step of batch dim b
    uint32_t shapeB[DN_MAX_DIMS];
    FastDiv divB[DN_MAX_DIMS];
    int32_t nbatch_dims;
    uint32_t sizeT, sizeS;
    uint32_t tilesT, tilesS;
    FastDiv divTilesT, divTilesS;
    uint32_t smem_mask;  // bit k: operand k is S-type (staged through shared memory)
};

template <class T>
__device__ __forceinline__ typename LoadedType<T>::type ew_load_elem(const char *addr) {
    if constexpr (std::is_same<T, IndexT>::value) return (int64_t)(intptr_t)addr;
    else return *reinterpret_cast<const T *>(addr);
}

template <class F>
__global__ void __launch_bounds__(kTile *kTileRows) ew_tiled_kernel(const __grid_constant__ EwTiledParams<F::NSRC + 1> p,
                                                                  const F f) {
    constexpr int NOPS = F::NSRC + 1;
    using Out = typename F::Out;
    using L0 = typename SmemType<typename F::In0>::type;
    using L1 = typename SmemType<typename F::In1>::type;
    using L2 = typename SmemType<typename F::In2>::type;
    __shared__ L0 s0[kTile][kTile + 1];
    __shared__ L1 s1[F::NSRC > 1 ? kTile : 1][kTile + 1];
    __shared__ L2 s2[F::NSRC > 2 ? kTile : 1][kTile + 1];

    // blockIdx.x -> (tileT, tileS, batch...)
    uint32_t rem = blockIdx.x;
    uint32_t q = p.divTilesT.div(rem);
    const uint32_t tT = rem - q * p.tilesT;
    rem = q;
    q = p.divTilesS.div(rem);
    const uint32_t tS = rem - q * p.tilesS;
    rem = q;
    int64_t boff[NOPS];
#pragma unroll
    for (int k = 0; k < NOPS; ++k) boff[k] = 0;
#pragma unroll
    for (int d = 0; d < DN_MAX_DIMS; ++d) {
        if (d >= p.nbatch_dims) break;
        const uint32_t qq = p.divB[d].div(rem);
        const uint32_t r = rem - qq * p.shapeB[d];
#pragma unroll
        for (int k = 0; k < NOPS; ++k) boff[k] += (int64_t)r * p.strideB[k][d];
        rem = qq;
    }
    const uint32_t t0 = tT * kTile, sbase = tS * kTile;
    const uint32_t tx = threadIdx.x, ty = threadIdx.y;

    // stage S-type sources: threadIdx.x runs along dS (their unit-stride dim)
#pragma unroll
    for (int i = 0; i < kTile; i += kTileRows) {
        const uint32_t tpos = t0 + ty + i, spos = sbase + tx;
        if (tpos < p.sizeT && spos < p.sizeS) {
            if constexpr (F::NSRC > 0)
                if ((p.smem_mask >> 1) & 1)
                    s0[ty + i][tx] = ew_load_elem<typename F::In0>(
                        p.ptr[1] + boff[1] + (int64_t)tpos * p.strideT[1] + (int64_t)spos * p.strideS[1]);
            if constexpr (F::NSRC > 1)
                if ((p.smem_mask >> 2) & 1)
                    s1[ty + i][tx] = ew_load_elem<typename F::In1>(
                        p.ptr[2] + boff[2] + (int64_t)tpos * p.strideT[2] + (int64_t)spos * p.strideS[2]);
            if constexpr (F::NSRC > 2)
                if ((p.smem_mask >> 3) & 1)
                    s2[ty + i][tx] = ew_load_elem<typename F::In2>(
                        p.ptr[3] + boff[3] + (int64_t)tpos * p.strideT[3] + (int64_t)spos * p.strideS[3]);
        }
    }
    __syncthreads();
    // compute: threadIdx.x runs along dT
#pragma unroll
    for (int i = 0; i < kTile; i += kTileRows) {
        const uint32_t tpos = t0 + tx, spos = sbase + ty + i;
        if (tpos < p.sizeT && spos < p.sizeS) {
            Out r;
            if constexpr (F::NSRC == 1) {
                L0 a = ((p.smem_mask >> 1) & 1) ? s0[tx][ty + i]
                                                : ew_load_elem<typename F::In0>(
                                                      p.ptr[1] + boff[1] + (int64_t)tpos * p.strideT[1] + (int64_t)spos * p.strideS[1]);
                r = f(a);
            } else if constexpr (F::NSRC == 2) {
                L0 a = ((p.smem_mask >> 1) & 1) ? s0[tx][ty + i]
                                                : ew_load_elem<typename F::In0>(
                                                      p.ptr[1] + boff[1] + (int64_t)tpos * p.strideT[1] + (int64_t)spos * p.strideS[1]);
                L1 b = ((p.smem_mask >> 2) & 1) ? s1[tx][ty + i]
                                                : ew_load_elem<typename F::In1>(
                                                      p.ptr[2] + boff[2] + (int64_t)tpos * p.strideT[2] + (int64_t)spos * p.strideS[2]);
                r = f(a, b);
            } else if constexpr (F::NSRC == 3) {
                L0 a = ((p.smem_mask >> 1) & 1) ? s0[tx][ty + i]
                                                : ew_load_elem<typename F::In0>(
                                                      p.ptr[1] + boff[1] + (int64_t)tpos * p.strideT[1] + (int64_t)spos * p.strideS[1]);
                L1 b = ((p.smem_mask >> 2) & 1) ? s1[tx][ty + i]
                                                : ew_load_elem<typename F::In1>(
                                                      p.ptr[2] + boff[2] + (int64_t)tpos * p.strideT[2] + (int64_t)spos * p.strideS[2]);
                L2 c = ((p.smem_mask >> 3) & 1) ? s2[tx][ty + i]
                                                : ew_load_elem<typename F::In2>(
                                                      p.ptr[3] + boff[3] + (int64_t)tpos * p.strideT[3] + (int64_t)spos * p.strideS[3]);
                r = f(a, b, c);
            }
            *reinterpret_cast<Out *>(p.ptr[0] + boff[0] + (int64_t)tpos * p.strideT[0] + (int64_t)spos * p.strideS[0]) = r;
        }
    }
}

// ---- host-side launch logic (non-template parts live in ew_plan.cu) ------------------------------------------
// Decides whether the plan can run with VEC-wide accesses; align[k] = required byte alignment of operand k.
bool ew_can_vectorize(const EwPlan &plan, int vec, const int *esize, int64_t *tail_elems);
// Picks dS for the tiled kernel, or returns -1 when the tiled kernel does not apply.
int ew_pick_tiled_dim(const EwPlan &plan);
// Chunking of the outermost dim so that one launch covers < 2^31 work items.
int64_t ew_chunk_rows(const EwPlan &plan, int64_t inner_items);
int ew_grid_for(int64_t work_items, int items_per_cta);

template <int NOPS>
void ew_fill_params(EwParams<NOPS> &p, const EwPlan &plan, int vec, int64_t outer_begin, int64_t outer_count) {
    const int nd = plan.ndims;
    p.ndims = nd;
    p.splat_mask = 0;
    uint64_t n = 1;
    for (int d = 0; d < DN_MAX_DIMS; ++d) {
        int64_t s = d < nd ? plan.shape[d] : 1;
        if (d == 0) s /= vec;
        if (d == nd - 1 && nd > 1) s = outer_count;
        if (nd == 1 && d == 0) s = outer_count;  // 1-D: the chunk is over work items directly
        p.shape[d] = (uint32_t)s;
        p.div[d].init((uint32_t)(s > 0 ? s : 1));
        if (d < nd) n *= (uint64_t)s;
    }
    p.n = (uint32_t)n;
    for (int k = 0; k < NOPS; ++k) {
        const EwOperand &o = plan.op[k];
        for (int d = 0; d < DN_MAX_DIMS; ++d) {
            int64_t st = d < nd ? o.stride[d] * o.esize : 0;
            if (d == 0) st *= vec;
            p.stride[k][d] = st;
        }
        if (o.stride[0] == 0) p.splat_mask |= 1u << k;
        // advance to the chunk start along the outermost dim (in work items for 1-D)
        const int od = nd - 1;
        int64_t step = o.stride[od] * o.esize * ((nd == 1) ? vec : 1);
        p.ptr[k] = o.ptr + outer_begin * step;
    }
}

template <class F>
dn_status ew_launch_tiled(const EwPlan &plan, const F &f, int dS);

template <class F, int VEC>
dn_status ew_launch_strided(const EwPlan &plan, const F &f, int64_t n_inner0_elems) {
    constexpr int NOPS = F::NSRC + 1;
    // U: keep >= 64 bytes of loads per source in flight per thread, within a sane register budget
    constexpr int maxsz = F::MaxSize;
    constexpr int U = (VEC * maxsz >= 64) ? 1 : ((VEC * maxsz >= 32) ? 2 : 4);
    const int nd = plan.ndims;
    // inner items = product of all dims but the outermost, in work items
    int64_t inner = 1;
    for (int d = 0; d < nd - 1; ++d) inner *= (d == 0 ? n_inner0_elems / VEC : plan.shape[d]);
    const int64_t outer_total = (nd == 1) ? n_inner0_elems / VEC : plan.shape[nd - 1];
    const int64_t max_items = (int64_t)1 << 30;
    int64_t chunk = inner > 0 ? max_items / inner : max_items;
    if (chunk < 1) return set_error(DN_ERR_UNSUPPORTED, "element-wise: inner extent exceeds 2^30 work items");
    EwPlan local = plan;
    local.shape[0] = n_inner0_elems;
    for (int64_t begin = 0; begin < outer_total; begin += chunk) {
        const int64_t count = (outer_total - begin < chunk) ? outer_total - begin : chunk;
        EwParams<NOPS> p;
        ew_fill_params<NOPS>(p, local, VEC, begin, count);
        if (p.n == 0) continue;
        const int grid = ew_grid_for(p.n, kEwThreads * U);
        DN_LAUNCH((ew_kernel<F, VEC, U>), grid, kEwThreads, 0, p, f);
    }
    return launch_status("element-wise kernel");
}

// Entry: run functor F over the plan. VECW = elements per vector work item (16 / smallest element size, capped).
template <class F>
dn_status ew_run(EwPlan &plan, const F &f) {
    if (plan.n == 0) return DN_OK;
    constexpr int NOPS = F::NSRC + 1;
    constexpr int VEC = F::Vec;
    int esize[kEwMaxOps];
    for (int k = 0; k < NOPS; ++k) esize[k] = plan.op[k].esize;
    if constexpr (VEC > 1) {
        int64_t tail = 0;
        if (ew_can_vectorize(plan, VEC, esize, &tail)) {
            const int64_t body = plan.shape[0] - tail;
            if (body > 0) {
                dn_status st = ew_launch_strided<F, VEC>(plan, f, body);
                if (st != DN_OK) return st;
            }
            if (tail > 0) {  // only possible for 1-D plans: finish the last < VEC elements with the scalar kernel
                EwPlan tp = plan;
                for (int k = 0; k < NOPS; ++k) tp.op[k].ptr += body * tp.op[k].stride[0] * tp.op[k].esize;
                tp.shape[0] = tail;
                return ew_launch_strided<F, 1>(tp, f, tail);
            }
            return DN_OK;
        }
    }
    if constexpr (F::NSRC > 0 && F::Tiled) {
        const int dS = ew_pick_tiled_dim(plan);
        if (dS > 0) return ew_launch_tiled<F>(plan, f, dS);
    }
    return ew_launch_strided<F, 1>(plan, f, plan.shape[0]);
}

template <class F>
dn_status ew_launch_tiled(const EwPlan &plan, const F &f, int dS) {
    constexpr int NOPS = F::NSRC + 1;
    EwTiledParams<NOPS> p;
    const int nd = plan.ndims;
    p.sizeT = (uint32_t)plan.shape[0];
    p.sizeS = (uint32_t)plan.shape[dS];
    p.tilesT = (p.sizeT + kTile - 1) / kTile;
    p.tilesS = (p.sizeS + kTile - 1) / kTile;
    p.divTilesT.init(p.tilesT);
    p.divTilesS.init(p.tilesS);
    p.smem_mask = 0;
    int nb = 0;
    int64_t nbatch = 1;
    for (int d = 0; d < DN_MAX_DIMS; ++d) {
        p.shapeB[d] = 1;
        p.divB[d].init(1);
    }
    for (int d = 1; d < nd; ++d) {
        if (d == dS) continue;
        p.shapeB[nb] = (uint32_t)plan.shape[d];
        p.divB[nb].init((uint32_t)plan.shape[d]);
        for (int k = 0; k < NOPS; ++k) p.strideB[k][nb] = plan.op[k].stride[d] * plan.op[k].esize;
        nbatch *= plan.shape[d];
        ++nb;
    }
    for (int k = 0; k < NOPS; ++k)
        for (int d = nb; d < DN_MAX_DIMS; ++d) p.strideB[k][d] = 0;
    p.nbatch_dims = nb;
    for (int k = 0; k < NOPS; ++k) {
        p.ptr[k] = plan.op[k].ptr;
        p.strideT[k] = plan.op[k].stride[0] * plan.op[k].esize;
        p.strideS[k] = plan.op[k].stride[dS] * plan.op[k].esize;
        if (k > 0 && !plan.op[k].is_index && plan.op[k].stride[dS] == 1 && plan.op[k].stride[0] != 1 &&
            plan.op[k].stride[0] != 0)
            p.smem_mask |= 1u << k;
    }
    const int64_t tiles = (int64_t)p.tilesT * p.tilesS * nbatch;
    if (tiles >= ((int64_t)1 << 31)) return ew_launch_strided<F, 1>(plan, f, plan.shape[0]);
    DN_LAUNCH((ew_tiled_kernel<F>), (unsigned)tiles, dim3(kTile, kTileRows), 0, p, f);
    return launch_status("element-wise tiled kernel");
}

}  // namespace dn
