// elemwise.cuh — kernel templates and launcher for every element-wise operator of ITensorBackend
// (Tensor/Tensor/TensorBackend.fs:67-118). Replaces Tensor/Tensor/Cuda/Kernels/Elemwise.cuh:13-308 and the launch
// scheme of Work.cuh:25-110 (one thread per N-d position, int64 index math per element, scalar 4-byte accesses).
//
// Design (B200, HBM-bound):
//   * The host canonicalises the operand layouts (ew_plan.cu): drops size-1 dims, flips negative target strides,
//     sorts dims by target stride, merges dims that are contiguous in EVERY operand. What is left is usually 1-D.
//   * ew_kernel<F, VEC, U, ND>: work items are VEC consecutive elements of the innermost dim (32 bytes on the
//     widest operand); every thread keeps U independent items in flight. VEC > 1 needs every operand's innermost
//     stride to be 1 (vector access) or 0 (splat) and 16-byte alignment; anything else runs VEC = 1, still
//     coalesced along the target's fastest dim, with the rank fixed at compile time (ND = 1, 2, 3) so that the
//     per-element index math is a couple of multiply-shift divisions. Index math is 32-bit.
//   * ew_xpose_kernel<F, M>: when a source's unit-stride dim (dS) differs from the target's (dT) — transposed
//     views — every thread owns an M x M micro-tile: it reads the "S-type" sources as M vectors along dS, the
//     "T-type" sources as M vectors along dT, transposes in REGISTERS and writes M vectors along dT. Lanes are
//     laid out 8 (dS) x 4 (dT), so every warp-level access covers whole 32-byte sectors on both sides; no shared
//     memory, no barriers.
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

// Builds the canonical plan. `index_operand >= 0` makes that source the virtual index operand along `index_dim`.
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
    uint32_t wide;        // every access of 32 bytes or more is 32-byte aligned: use 256-bit LDG/STG (sm_100)
    uint32_t rev_mask;    // bit k: source k has innermost stride -1 (VEC mode: load the pack and reverse it)
};

struct IndexT {};  // tag type of the virtual index operand; its loaded value is an int64_t position

template <class T> struct LoadedType { using type = T; };
template <> struct LoadedType<IndexT> { using type = int64_t; };
template <> struct LoadedType<void> { using type = char; };

// ---- packs -------------------------------------------------------------------------------------------------
template <class T, int N>
struct alignas((N * sizeof(T) >= 16) ? 16 : (N * sizeof(T))) Pack {
    T v[N];
};

// 256-bit global accesses (sm_100: LDG/STG.E.256). Measured on B200 (tools/fill_probe.cu): a thread that writes
// its 32 bytes as two adjacent 128-bit stores makes every warp-level store touch half of each 32-byte sector
// (3.7 TB/s write-only); one 256-bit store per thread writes whole sectors (7.0 TB/s).
__device__ __forceinline__ void ldg256(const char *addr, uint4 &lo, uint4 &hi) {
    asm volatile("ld.global.cs.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
                 : "l"(addr)
                 : "memory");
}
__device__ __forceinline__ void stg256(char *addr, const uint4 &lo, const uint4 &hi) {
    asm volatile("st.global.cs.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(addr), "r"(lo.x), "r"(lo.y), "r"(lo.z),
                 "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w)
                 : "memory");
}

template <class P>
__device__ __forceinline__ P load_pack(const char *addr, bool wide = false) {
    P out;
    if constexpr (sizeof(P) % 32 == 0) {
        uint4 *dst = reinterpret_cast<uint4 *>(&out);
        if (wide) {
#pragma unroll
            for (int i = 0; i < (int)(sizeof(P) / 32); ++i) ldg256(addr + 32 * i, dst[2 * i], dst[2 * i + 1]);
        } else {
#pragma unroll
            for (int i = 0; i < (int)(sizeof(P) / 16); ++i) dst[i] = __ldcs(reinterpret_cast<const uint4 *>(addr) + i);
        }
    } else if constexpr (sizeof(P) % 16 == 0) {
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
__device__ __forceinline__ void store_pack(char *addr, const P &val, bool wide = false) {
    if constexpr (sizeof(P) % 32 == 0) {
        const uint4 *src = reinterpret_cast<const uint4 *>(&val);
        if (wide) {
#pragma unroll
            for (int i = 0; i < (int)(sizeof(P) / 32); ++i) stg256(addr + 32 * i, src[2 * i], src[2 * i + 1]);
        } else {
#pragma unroll
            for (int i = 0; i < (int)(sizeof(P) / 16); ++i) __stcs(reinterpret_cast<uint4 *>(addr) + i, src[i]);
        }
    } else if constexpr (sizeof(P) % 16 == 0) {
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
    __device__ __forceinline__ void load(const char *addr, bool splat, bool wide = false, bool rev = false) {
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
                p = load_pack<Pack<T, VEC>>(addr, wide);
                if (rev) {  // reversed view: addr is the pack's lowest address, element j lives at VEC-1-j
#pragma unroll
                    for (int j = 0; j < VEC / 2; ++j) {
                        const L tmp = p.v[j];
                        p.v[j] = p.v[VEC - 1 - j];
                        p.v[VEC - 1 - j] = tmp;
                    }
                }
            }
        }
    }
};
template <int VEC>
struct InPack<void, VEC> {
    __device__ __forceinline__ void load(const char *, bool, bool = false, bool = false) {}
};

// Functor signature helper: every functor derives from EwSig<Out, In0[, In1[, In2]]>.
template <class T> constexpr int ew_vec_size() {
    if constexpr (std::is_void<T>::value || std::is_same<T, IndexT>::value) return 0;
    else return (int)sizeof(T);
}
constexpr int ew_cmax(int a, int b) { return a > b ? a : b; }
constexpr int ew_cmin(int a, int b) { return a < b ? a : b; }
constexpr int ew_cmin_nz(int a, int b) { return a == 0 ? b : (b == 0 ? a : (a < b ? a : b)); }
constexpr int ew_pick_vec(int min_size, int max_size) {
    int v = ew_cmax(16 / min_size, 32 / max_size);
    while (v > 1 && v * max_size > 64) v /= 2;
    return v;
}

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
    // elements per vector work item: at least 16 bytes on the narrowest operand, 32 bytes on the widest if that
    // does not push the widest past 64 bytes
    static constexpr int Vec = ew_pick_vec(MinSize, MaxSize);
    static constexpr bool Tiled = true;  // instantiate the register-transpose kernel for this functor
    // Packed functors (1-byte bool operators) also provide `uint32_t packed(uint32_t...)` over four elements per
    // 32-bit word; the vector kernel uses it instead of sixteen byte-wide evaluations per 128-bit access.
    static constexpr bool Packed = false;
    // VectorEval functors evaluate a whole work item at once: `eval(Out (&r)[VEC], const In0 (&a)[VEC], ...)`
    // (the fused-expression interpreter amortises its instruction decode over the VEC elements).
    static constexpr bool VectorEval = false;
    // MultiEval functors evaluate ALL U work items of a thread in one call:
    // `eval_multi<VEC, U>(Pack<Out,VEC> (&r)[U], const P0 (&a)[U], const P1 (&b)[U], const P2 (&c)[U])`
    // (the interpreter then decodes every program instruction once per thread iteration, not once per item).
    static constexpr bool MultiEval = false;
    static constexpr int MaxInFlight = 8;  // cap on U, the work items a thread keeps in flight
    static constexpr int MinBlocks = 1;    // __launch_bounds__ minimum resident CTAs per SM
};

template <int NOPS, int ND>
__device__ __forceinline__ void ew_offsets(const EwParams<NOPS> &p, uint32_t idx, int64_t (&off)[NOPS]) {
#pragma unroll
    for (int k = 0; k < NOPS; ++k) off[k] = 0;
    uint32_t rem = idx;
    if constexpr (ND > 0) {
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            uint32_t x = rem;
            if (d < ND - 1) {
                const uint32_t q = p.div[d].div(rem);
                x = rem - q * p.shape[d];
                rem = q;
            }
#pragma unroll
            for (int k = 0; k < NOPS; ++k) off[k] += (int64_t)x * p.stride[k][d];
        }
    } else {
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
}

// F: struct : EwSig<...> { __device__ Out operator()(In...) const; }
template <class F, int VEC, int U, int ND>
__global__ void __launch_bounds__(kEwThreads, F::MinBlocks) ew_kernel(const __grid_constant__ EwParams<F::NSRC + 1> p, const __grid_constant__ F f) {
    constexpr int NOPS = F::NSRC + 1;
    using Out = typename F::Out;
    using P0 = InPack<typename F::In0, VEC>;
    using P1 = InPack<typename F::In1, VEC>;
    using P2 = InPack<typename F::In2, VEC>;

    const uint32_t tile = kEwThreads * U;
    const bool wide = VEC > 1 && p.wide != 0;
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
                ew_offsets<NOPS, ND>(p, (uint32_t)idx, off);
                toff[j] = off[0];
                if constexpr (F::NSRC > 0) a[j].load(p.ptr[1] + off[1], (p.splat_mask >> 1) & 1, wide, VEC > 1 && ((p.rev_mask >> 1) & 1));
                if constexpr (F::NSRC > 1) b[j].load(p.ptr[2] + off[2], (p.splat_mask >> 2) & 1, wide, VEC > 1 && ((p.rev_mask >> 2) & 1));
                if constexpr (F::NSRC > 2) c[j].load(p.ptr[3] + off[3], (p.splat_mask >> 3) & 1, wide, VEC > 1 && ((p.rev_mask >> 3) & 1));
            }
        }
        if constexpr (F::MultiEval) {
            // inactive items compute on whatever their registers hold; only active results are stored
            Pack<Out, VEC> r[U];
            f.template eval_multi<VEC, U>(r, a, b, c);
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const uint64_t idx = base + (uint32_t)j * kEwThreads + threadIdx.x;
                if (idx < p.n) store_pack(p.ptr[0] + toff[j], r[j], wide);
            }
        } else {
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const uint64_t idx = base + (uint32_t)j * kEwThreads + threadIdx.x;
            if (idx < p.n) {
                Pack<Out, VEC> r;
                if constexpr (F::VectorEval) {
                    if constexpr (F::NSRC == 1) f.template eval<VEC>(r.v, a[j].p.v, a[j].p.v, a[j].p.v);
                    else if constexpr (F::NSRC == 2) f.template eval<VEC>(r.v, a[j].p.v, b[j].p.v, b[j].p.v);
                    else f.template eval<VEC>(r.v, a[j].p.v, b[j].p.v, c[j].p.v);
                } else if constexpr (F::Packed && VEC % 4 == 0) {
                    uint32_t *wr = reinterpret_cast<uint32_t *>(&r);
#pragma unroll
                    for (int e = 0; e < VEC / 4; ++e) {
                        if constexpr (F::NSRC == 1) wr[e] = f.packed(reinterpret_cast<const uint32_t *>(&a[j].p)[e]);
                        else wr[e] = f.packed(reinterpret_cast<const uint32_t *>(&a[j].p)[e],
                                              reinterpret_cast<const uint32_t *>(&b[j].p)[e]);
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) {
                        if constexpr (F::NSRC == 0) r.v[e] = f();
                        else if constexpr (F::NSRC == 1) r.v[e] = f(a[j].p.v[e]);
                        else if constexpr (F::NSRC == 2) r.v[e] = f(a[j].p.v[e], b[j].p.v[e]);
                        else r.v[e] = f(a[j].p.v[e], b[j].p.v[e], c[j].p.v[e]);
                    }
                }
                store_pack(p.ptr[0] + toff[j], r, wide);
            }
        }
        }  // !MultiEval
    }
}

// ---- register-transpose kernel for transposed sources ----------------------------------------------------------
// Dims: dT = canonical dim 0 (target stride 1), dS = the dim in which the "S-type" sources have stride 1; every
// other dim is a batch dim decomposed once per CTA. Operand access modes (2 bits each in `modes`):
//   0 = T-vector: unit stride along dT, aligned   -> M vector accesses along dT (one per dS position)
//   1 = S-vector: unit stride along dS, aligned   -> M vector accesses along dS (one per dT position)
//   2 = scalar:   anything else (broadcast, misaligned, arbitrary strides) -> M*M element accesses
// CTA = 8 warps as 2 (dS) x 4 (dT); warp = 8 lanes (dS) x 4 lanes (dT); thread = K micro-tiles of M x M elements
// stepped along dS. CTA tile: 16*M*K (dS) x 16*M (dT).
template <int NOPS>
struct EwXposeParams {
    char *ptr[NOPS];
    int64_t strideT[NOPS];  // bytes per element step along dT
    int64_t strideS[NOPS];  // bytes per element step along dS
    int64_t strideB[NOPS][DN_MAX_DIMS];  // bytes per step of batch dim b
    uint32_t shapeB[DN_MAX_DIMS];
    FastDiv divB[DN_MAX_DIMS];
    int32_t nbatch_dims;
    uint32_t sizeT, sizeS;
    uint32_t tilesT, tilesS;
    FastDiv divTilesT, divTilesS;
    uint32_t modes;  // 2 bits per operand
};

template <class T>
__device__ __forceinline__ typename LoadedType<T>::type ew_load_elem(const char *addr) {
    if constexpr (std::is_same<T, IndexT>::value) return (int64_t)(intptr_t)addr;
    else if constexpr (std::is_void<T>::value) return 0;
    else return *reinterpret_cast<const T *>(addr);
}

// val[j][i] = element at (dS position s0+j, dT position t0+i)
template <class T, int M>
__device__ __forceinline__ void xpose_load(typename LoadedType<T>::type (&val)[M][M], const char *base, int64_t strideT,
                                           int64_t strideS, int mode) {
    if constexpr (std::is_void<T>::value) {
        return;
    } else if constexpr (std::is_same<T, IndexT>::value) {
#pragma unroll
        for (int j = 0; j < M; ++j)
#pragma unroll
            for (int i = 0; i < M; ++i) val[j][i] = (int64_t)(intptr_t)(base + i * strideT + j * strideS);
    } else {
        if (mode == 0) {
#pragma unroll
            for (int j = 0; j < M; ++j) {
                const Pack<T, M> v = load_pack<Pack<T, M>>(base + j * strideS);
#pragma unroll
                for (int i = 0; i < M; ++i) val[j][i] = v.v[i];
            }
        } else if (mode == 1) {
#pragma unroll
            for (int i = 0; i < M; ++i) {
                const Pack<T, M> v = load_pack<Pack<T, M>>(base + i * strideT);
#pragma unroll
                for (int j = 0; j < M; ++j) val[j][i] = v.v[j];
            }
        } else {
#pragma unroll
            for (int j = 0; j < M; ++j)
#pragma unroll
                for (int i = 0; i < M; ++i) val[j][i] = *reinterpret_cast<const T *>(base + i * strideT + j * strideS);
        }
    }
}

template <class F, int M, int K>
__global__ void __launch_bounds__(kEwThreads) ew_xpose_kernel(const __grid_constant__ EwXposeParams<F::NSRC + 1> p,
                                                            const F f) {
    constexpr int NOPS = F::NSRC + 1;
    using Out = typename F::Out;
    using L0 = typename LoadedType<typename F::In0>::type;
    using L1 = typename LoadedType<typename F::In1>::type;
    using L2 = typename LoadedType<typename F::In2>::type;

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
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t t0 = (tT * 16 + (warp >> 1) * 4 + (lane >> 3)) * M;
    const uint32_t sbase = (tS * 16 * K + (warp & 1) * 8 * K + (lane & 7)) * M;  // + kk*8*M per micro-tile
    if (t0 >= p.sizeT) return;

    const bool t_full = t0 + M <= p.sizeT;
    // Fast path: every operand is accessed with vectors (modes 0/1) and all K micro-tiles are interior. All loads
    // of all operands are issued back to back (no per-operand branches), then transposed with selects.
    if ((p.modes & 0xAAu) == 0 && (p.modes & 3u) == 0 && t_full && sbase + (K - 1) * 8 * M + M <= p.sizeS) {
        using V0 = Pack<std::conditional_t<std::is_void<typename F::In0>::value, char, typename F::In0>, M>;
        using V1 = Pack<std::conditional_t<std::is_void<typename F::In1>::value, char, typename F::In1>, M>;
        using V2 = Pack<std::conditional_t<std::is_void<typename F::In2>::value, char, typename F::In2>, M>;
        V0 va[K][M];
        V1 vb[K][M];
        V2 vc[K][M];
        const bool m1 = (p.modes >> 2) & 1, m2 = (p.modes >> 4) & 1, m3 = (p.modes >> 6) & 1;
        const int64_t st1 = m1 ? p.strideT[1] : p.strideS[1];  // step between the M vectors of operand 1
        const int64_t st2 = m2 ? p.strideT[2] : p.strideS[2];
        const int64_t st3 = m3 ? p.strideT[3] : p.strideS[3];
#pragma unroll
        for (int kk = 0; kk < K; ++kk) {
            const uint32_t s0 = sbase + kk * 8 * M;
#pragma unroll
            for (int m = 0; m < M; ++m) {
                if constexpr (F::NSRC > 0 && !std::is_same<typename F::In0, IndexT>::value)
                    va[kk][m] = load_pack<V0>(p.ptr[1] + boff[1] + (int64_t)t0 * p.strideT[1] + (int64_t)s0 * p.strideS[1] + m * st1);
                if constexpr (F::NSRC > 1 && !std::is_same<typename F::In1, IndexT>::value)
                    vb[kk][m] = load_pack<V1>(p.ptr[2] + boff[2] + (int64_t)t0 * p.strideT[2] + (int64_t)s0 * p.strideS[2] + m * st2);
                if constexpr (F::NSRC > 2 && !std::is_same<typename F::In2, IndexT>::value)
                    vc[kk][m] = load_pack<V2>(p.ptr[3] + boff[3] + (int64_t)t0 * p.strideT[3] + (int64_t)s0 * p.strideS[3] + m * st3);
            }
        }
#pragma unroll
        for (int kk = 0; kk < K; ++kk) {
            const uint32_t s0 = sbase + kk * 8 * M;
            char *tbase = p.ptr[0] + boff[0] + (int64_t)t0 * p.strideT[0] + (int64_t)s0 * p.strideS[0];
#pragma unroll
            for (int j = 0; j < M; ++j) {
                Pack<Out, M> r;
#pragma unroll
                for (int i = 0; i < M; ++i) {
                    // element (dS = s0+j, dT = t0+i): mode 0 -> vector j, lane i; mode 1 -> vector i, lane j
                    if constexpr (F::NSRC == 1) r.v[i] = f(m1 ? va[kk][i].v[j] : va[kk][j].v[i]);
                    else if constexpr (F::NSRC == 2)
                        r.v[i] = f(m1 ? va[kk][i].v[j] : va[kk][j].v[i], m2 ? vb[kk][i].v[j] : vb[kk][j].v[i]);
                    else
                        r.v[i] = f(m1 ? va[kk][i].v[j] : va[kk][j].v[i], m2 ? vb[kk][i].v[j] : vb[kk][j].v[i],
                                   m3 ? vc[kk][i].v[j] : vc[kk][j].v[i]);
                }
                store_pack(tbase + j * p.strideS[0], r);
            }
        }
        return;
    }

    L0 a[K][M][M];
    L1 b[K][M][M];
    L2 c[K][M][M];
#pragma unroll
    for (int kk = 0; kk < K; ++kk) {
        const uint32_t s0 = sbase + kk * 8 * M;
        if (t_full && s0 + M <= p.sizeS) {
            if constexpr (F::NSRC > 0)
                xpose_load<typename F::In0, M>(a[kk], p.ptr[1] + boff[1] + (int64_t)t0 * p.strideT[1] + (int64_t)s0 * p.strideS[1],
                                               p.strideT[1], p.strideS[1], (p.modes >> 2) & 3);
            if constexpr (F::NSRC > 1)
                xpose_load<typename F::In1, M>(b[kk], p.ptr[2] + boff[2] + (int64_t)t0 * p.strideT[2] + (int64_t)s0 * p.strideS[2],
                                               p.strideT[2], p.strideS[2], (p.modes >> 4) & 3);
            if constexpr (F::NSRC > 2)
                xpose_load<typename F::In2, M>(c[kk], p.ptr[3] + boff[3] + (int64_t)t0 * p.strideT[3] + (int64_t)s0 * p.strideS[3],
                                               p.strideT[3], p.strideS[3], (p.modes >> 6) & 3);
        }
    }
#pragma unroll
    for (int kk = 0; kk < K; ++kk) {
        const uint32_t s0 = sbase + kk * 8 * M;
        if (s0 >= p.sizeS) continue;
        char *tbase = p.ptr[0] + boff[0] + (int64_t)t0 * p.strideT[0] + (int64_t)s0 * p.strideS[0];
        if (t_full && s0 + M <= p.sizeS) {
            const int tmode = p.modes & 3;
#pragma unroll
            for (int j = 0; j < M; ++j) {
                Pack<Out, M> r;
#pragma unroll
                for (int i = 0; i < M; ++i) {
                    if constexpr (F::NSRC == 1) r.v[i] = f(a[kk][j][i]);
                    else if constexpr (F::NSRC == 2) r.v[i] = f(a[kk][j][i], b[kk][j][i]);
                    else r.v[i] = f(a[kk][j][i], b[kk][j][i], c[kk][j][i]);
                }
                if (tmode == 0) {
                    store_pack(tbase + j * p.strideS[0], r);
                } else {
#pragma unroll
                    for (int i = 0; i < M; ++i)
                        *reinterpret_cast<Out *>(tbase + i * p.strideT[0] + j * p.strideS[0]) = r.v[i];
                }
            }
        } else {
            // edge micro-tile: element-wise, predicated
            for (int j = 0; j < M; ++j)
                for (int i = 0; i < M; ++i) {
                    if (t0 + i >= p.sizeT || s0 + j >= p.sizeS) continue;
                    Out r;
                    if constexpr (F::NSRC == 1) {
                        r = f(ew_load_elem<typename F::In0>(p.ptr[1] + boff[1] + (int64_t)(t0 + i) * p.strideT[1] + (int64_t)(s0 + j) * p.strideS[1]));
                    } else if constexpr (F::NSRC == 2) {
                        r = f(ew_load_elem<typename F::In0>(p.ptr[1] + boff[1] + (int64_t)(t0 + i) * p.strideT[1] + (int64_t)(s0 + j) * p.strideS[1]),
                              ew_load_elem<typename F::In1>(p.ptr[2] + boff[2] + (int64_t)(t0 + i) * p.strideT[2] + (int64_t)(s0 + j) * p.strideS[2]));
                    } else {
                        r = f(ew_load_elem<typename F::In0>(p.ptr[1] + boff[1] + (int64_t)(t0 + i) * p.strideT[1] + (int64_t)(s0 + j) * p.strideS[1]),
                              ew_load_elem<typename F::In1>(p.ptr[2] + boff[2] + (int64_t)(t0 + i) * p.strideT[2] + (int64_t)(s0 + j) * p.strideS[2]),
                              ew_load_elem<typename F::In2>(p.ptr[3] + boff[3] + (int64_t)(t0 + i) * p.strideT[3] + (int64_t)(s0 + j) * p.strideS[3]));
                    }
                    *reinterpret_cast<Out *>(tbase + i * p.strideT[0] + j * p.strideS[0]) = r;
                }
        }
    }
}

// ---- host-side launch logic (non-template parts live in ew_plan.cu) ------------------------------------------
bool ew_can_vectorize(const EwPlan &plan, int vec, const int *esize, int64_t *tail_elems);
bool ew_find_peel(const EwPlan &plan, int vec, const int *esize, int64_t *head);
// Picks dS for the transpose kernel, or returns -1 when it does not apply.
int ew_pick_tiled_dim(const EwPlan &plan);
int ew_grid_for(int64_t work_items, int items_per_cta);
// Access mode (0 T-vector, 1 S-vector, 2 scalar) of operand k for micro-tile size m.
int ew_xpose_mode(const EwPlan &plan, int k, int dS, int m);

template <int NOPS>
void ew_fill_params(EwParams<NOPS> &p, const EwPlan &plan, int vec) {
    const int nd = plan.ndims;
    p.ndims = nd;
    p.splat_mask = 0;
    uint64_t n = 1;
    for (int d = 0; d < DN_MAX_DIMS; ++d) {
        int64_t s = d < nd ? plan.shape[d] : 1;
        if (d == 0) s /= vec;
        p.shape[d] = (uint32_t)s;
        p.div[d].init((uint32_t)(s > 0 ? s : 1));
        if (d < nd) n *= (uint64_t)s;
    }
    p.n = (uint32_t)n;
    p.wide = vec > 1 ? 1u : 0u;
    p.rev_mask = 0;
    for (int k = 0; k < NOPS; ++k) {
        const EwOperand &o = plan.op[k];
        for (int d = 0; d < DN_MAX_DIMS; ++d) {
            int64_t st = d < nd ? o.stride[d] * o.esize : 0;
            if (d == 0) st *= vec;
            p.stride[k][d] = st;
        }
        if (o.stride[0] == 0) p.splat_mask |= 1u << k;
        p.ptr[k] = o.ptr;
        if (vec > 1 && !o.is_index && o.stride[0] == -1) {  // reversed source: packs are addressed by their lowest byte
            p.rev_mask |= 1u << k;
            p.ptr[k] = o.ptr - (int64_t)(vec - 1) * o.esize;
        }
        // 256-bit accesses need 32-byte alignment of every access this operand makes with packs of >= 32 bytes
        if (!o.is_index && o.stride[0] != 0 && (int64_t)vec * o.esize >= 32) {
            if (((uintptr_t)p.ptr[k]) % 32 != 0) p.wide = 0;
            for (int d = 1; d < nd; ++d)
                if ((o.stride[d] * o.esize) % 32 != 0) p.wide = 0;
        }
    }
}

template <class F, int VEC, int U, int ND>
void ew_launch_one(const EwParams<F::NSRC + 1> &p, const F &f) {
    const int grid = ew_grid_for(p.n, kEwThreads * U);
    DN_LAUNCH((ew_kernel<F, VEC, U, ND>), grid, kEwThreads, 0, p, f);
}

// Launches the strided kernel over `plan` (dim 0 holds a multiple of VEC elements). One launch covers at most
// 2^30 work items (32-bit index math); larger plans are halved along their outermost non-trivial dim, recursively.
template <class F, int VEC>
dn_status ew_launch_strided(const EwPlan &plan, const F &f, int64_t n_inner0_elems) {
    constexpr int NOPS = F::NSRC + 1;
    // U: independent work items per thread; 64-128 bytes of loads per source in flight per thread
    constexpr int U0 = VEC == 1 ? (F::MaxSize >= 8 ? 4 : 8) : ((VEC * F::MaxSize >= 64) ? 1 : ((VEC * F::MaxSize >= 32) ? 2 : 4));
    constexpr int U = U0 < F::MaxInFlight ? U0 : F::MaxInFlight;
    EwPlan local = plan;
    local.shape[0] = n_inner0_elems;
    const int nd = local.ndims;
    int64_t items = local.shape[0] / VEC;
    for (int d = 1; d < nd; ++d) items *= local.shape[d];
    if (items == 0) return DN_OK;
    if (items > ((int64_t)1 << 30)) {
        int d = nd - 1;
        while (d > 0 && local.shape[d] == 1) --d;
        const int64_t unit = d == 0 ? VEC : 1;
        const int64_t half = (local.shape[d] / unit / 2) * unit;
        if (half <= 0) return set_error(DN_ERR_UNSUPPORTED, "element-wise: cannot split a plan of %lld work items", (long long)items);
        EwPlan lo = local, hi = local;
        lo.shape[d] = half;
        hi.shape[d] = local.shape[d] - half;
        for (int k = 0; k < NOPS; ++k) hi.op[k].ptr += half * hi.op[k].stride[d] * hi.op[k].esize;
        dn_status st = ew_launch_strided<F, VEC>(lo, f, lo.shape[0]);
        if (st != DN_OK) return st;
        return ew_launch_strided<F, VEC>(hi, f, hi.shape[0]);
    }
    EwParams<NOPS> p;
    ew_fill_params<NOPS>(p, local, VEC);
    if constexpr (VEC == 1) {
        // scalar path: the rank is a compile-time constant for the common cases
        if (nd == 1) ew_launch_one<F, 1, U, 1>(p, f);
        else if (nd == 2) ew_launch_one<F, 1, U, 2>(p, f);
        else if (nd == 3) ew_launch_one<F, 1, U, 3>(p, f);
        else ew_launch_one<F, 1, U, 0>(p, f);
    } else {
        ew_launch_one<F, VEC, U, 0>(p, f);
    }
    return launch_status("element-wise kernel");
}

template <class F>
dn_status ew_launch_xpose(const EwPlan &plan, const F &f, int dS);

// Entry: run functor F over the plan.
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
        int64_t head = 0;
        if (ew_find_peel(plan, VEC, esize, &head)) {
            // misaligned rows: vector body + thin scalar head / tail columns
            const int64_t tail = (plan.shape[0] - head) % VEC;
            const int64_t body = plan.shape[0] - head - tail;
            auto columns = [&](int64_t first, int64_t count) {
                EwPlan sub = plan;
                for (int k = 0; k < NOPS; ++k) sub.op[k].ptr += first * sub.op[k].stride[0] * sub.op[k].esize;
                sub.shape[0] = count;
                return sub;
            };
            dn_status st = ew_launch_strided<F, VEC>(columns(head, body), f, body);
            if (st == DN_OK && head > 0) st = ew_launch_strided<F, 1>(columns(0, head), f, head);
            if (st == DN_OK && tail > 0) st = ew_launch_strided<F, 1>(columns(head + body, tail), f, tail);
            return st;
        }
    }
    if constexpr (F::NSRC > 0 && F::Tiled) {
        const int dS = ew_pick_tiled_dim(plan);
        if (dS > 0) return ew_launch_xpose<F>(plan, f, dS);
    }
    return ew_launch_strided<F, 1>(plan, f, plan.shape[0]);
}

template <class F>
dn_status ew_launch_xpose(const EwPlan &plan, const F &f, int dS) {
    constexpr int NOPS = F::NSRC + 1;
    // micro-tile: 4x4 elements, 2x2 when an 8-byte type is involved (vectors stay <= 16 bytes)
    constexpr int M = F::MaxSize >= 8 ? 2 : 4;
    // micro-tiles per thread: 8-byte types 4 (128 bytes per operand in flight); 4-byte types 2 for one-source
    // operators (copy a.T / abs(a.T): +3-4 %), 1 for two sources (2 costs a.T + b 2 %, a < b.T 3 %:
    // profiles/r02zg_ab_xpose.txt)
    constexpr int K = F::MaxSize >= 8 ? 4 : (F::NSRC == 1 ? 2 : 1);
    EwXposeParams<NOPS> p;
    const int nd = plan.ndims;
    p.sizeT = (uint32_t)plan.shape[0];
    p.sizeS = (uint32_t)plan.shape[dS];
    p.tilesT = (p.sizeT + 16 * M - 1) / (16 * M);
    p.tilesS = (p.sizeS + 16 * M * K - 1) / (16 * M * K);
    p.divTilesT.init(p.tilesT);
    p.divTilesS.init(p.tilesS);
    int nb = 0;
    int64_t nbatch = 1;
    for (int d = 0; d < DN_MAX_DIMS; ++d) {
        p.shapeB[d] = 1;
        p.divB[d].init(1);
        for (int k = 0; k < NOPS; ++k) p.strideB[k][d] = 0;
    }
    for (int d = 1; d < nd; ++d) {
        if (d == dS) continue;
        p.shapeB[nb] = (uint32_t)plan.shape[d];
        p.divB[nb].init((uint32_t)plan.shape[d]);
        for (int k = 0; k < NOPS; ++k) p.strideB[k][nb] = plan.op[k].stride[d] * plan.op[k].esize;
        nbatch *= plan.shape[d];
        ++nb;
    }
    p.nbatch_dims = nb;
    p.modes = 0;
    for (int k = 0; k < NOPS; ++k) {
        p.ptr[k] = plan.op[k].ptr;
        p.strideT[k] = plan.op[k].stride[0] * plan.op[k].esize;
        p.strideS[k] = plan.op[k].stride[dS] * plan.op[k].esize;
        p.modes |= (uint32_t)ew_xpose_mode(plan, k, dS, M) << (2 * k);
    }
    const int64_t tiles = (int64_t)p.tilesT * p.tilesS * nbatch;
    if (tiles >= ((int64_t)1 << 31)) return ew_launch_strided<F, 1>(plan, f, plan.shape[0]);
    DN_LAUNCH((ew_xpose_kernel<F, M, K>), (unsigned)tiles, kEwThreads, 0, p, f);
    return launch_status("element-wise transpose kernel");
}

}  // namespace dn
