// reduce.cuh — last-axis reductions: Sum/Product/Min/Max/All/Any/CountTrue, ArgMin/ArgMax, Find
// (Tensor/Tensor/TensorBackend.fs:125-135). Replaces Tensor/Tensor/Cuda/Kernels/Reduction.cuh:10-159, where ONE
// thread folds a whole row serially and adjacent threads read addresses a row apart (5.3 GB/s on K40c).
//
// Semantics follow the HOST backend (Tensor/Tensor/Host/ScalarOps.fs:606-665), see SURVEY.md §8c rules 4, 5, 8:
//   * ArgMin/ArgMax: strict compare, first occurrence, DN_NOT_FOUND when nothing beats the initial min/maxValue.
//   * Min/Max on floats: `if res > v then res else v` from the FINITE initial value; a NaN replaces the running
//     value and is itself replaced by the next element. This is an associative, order-sensitive monoid
//     (value after the last NaN, saw-a-NaN flag); both kernel families below partition the axis into CONTIGUOUS,
//     ORDERED pieces so it is reproduced exactly, not approximated.
//   * Float Sum/Product: fixed-shape tree (deterministic run to run), within rel 1e-4*log2(n) of the host's
//     left-to-right fold; integer folds wrap and are bit-exact in any order.
//
// Two kernel families, both HBM-bound by design:
//   rows  — the reduced axis is contiguous (stride 1). A warp streams a contiguous part of a row with 128-bit
//           loads (scalar head/tail peeled to 16-byte alignment), 4 loads in flight per lane, lane-interleaved
//           inside a 32*VEC round, rounds in index order. A row is 1 part (warp per row), 8 parts (CTA per row,
//           ordered combine in shared memory) or 8*S parts (S CTAs per row + finalize kernel) depending on R and L.
//   cols  — the reduced axis is strided. Threads run along the flattened OUTER index sorted by source stride (so a
//           warp reads consecutive addresses when some outer dim is contiguous) and each thread folds its chunk of
//           the axis sequentially; the axis is split over threadIdx.y / blockIdx.y with an ordered combine.
#pragma once

#include "common.cuh"
#include "ew_ops.cuh"
#include "peer.cuh"

namespace dn {

constexpr int kRedThreads = 256;
constexpr int kRedWarps = kRedThreads / 32;
constexpr unsigned kFull = 0xffffffffu;

// Outer (non-reduced) dims in canonical form, innermost-first.
struct RedOuter {
    int32_t ndims;
    uint32_t shape[DN_MAX_DIMS];
    FastDiv div[DN_MAX_DIMS];
    int64_t sstride[DN_MAX_DIMS];  // source, bytes
    int64_t tstride[DN_MAX_DIMS];  // target, bytes
};

__device__ __forceinline__ void red_offsets(const RedOuter &o, uint32_t r, int64_t &soff, int64_t &toff) {
    soff = 0;
    toff = 0;
    uint32_t rem = r;
#pragma unroll
    for (int d = 0; d < DN_MAX_DIMS; ++d) {
        if (d >= o.ndims) break;
        uint32_t q, x;
        if (d == o.ndims - 1) {
            x = rem;
            q = 0;
        } else {
            q = o.div[d].div(rem);
            x = rem - q * o.shape[d];
        }
        soff += (int64_t)x * o.sstride[d];
        toff += (int64_t)x * o.tstride[d];
        rem = q;
    }
}

struct RedParams {
    const char *src;
    char *dst;
    RedOuter outer;
    uint32_t nrows;       // outputs in this launch
    int64_t len;          // L
    int64_t lstride;      // bytes per step along the reduced axis (cols family)
    int32_t parts;        // rows: parts per row (1, 8 or 8*S); cols: chunks along the axis (blockDim.y*gridDim.y)
    int64_t part_len;     // elements per part / chunk
    void *partials;       // [nrows][ctas_per_row] states when more than one CTA shares a row
    int32_t ctas_per_row;
    PeerSync peer;        // sharded launches (shard.cu): every output is also stored into the peers' result buffers
    char *dst2;           // fused Min/Max + ArgMin/ArgMax: the int64 index target (same ELEMENT strides as dst)
    int32_t dst2_scale;   // byte offset in dst2 = byte offset in dst * dst2_scale (8 / sizeof(value type))
};

// The one place an output element is written: the local target and, in a sharded launch, the same offset of every
// peer's copy of the result (plain stores to peer-mapped memory over NVLink).
template <class Out>
__device__ __forceinline__ void red_store(const RedParams &p, char *dst, int64_t toff, Out v) {
    *reinterpret_cast<Out *>(dst + toff) = v;
    for (int k = 0; k < p.peer.npeers; ++k) *reinterpret_cast<Out *>(dst + toff + p.peer.delta[k]) = v;
}

// ---------------------------------------------------------------------------------------------------------------
// Operators.  State must be trivially copyable.  `ordered` ops get the per-round NaN vote in the rows family.
// ---------------------------------------------------------------------------------------------------------------
template <class T> __device__ __forceinline__ T shfl_xor_any(T v, int m) {
    static_assert(sizeof(T) % 4 == 0 || sizeof(T) < 4, "state size");
    if constexpr (sizeof(T) <= 4) {
        unsigned u = 0;
        memcpy(&u, &v, sizeof(T));
        u = __shfl_xor_sync(kFull, u, m);
        T r;
        memcpy(&r, &u, sizeof(T));
        return r;
    } else {
        constexpr int N = sizeof(T) / 4;
        unsigned u[N];
        memcpy(u, &v, sizeof(T));
#pragma unroll
        for (int i = 0; i < N; ++i) u[i] = __shfl_xor_sync(kFull, u[i], m);
        T r;
        memcpy(&r, u, sizeof(T));
        return r;
    }
}

template <class T> __device__ __forceinline__ T shfl_idx_any(T v, int src) {
    constexpr int N = (sizeof(T) + 3) / 4;
    unsigned u[N] = {};
    memcpy(u, &v, sizeof(T));
#pragma unroll
    for (int i = 0; i < N; ++i) u[i] = __shfl_sync(kFull, u[i], src);
    T r;
    memcpy(&r, u, sizeof(T));
    return r;
}

template <class T>
struct SumOp {
    using In = T; using Out = T; using State = T;
    static constexpr bool ordered = false;
    __device__ static State identity() { return T(0); }
    __device__ static void step(State &s, T v, int64_t) {
        if constexpr (kIsInt<T>) s = (T)((UnsignedT<T>)s + (UnsignedT<T>)v);
        else s = s + v;
    }
    __device__ static State combine(State a, State b) {
        if constexpr (kIsInt<T>) return (T)((UnsignedT<T>)a + (UnsignedT<T>)b);
        else return a + b;
    }
    __device__ static Out finalize(State s) { return s; }
    // 8- and 16-bit integers: the sixteen bytes of one 128-bit load through the dot-product unit (4 x IDP.4A /
    // IDP.2A with a vector of ones); the sum wraps in the element type, so any wider accumulator gives the same bits
    __device__ static void packed16(State &s, const uint32_t *w) {
        uint32_t acc = (uint32_t)(UnsignedT<T>)s;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if constexpr (sizeof(T) == 1) acc = __dp4a(w[i], 0x01010101u, acc);
            else acc = __dp2a_lo(w[i], 0x00000101u, acc);
        }
        s = (T)(UnsignedT<T>)acc;
    }
};

template <class T>
struct ProductOp {
    using In = T; using Out = T; using State = T;
    static constexpr bool ordered = false;
    __device__ static State identity() { return T(1); }
    __device__ static T mul(T a, T b) {
        if constexpr (kIsInt<T> && sizeof(T) < 4) return (T)((uint32_t)(UnsignedT<T>)a * (uint32_t)(UnsignedT<T>)b);
        else if constexpr (kIsInt<T>) return (T)((UnsignedT<T>)a * (UnsignedT<T>)b);
        else return a * b;
    }
    __device__ static void step(State &s, T v, int64_t) { s = mul(s, v); }
    __device__ static State combine(State a, State b) { return mul(a, b); }
    __device__ static Out finalize(State s) { return s; }
};

// Integer Min/Max: plain lattice, identity == the host's initial value.
template <class T, bool IsMax>
struct MinMaxIntOp {
    using In = T; using Out = T; using State = T;
    static constexpr bool ordered = false;
    __device__ static State identity() { return IsMax ? Limits<T>::lowest() : Limits<T>::max(); }
    __device__ static void step(State &s, T v, int64_t) { s = IsMax ? (s > v ? s : v) : (s < v ? s : v); }
    __device__ static State combine(State a, State b) { return IsMax ? (a > b ? a : b) : (a < b ? a : b); }
    __device__ static Out finalize(State s) { return s; }
};

// Float Min/Max with the host's NaN behaviour (see header).  val: fold over the elements after the last NaN
// (NaN itself if the piece ends with a NaN); flags bit0: piece contains a NaN, bit1: piece is non-empty.
template <class T, bool IsMax>
struct MinMaxFloatOp {
    using In = T; using Out = T;
    struct State { T val; int32_t flags; };
    static constexpr bool ordered = true;
    __device__ static T pos_inf() {
        if constexpr (std::is_same<T, float>::value) return __int_as_float(0x7f800000);
        else return __longlong_as_double(0x7ff0000000000000LL);
    }
    __device__ static T neutral() { return IsMax ? -pos_inf() : pos_inf(); }
    __device__ static bool better(T a, T b) { return IsMax ? a > b : a < b; }  // `res > v` / `res < v`
    // NaN-free operands only: one FMNMX; equals the host's pick in value (the sign of a +0/-0 tie may differ)
    __device__ static T pick(T a, T b) { return IsMax ? fmax(a, b) : fmin(a, b); }
    __device__ static State identity() { return State{neutral(), 0}; }
    // exact sequential step (ScalarOps.fs:620-628): if res `better` v then res else v
    __device__ static void step(State &s, T v, int64_t) {
        s.val = better(s.val, v) ? s.val : v;
        s.flags |= 2 | (v != v ? 1 : 0);
    }
    __device__ static State combine(State a, State b) {  // a covers earlier indices than b
        if (!(b.flags & 2)) return a;
        if (!(a.flags & 2)) return b;
        if (b.flags & 1) return State{b.val, 3};
        return State{better(a.val, b.val) ? a.val : b.val, a.flags};
    }
    __device__ static Out finalize(State s) {
        const T init = IsMax ? Limits<T>::lowest() : Limits<T>::max();
        if (s.flags & 1) return s.val;
        return better(init, s.val) ? init : s.val;
    }
};

template <bool IsAll>
struct AllAnyOp {
    using In = bool8; using Out = bool8; using State = int32_t;
    static constexpr bool ordered = false;
    __device__ static State identity() { return IsAll ? 1 : 0; }
    __device__ static void step(State &s, bool8 v, int64_t) { s = IsAll ? (s & (int)bool(v)) : (s | (int)bool(v)); }
    __device__ static State combine(State a, State b) { return IsAll ? (a & b) : (a | b); }
    __device__ static Out finalize(State s) { return bool8(s != 0); }
    // sixteen bools of one 128-bit load at once
    __device__ static void packed16(State &s, const uint32_t *w) {
        if constexpr (IsAll) s &= (int)((__vcmpeq4(w[0], 0u) | __vcmpeq4(w[1], 0u) | __vcmpeq4(w[2], 0u) | __vcmpeq4(w[3], 0u)) == 0u);
        else s |= (int)((w[0] | w[1] | w[2] | w[3]) != 0u);
    }
};

struct CountTrueOp {
    using In = bool8; using Out = int64_t; using State = int64_t;
    static constexpr bool ordered = false;
    __device__ static State identity() { return 0; }
    __device__ static void step(State &s, bool8 v, int64_t) { s += bool(v) ? 1 : 0; }
    __device__ static State combine(State a, State b) { return a + b; }
    __device__ static Out finalize(State s) { return s; }
    __device__ static void packed16(State &s, const uint32_t *w) {
        // per-byte 0/1 flags of the four words add without carry (at most 4 per byte)
        const uint32_t t = bool4_norm(w[0]) + bool4_norm(w[1]) + bool4_norm(w[2]) + bool4_norm(w[3]);
        s += (int64_t)__vsadu4(t, 0u);
    }
};

template <class Op> struct HasPacked16 : std::false_type {};
template <bool IsAll> struct HasPacked16<AllAnyOp<IsAll>> : std::true_type {};
template <> struct HasPacked16<CountTrueOp> : std::true_type {};
template <> struct HasPacked16<SumOp<int8_t>> : std::true_type {};
template <> struct HasPacked16<SumOp<uint8_t>> : std::true_type {};
template <> struct HasPacked16<SumOp<int16_t>> : std::true_type {};
template <> struct HasPacked16<SumOp<uint16_t>> : std::true_type {};

template <class T, bool IsMax>
struct ArgOp {
    using In = T; using Out = int64_t;
    struct State { T val; int64_t idx; };
    static constexpr bool ordered = false;
    __device__ static bool better(T a, T b) { return IsMax ? a > b : a < b; }
    // extremum of two candidates, NaNs ignored (a NaN never wins an arg fold): one FMNMX / IMNMX
    __device__ static T pick(T a, T b) {
        if constexpr (kIsFloat<T>) return IsMax ? fmax(a, b) : fmin(a, b);
        else return IsMax ? (a > b ? a : b) : (a < b ? a : b);
    }
    __device__ static State identity() {
        return State{IsMax ? Limits<T>::lowest() : Limits<T>::max(), (int64_t)DN_NOT_FOUND};
    }
    __device__ static void step(State &s, T v, int64_t i) {
        if (better(v, s.val)) { s.val = v; s.idx = i; }
    }
    __device__ static State combine(State a, State b) {
        if (better(b.val, a.val) || (b.val == a.val && b.idx < a.idx && b.idx != (int64_t)DN_NOT_FOUND)) return b;
        return a;
    }
    __device__ static Out finalize(State s) { return s.idx; }
};

template <class Op> struct IsArgOp : std::false_type {};
template <class T, bool IsMax> struct IsArgOp<ArgOp<T, IsMax>> : std::true_type {};

// Min/Max AND ArgMin/ArgMax of a float row in one pass: the two folds have different NaN rules (the value fold lets
// a NaN replace the running value, the arg fold never lets one win), so the state carries both.
template <class T, bool IsMax>
struct MinMaxArgOp {
    using In = T; using Out = T;
    using V = MinMaxFloatOp<T, IsMax>;
    using A = ArgOp<T, IsMax>;
    struct State { T val; int32_t flags; T aval; int64_t idx; };
    static constexpr bool ordered = true;
    __device__ static T neutral() { return V::neutral(); }
    __device__ static bool better(T a, T b) { return V::better(a, b); }
    __device__ static T pick(T a, T b) { return V::pick(a, b); }
    __device__ static State identity() {
        return State{V::neutral(), 0, IsMax ? Limits<T>::lowest() : Limits<T>::max(), (int64_t)DN_NOT_FOUND};
    }
    __device__ static void step(State &s, T v, int64_t i) {
        s.val = better(s.val, v) ? s.val : v;
        s.flags |= 2 | (v != v ? 1 : 0);
        if (better(v, s.aval)) { s.aval = v; s.idx = i; }
    }
    __device__ static State combine(State a, State b) {
        const typename V::State v = V::combine(typename V::State{a.val, a.flags}, typename V::State{b.val, b.flags});
        const typename A::State x = A::combine(typename A::State{a.aval, a.idx}, typename A::State{b.aval, b.idx});
        return State{v.val, v.flags, x.val, x.idx};
    }
    __device__ static Out finalize(State s) { return V::finalize(typename V::State{s.val, s.flags}); }
};
template <class Op> struct IsFusedArgOp : std::false_type {};
template <class T, bool IsMax> struct IsFusedArgOp<MinMaxArgOp<T, IsMax>> : std::true_type {};
template <class Op> struct IsMaxOp : std::false_type {};
template <class T, bool IsMax> struct IsMaxOp<ArgOp<T, IsMax>> : std::integral_constant<bool, IsMax> {};
template <class T, bool IsMax> struct IsMaxOp<MinMaxArgOp<T, IsMax>> : std::integral_constant<bool, IsMax> {};

// Writes the result(s) of one output element.
template <class Op>
__device__ __forceinline__ void red_emit(const RedParams &p, int64_t toff, const typename Op::State &s) {
    red_store<typename Op::Out>(p, p.dst, toff, Op::finalize(s));
    if constexpr (IsFusedArgOp<Op>::value) red_store<int64_t>(p, p.dst2, toff * p.dst2_scale, s.idx);
}

template <class T>
struct FindOp {
    using In = T; using Out = int64_t; using State = int64_t;
    static constexpr bool ordered = false;
    T value;
    __device__ static State identity() { return INT64_MAX; }
    __device__ void stepv(State &s, T v, int64_t i) const {
        bool eq;
        if constexpr (kIsBool<T>) eq = bool(v) == bool(value);
        else eq = v == value;
        if (eq && i < s) s = i;
    }
    __device__ static State combine(State a, State b) { return a < b ? a : b; }
    __device__ static Out finalize(State s) { return s == INT64_MAX ? (int64_t)DN_NOT_FOUND : s; }
};

// Uniform access to step (FindOp carries a runtime value, the others are stateless).
template <class Op>
__device__ __forceinline__ void op_step(const Op &op, typename Op::State &s, typename Op::In v, int64_t i) {
    if constexpr (std::is_same<Op, FindOp<typename Op::In>>::value) op.stepv(s, v, i);
    else Op::step(s, v, i);
}

template <class Op>
__device__ __forceinline__ typename Op::State warp_combine_unordered(typename Op::State s) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) s = Op::combine(s, shfl_xor_any(s, m));
    return s;
}

// Warp-wide combine of per-lane (best value, 32-bit index relative to the part's begin; -1 = none) pairs with
// first-occurrence semantics: two shuffles per step instead of the four a (value, int64 index) state needs.
template <class T, bool IsMax>
__device__ __forceinline__ void warp_arg_combine32(T &val, int &ridx) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        const T ov = shfl_xor_any(val, m);
        const int oi = __shfl_xor_sync(kFull, ridx, m);
        const bool better = IsMax ? ov > val : ov < val;
        if (oi >= 0 && (ridx < 0 || better || (ov == val && oi < ridx))) {
            val = ov;
            ridx = oi;
        }
    }
}

// The arg fold of an op (ArgOp itself, or the arg half of the fused value+index op): extremum of two candidates
// and the fold's start value.
template <class Op> struct ArgPick {
    using T = typename Op::In;
    __device__ static T pick(T a, T b) { return a; }
    __device__ static T lowest() { return T(); }
};
template <class T, bool IsMax> struct ArgPick<ArgOp<T, IsMax>> {
    __device__ static T pick(T a, T b) { return ArgOp<T, IsMax>::pick(a, b); }
    __device__ static T lowest() { return IsMax ? Limits<T>::lowest() : Limits<T>::max(); }
};
template <class T, bool IsMax> struct ArgPick<MinMaxArgOp<T, IsMax>> {
    __device__ static T pick(T a, T b) { return ArgOp<T, IsMax>::pick(a, b); }
    __device__ static T lowest() { return IsMax ? Limits<T>::lowest() : Limits<T>::max(); }
};

// ---------------------------------------------------------------------------------------------------------------
// rows family
// ---------------------------------------------------------------------------------------------------------------
// Min / Max of 1- and 2-byte integers on two 16-bit SIMD lanes (VIMNMX.S16x2 / .U16x2 are native; the 8-bit x4 forms
// are emulated): a lane keeps a packed 2 x 16-bit accumulator for the whole part; bytes are widened on the way in
// (one PRMT with sign replication / a zero byte per pair) — 1 instruction per byte, 0.5 per 16-bit element.
template <class Op> struct PackedMinMax { static constexpr bool value = false; };
template <class T, bool IsMax> struct PackedMinMax<MinMaxIntOp<T, IsMax>> {
    static constexpr bool value = sizeof(T) < 4;
    static constexpr bool kSigned = std::is_signed<T>::value;
    __device__ static uint32_t mnmx(uint32_t a, uint32_t b) {
        if constexpr (kSigned) return IsMax ? __vmaxs2(a, b) : __vmins2(a, b);
        else return IsMax ? __vmaxu2(a, b) : __vminu2(a, b);
    }
    __device__ static uint32_t identity() {  // the fold's start value in both 16-bit lanes
        const uint32_t v = (uint32_t)(uint16_t)(int16_t)MinMaxIntOp<T, IsMax>::identity();
        return v | (v << 16);
    }
    __device__ static uint32_t prmt(uint32_t a, uint32_t sel) {
        uint32_t d;
        asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(0u), "r"(sel));
        return d;
    }
    __device__ static void fold16(uint32_t &acc, const uint32_t *w) {  // the 16 bytes of one vector
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if constexpr (sizeof(T) == 2) {
                acc = mnmx(acc, w[i]);
            } else {  // bytes 0,2 and bytes 1,3 as 16-bit lanes (selector msb: replicate the byte's sign; 4: zero)
                acc = mnmx(acc, prmt(w[i], kSigned ? 0xA280u : 0x4240u));
                acc = mnmx(acc, prmt(w[i], kSigned ? 0xB391u : 0x4341u));
            }
        }
    }
    __device__ static T finish(uint32_t acc) {
        acc = mnmx(acc, acc >> 16);
        return (T)(uint16_t)(acc & 0xFFFFu);
    }
};

// The same for 1- and 2-byte elements (integers and bools; never ordered). A vector holds 8 or 16 elements, so the
// general structure below — per-element predicates in every ragged round, scalar loads assembled into vectors — costs
// ~200 registers (one per element in flight) and one CTA per SM. Here a lane only ever folds ONE element (head up to
// 16-byte alignment and the < VEC elements after the last whole vector, one per lane) or a WHOLE vector that stays
// packed in four registers: ops with a packed form (Sum through the dot-product unit, All / Any / CountTrue through
// byte-wise SIMD, Min / Max on 16-bit SIMD lanes) never unpack it. 262144 int8 rows of 1000: Sum 0.47 -> 3.35 TB/s,
// int16 Max 0.91 -> 6.56 (profiles/r02za_ab_reduce.txt).
template <class Op>
__device__ __forceinline__ typename Op::State warp_fold_part_subword(const Op &op, const char *row, int64_t begin,
                                                                    int64_t end, int lane) {
    using T = typename Op::In;
    using State = typename Op::State;
    static_assert(sizeof(T) < 4 && !Op::ordered, "sub-word integer / bool folds only");
    constexpr int VEC = 16 / (int)sizeof(T);
    // vectors in flight per lane: bytes that are folded one by one get unpacked into a register each
    constexpr int UNR = (HasPacked16<Op>::value || PackedMinMax<Op>::value || sizeof(T) == 2) ? 4 : 2;
    State st = Op::identity();
    int ridx = -1;  // arg ops: lane-local best index relative to `begin`
    uint32_t pacc = 0;  // Min / Max: packed 2 x 16-bit accumulator
    if constexpr (PackedMinMax<Op>::value) pacc = PackedMinMax<Op>::identity();
    const T *p = reinterpret_cast<const T *>(row);
    auto fold1 = [&](T v, int64_t i) {
        if constexpr (IsArgOp<Op>::value) {
            if (Op::better(v, st.val)) {
                st.val = v;
                ridx = (int)(i - begin);
            }
        } else {
            op_step(op, st, v, i);
        }
    };
    int64_t i = begin;
    if (i < end) {
        const uintptr_t addr = reinterpret_cast<uintptr_t>(p + i);
        int head = (int)(((16 - (addr & 15)) & 15) / sizeof(T));
        if (head > end - i) head = (int)(end - i);
        if (lane < head) fold1(p[i + lane], i + lane);
        i += head;
        const int64_t nvec = (end - i) / VEC;  // whole vectors; vector k belongs to lane k % 32
        const T *g = p + i;
        for (int64_t base = 0; base < nvec; base += 32 * UNR) {
            Pack<T, VEC> buf[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int64_t k = base + u * 32 + lane;
                if (k < nvec) buf[u] = load_pack<Pack<T, VEC>>(reinterpret_cast<const char *>(g + k * VEC));
            }
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int64_t k = base + u * 32 + lane;
                if (k < nvec) {
                    if constexpr (HasPacked16<Op>::value) {
                        Op::packed16(st, reinterpret_cast<const uint32_t *>(buf[u].v));
                    } else if constexpr (PackedMinMax<Op>::value) {
                        PackedMinMax<Op>::fold16(pacc, reinterpret_cast<const uint32_t *>(buf[u].v));
                    } else {
#pragma unroll
                        for (int j = 0; j < VEC; ++j) fold1(buf[u].v[j], i + k * VEC + j);
                    }
                }
            }
        }
        const int64_t ts = i + nvec * VEC;
        if (lane < (int)(end - ts)) fold1(p[ts + lane], ts + lane);
    }
    if constexpr (PackedMinMax<Op>::value) st = Op::combine(st, PackedMinMax<Op>::finish(pacc));
    if constexpr (IsArgOp<Op>::value) {
        warp_arg_combine32<T, IsMaxOp<Op>::value>(st.val, ridx);
        st.idx = ridx >= 0 ? begin + ridx : (int64_t)DN_NOT_FOUND;
        return st;
    } else {
        return warp_combine_unordered<Op>(st);
    }
}

// One warp folds elements [begin, end) of a contiguous row. All lanes return the same state.
// Structure: scalar head up to 16-byte alignment, then rounds of 32*VEC consecutive elements (lane l owns VEC
// consecutive elements of a round), UNR rounds loaded before any is folded; the last group is the same code with
// per-lane predicates (vector loads where a whole vector is in range, scalar loads for the final < VEC elements).
template <class Op>
__device__ __forceinline__ typename Op::State warp_fold_part(const Op &op, const char *row, int64_t begin,
                                                            int64_t end, int lane) {
    using T = typename Op::In;
    using State = typename Op::State;
    if constexpr (sizeof(T) < 4) return warp_fold_part_subword<Op>(op, row, begin, end, lane);
    constexpr int VEC = 16 / (int)sizeof(T);
    constexpr int ROUND = 32 * VEC;
    constexpr int UNR = 4;
    constexpr bool kVecArg = sizeof(T) >= 4;  // arg folds per vector (below)
    State st = Op::identity();
    int64_t last_nan = -1;  // ordered ops only: index of the last NaN seen by the WARP (uniform)
    int ridx = -1;          // arg ops only: lane-local best index relative to `begin`
    const T *p = reinterpret_cast<const T *>(row);

    // fold one round: this lane holds n (0..VEC) consecutive elements starting at index i0; FULL: n == VEC, known at
    // compile time (no per-element predicates — every round but the head and the last one of a part).
    // The arg folds work per VECTOR: the vector's extremum (NaNs dropped by FMNMX, as the strict compare drops them)
    // is compared with the lane's best and the lane remembers where the winning VECTOR starts — VEC-1 min/max plus
    // one compare and two selects per vector instead of a compare and two selects per element; which element of the
    // winning vector it was is settled once per part, after the warp-wide combine (resolve_arg below).
    auto round = [&](auto full_tag, const T *v, int n, int64_t i0) {
        constexpr bool FULL = decltype(full_tag)::value;
        auto vec_extremum = [&]() {
            using A = ArgPick<Op>;
            T m = FULL ? v[0] : A::lowest();
#pragma unroll
            for (int j = FULL ? 1 : 0; j < VEC; ++j)
                if (FULL || j < n) m = A::pick(m, v[j]);
            return m;
        };
        if constexpr (Op::ordered) {
            // cheap NaN probe: the sum of the lane's elements is NaN iff one of them is (or +inf meets -inf, a
            // false positive that the exact slow path below resolves)
            T probe = T(0);
#pragma unroll
            for (int j = 0; j < VEC; ++j)
                if (FULL || j < n) probe += v[j];
            T m = T(0);
            if constexpr (IsFusedArgOp<Op>::value) {  // the arg fold: NaNs never win (strict compare)
                m = vec_extremum();
                if (Op::better(m, st.aval)) {
                    st.aval = m;
                    ridx = (int)(i0 - begin);
                }
            }
            if (__any_sync(kFull, probe != probe)) {  // rare
                int64_t my_nan = -1;
#pragma unroll
                for (int j = 0; j < VEC; ++j)
                    if ((FULL || j < n) && v[j] != v[j]) my_nan = i0 + j;
#pragma unroll
                for (int m2 = 16; m2 >= 1; m2 >>= 1) {
                    const int64_t o = __shfl_xor_sync(kFull, my_nan, m2);
                    my_nan = o > my_nan ? o : my_nan;
                }
                if (my_nan >= 0) {  // a real NaN in this round resets every lane
                    last_nan = my_nan;
                    st.val = Op::neutral();
                }
#pragma unroll
                for (int j = 0; j < VEC; ++j)
                    if ((FULL || j < n) && i0 + j > last_nan) st.val = Op::better(st.val, v[j]) ? st.val : v[j];
            } else {
#pragma unroll
                for (int j = 0; j < VEC; ++j)
                    if (FULL || j < n) st.val = Op::pick(st.val, v[j]);
            }
        } else if constexpr (IsArgOp<Op>::value) {
            // (testing the vector's extremum first and searching the vector only on a hit was tried: the branch costs
            // more than the selects it saves — C3 ArgMax 6.42 -> 5.98 TB/s; the deferred search has no branch)
            if constexpr (kVecArg) {
                const T m = vec_extremum();
                if (Op::better(m, st.val)) {
                    st.val = m;
                    ridx = (int)(i0 - begin);
                }
            } else {  // sub-word elements: per element (the byte extraction dominates; the deferred search would add
                      // VEC scalar loads per part — int8 rows of 1000: 394 GB/s this way, 317 the other)
                const int r0 = (int)(i0 - begin);
#pragma unroll
                for (int j = 0; j < VEC; ++j)
                    if ((FULL || j < n) && Op::better(v[j], st.val)) {
                        st.val = v[j];
                        ridx = r0 + j;
                    }
            }
        } else {
            if constexpr (HasPacked16<Op>::value) {
                if constexpr (FULL) {  // v is a 16-byte aligned Pack
                    Op::packed16(st, reinterpret_cast<const uint32_t *>(v));
                    return;
                } else if (n == VEC) {
                    Op::packed16(st, reinterpret_cast<const uint32_t *>(v));
                    return;
                }
            }
            if constexpr (!(FULL && HasPacked16<Op>::value)) {
#pragma unroll
                for (int j = 0; j < VEC; ++j)
                    if (FULL || j < n) op_step(op, st, v[j], i0 + j);
            }
        }
    };
    // arg folds: `ridx` is where the winning vector starts (relative to `begin`); the element is the first one of
    // that vector equal to the winning value (a NaN equals nothing; elements past the vector belong to later
    // vectors and are only reached if nothing before them matched, which cannot happen). Uniform across the warp.
    auto resolve_arg = [&](T best, int r) -> int64_t {
        if (r < 0) return (int64_t)DN_NOT_FOUND;
        if constexpr (!kVecArg) return begin + r;
        int first = 0;
#pragma unroll
        for (int j = VEC - 1; j >= 0; --j) {
            const int64_t e = begin + r + j;
            if (e < end && p[e] == best) first = j;
        }
        return begin + r + first;
    };
    const std::true_type kFullRound{};
    const std::false_type kPartRound{};

    int64_t i = begin;
    if (i < end) {
        const uintptr_t addr = reinterpret_cast<uintptr_t>(p + i);
        int head = (int)(((16 - (addr & 15)) & 15) / sizeof(T));
        if (head > end - i) head = (int)(end - i);
        if (head > 0) {
            T v[VEC];
            const bool on = lane < head;
            if (on) v[0] = p[i + lane];
            round(kPartRound, v, on ? 1 : 0, i + lane);
            i += head;
        }
        // groups of UNR rounds; only the last group of a part can be partial (per-lane predicates, scalar loads for
        // the final < VEC elements); offsets inside a group are 32-bit
        for (; i < end; i += (int64_t)ROUND * UNR) {
            const int64_t rem64 = end - i;
            const int rem = rem64 > (int64_t)ROUND * UNR ? ROUND * UNR : (int)rem64;  // elements left in this group
            const T *g = p + i;
            Pack<T, VEC> buf[UNR];
            if (rem == ROUND * UNR) {  // full group: UNR unconditional 128-bit loads back to back
#pragma unroll
                for (int u = 0; u < UNR; ++u)
                    buf[u] = load_pack<Pack<T, VEC>>(reinterpret_cast<const char *>(g + u * ROUND + lane * VEC));
            } else {
#pragma unroll
                for (int u = 0; u < UNR; ++u) {
                    const int off = u * ROUND + lane * VEC;
                    if (off + VEC <= rem) {
                        buf[u] = load_pack<Pack<T, VEC>>(reinterpret_cast<const char *>(g + off));
                    } else {
#pragma unroll
                        for (int j = 0; j < VEC; ++j)
                            if (off + j < rem) buf[u].v[j] = g[off + j];
                    }
                }
            }
            if (rem == ROUND * UNR) {
#pragma unroll
                for (int u = 0; u < UNR; ++u) round(kFullRound, buf[u].v, VEC, i + u * ROUND + lane * VEC);
            } else {
#pragma unroll
                for (int u = 0; u < UNR; ++u) {
                    const int off = u * ROUND + lane * VEC;
                    if ((u + 1) * ROUND <= rem) {  // warp-uniform: a whole round
                        round(kFullRound, buf[u].v, VEC, i + off);
                    } else if (u * ROUND < rem) {  // warp-uniform: the part's last, ragged round
                        const int left = rem - off;
                        round(kPartRound, buf[u].v, left >= VEC ? VEC : (left > 0 ? left : 0), i + off);
                    }
                }
            }
        }
    }
    if constexpr (Op::ordered) {
        T v = st.val;
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            const T o = __shfl_xor_sync(kFull, v, m);
            v = Op::pick(v, o);
        }
        State out;
        if constexpr (IsFusedArgOp<Op>::value) {
            T av = st.aval;
            warp_arg_combine32<T, IsMaxOp<Op>::value>(av, ridx);
            out.aval = av;
            out.idx = resolve_arg(av, ridx);
        }
        out.flags = (end > begin ? 2 : 0) | (last_nan >= 0 ? 1 : 0);
        out.val = v;
        if (last_nan >= 0 && last_nan == end - 1) out.val = p[last_nan];  // piece ends with the NaN itself
        return out;
    } else {
        if constexpr (IsArgOp<Op>::value) {
            warp_arg_combine32<T, IsMaxOp<Op>::value>(st.val, ridx);
            st.idx = resolve_arg(st.val, ridx);
            return st;
        } else {
            return warp_combine_unordered<Op>(st);
        }
    }
}

// Register target of the rows kernel: the folds over 4- and 8-byte elements need <= 64 registers and run best at
// exactly that (4 CTAs per SM: the loads of a group stay in flight; left alone ptxas squeezes them to 48 registers
// for a fifth CTA and serialises the loads — measured 5-7 % slower on C3). Byte-wise folds that unpack their
// vectors (ArgMin/ArgMax, Find, Product of 1-byte elements) get 85; the fused value+index fold fits with 12 bytes of spill and is 24 % faster for it (C3 Max+ArgMax in one pass:
// 4.56 -> 5.67 TB/s, same box, `profiles/r02w_ab_reduce.txt`).
template <class Op> struct RowsMinBlocks {
    static constexpr int value = (sizeof(typename Op::In) == 1 && !HasPacked16<Op>::value && !PackedMinMax<Op>::value) ? 3 : 4;
};
template <class T, bool IsMax> struct RowsMinBlocks<MinMaxArgOp<T, IsMax>> { static constexpr int value = 4; };

// parts == 1: warp per row.  parts == 8*S: one CTA per (row, s); its 8 warps take consecutive parts.
template <class Op>
__global__ void __launch_bounds__(kRedThreads, RowsMinBlocks<Op>::value) reduce_rows_kernel(const __grid_constant__ RedParams p, const Op op) {
    using State = typename Op::State;
    using Out = typename Op::Out;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ State sm[kRedWarps];
    if (p.parts == 1 && p.peer.npeers > 0) {
        // Sharded launch: every warp owns a CONTIGUOUS range of rows and parks the result of row j of the current
        // group of 32 in lane j, so that the results leave as one coalesced store per group — to the local target and
        // to every peer's copy. (One 8-byte store per row and peer, issued by a single lane, makes every output its own
        // NVLink write transaction: measured 26 us on top of an 84 us kernel at 2 GPUs.)
        const uint64_t nwarps = (uint64_t)gridDim.x * kRedWarps, gw = (uint64_t)blockIdx.x * kRedWarps + warp;
        const uint64_t per = (p.nrows + nwarps - 1) / nwarps;
        const uint64_t r_begin = gw * per;
        uint64_t r_end = r_begin + per;
        if (r_end > p.nrows) r_end = p.nrows;
        for (uint64_t g0 = r_begin; g0 < r_end; g0 += 32) {
            const int cnt = r_end - g0 < 32 ? (int)(r_end - g0) : 32;
            Out keep = Out();
            int64_t keep_idx = 0, keep_toff = 0;
            for (int j = 0; j < cnt; ++j) {
                int64_t soff, toff;
                red_offsets(p.outer, (uint32_t)(g0 + j), soff, toff);
                const State s = warp_fold_part<Op>(op, p.src + soff, 0, p.len, lane);
                if (lane == j) {
                    keep = Op::finalize(s);
                    keep_toff = toff;
                    if constexpr (IsFusedArgOp<Op>::value) keep_idx = s.idx;
                }
            }
            if (lane < cnt) {
                *reinterpret_cast<Out *>(p.dst + keep_toff) = keep;
                if constexpr (IsFusedArgOp<Op>::value) *reinterpret_cast<int64_t *>(p.dst2 + keep_toff * p.dst2_scale) = keep_idx;
                for (int k = 0; k < p.peer.npeers; ++k) {
                    *reinterpret_cast<Out *>(p.dst + keep_toff + p.peer.delta[k]) = keep;
                    if constexpr (IsFusedArgOp<Op>::value)
                        *reinterpret_cast<int64_t *>(p.dst2 + keep_toff * p.dst2_scale + p.peer.delta[k]) = keep_idx;
                }
            }
        }
    } else if (p.parts == 1) {
        for (uint64_t r = (uint64_t)blockIdx.x * kRedWarps + warp; r < p.nrows; r += (uint64_t)gridDim.x * kRedWarps) {
            int64_t soff, toff;
            red_offsets(p.outer, (uint32_t)r, soff, toff);
            State s = warp_fold_part<Op>(op, p.src + soff, 0, p.len, lane);
            if (lane == 0) {
                *reinterpret_cast<Out *>(p.dst + toff) = Op::finalize(s);
                if constexpr (IsFusedArgOp<Op>::value) *reinterpret_cast<int64_t *>(p.dst2 + toff * p.dst2_scale) = s.idx;
            }
        }
    } else {
        const uint64_t total = (uint64_t)p.nrows * p.ctas_per_row;
        for (uint64_t w = blockIdx.x; w < total; w += gridDim.x) {
            const uint32_t r = (uint32_t)(w / p.ctas_per_row);
            const int s_idx = (int)(w % p.ctas_per_row);
            int64_t soff, toff;
            red_offsets(p.outer, r, soff, toff);
            const int64_t part = (int64_t)s_idx * kRedWarps + warp;
            int64_t b = part * p.part_len, e = b + p.part_len;
            if (b > p.len) b = p.len;
            if (e > p.len) e = p.len;
            State s = warp_fold_part<Op>(op, p.src + soff, b, e, lane);
            if (lane == 0) sm[warp] = s;
            __syncthreads();
            if (threadIdx.x == 0) {
                State acc = sm[0];
#pragma unroll
                for (int k = 1; k < kRedWarps; ++k) acc = Op::combine(acc, sm[k]);
                if (p.ctas_per_row == 1) red_emit<Op>(p, toff, acc);
                else reinterpret_cast<State *>(p.partials)[(uint64_t)r * p.ctas_per_row + s_idx] = acc;
            }
            __syncthreads();
        }
    }
    peer_exit(p.peer);
}

// Ordered combine of the per-CTA partial states of each row; one thread per row.
template <class Op>
__global__ void __launch_bounds__(kRedThreads) reduce_finalize_kernel(const __grid_constant__ RedParams p) {
    using State = typename Op::State;
    using Out = typename Op::Out;
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < p.nrows) {
        const State *part = reinterpret_cast<const State *>(p.partials) + r * p.ctas_per_row;
        State acc = part[0];
        for (int k = 1; k < p.ctas_per_row; ++k) acc = Op::combine(acc, part[k]);
        int64_t soff, toff;
        red_offsets(p.outer, (uint32_t)r, soff, toff);
        red_emit<Op>(p, toff, acc);
    }
    peer_exit(p.peer);
}

// The same for rows shared by MANY CTAs (whole-tensor folds: up to 1024 partials per row): one warp per row, lane l
// folds the contiguous block l of the partials (independent loads in flight), the 32 lane results are combined in
// lane order — still an ordered fold, without a chain of ~1000 dependent L2 round trips in one thread.
template <class Op>
__global__ void __launch_bounds__(kRedThreads) reduce_finalize_warp_kernel(const __grid_constant__ RedParams p) {
    using State = typename Op::State;
    using Out = typename Op::Out;
    const int lane = threadIdx.x & 31;
    const uint64_t r = (uint64_t)blockIdx.x * kRedWarps + (threadIdx.x >> 5);
    if (r < p.nrows) {  // warp-uniform
        const State *part = reinterpret_cast<const State *>(p.partials) + r * p.ctas_per_row;
        const int per = (p.ctas_per_row + 31) / 32;
        const int b = lane * per;
        int e = b + per;
        if (e > p.ctas_per_row) e = p.ctas_per_row;
        State acc = Op::identity();
        bool any = false;
        for (int k = b; k < e; ++k) {
            const State v = part[k];
            acc = any ? Op::combine(acc, v) : v;
            any = true;
        }
        // ordered combine over the lanes that hold something (lane order == partial order)
        State total = shfl_idx_any(acc, 0);
        const int used = (p.ctas_per_row + per - 1) / per;
        for (int l = 1; l < used; ++l) total = Op::combine(total, shfl_idx_any(acc, l));
        if (lane == 0) {
            int64_t soff, toff;
            red_offsets(p.outer, (uint32_t)r, soff, toff);
            red_emit<Op>(p, toff, total);
        }
    }
    peer_exit(p.peer);
}

// ---------------------------------------------------------------------------------------------------------------
// cols family: blockDim = (TX, TY); thread (x, y) folds chunk y of output x sequentially.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kColsTX = 32;

// VEC > 1: the innermost outer dim is contiguous in the source (stride 1), its extent is a multiple of VEC and
// everything is 16-byte aligned, so a thread owns VEC adjacent outputs and reads them with one 128-bit load per
// step of the reduced axis (4 steps in flight).
template <class Op, int VEC>
__global__ void __launch_bounds__(kRedThreads) reduce_cols_kernel(const __grid_constant__ RedParams p, const Op op) {
    using T = typename Op::In;
    using State = typename Op::State;
    using Out = typename Op::Out;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    State *sm = reinterpret_cast<State *>(smem_raw);  // [blockDim.y][blockDim.x][VEC]
    const uint64_t o0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;  // first output of this thread
    const int chunk = blockIdx.y * blockDim.y + threadIdx.y;
    State st[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) st[j] = Op::identity();
    int64_t soff = 0, toff = 0;
    const bool active = o0 < p.nrows;
    if (active) {
        red_offsets(p.outer, (uint32_t)o0, soff, toff);
        int64_t b = (int64_t)chunk * p.part_len, e = b + p.part_len;
        if (e > p.len) e = p.len;
        const char *q = p.src + soff + b * p.lstride;
        int64_t i = b;
        for (; i + 4 <= e; i += 4) {
            Pack<T, VEC> v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = load_pack<Pack<T, VEC>>(q + u * p.lstride);
            q += 4 * p.lstride;
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int j = 0; j < VEC; ++j) op_step(op, st[j], v[u].v[j], i + u);
        }
        for (; i < e; ++i) {
            const Pack<T, VEC> v = load_pack<Pack<T, VEC>>(q);
            q += p.lstride;
#pragma unroll
            for (int j = 0; j < VEC; ++j) op_step(op, st[j], v.v[j], i);
        }
    }
    bool writer = active;
    if (blockDim.y > 1) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) sm[(threadIdx.y * blockDim.x + threadIdx.x) * VEC + j] = st[j];
        __syncthreads();
        writer = active && threadIdx.y == 0;
        if (writer)
            for (int y = 1; y < (int)blockDim.y; ++y)
#pragma unroll
                for (int j = 0; j < VEC; ++j) st[j] = Op::combine(st[j], sm[(y * blockDim.x + threadIdx.x) * VEC + j]);
    }
    if (writer) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            if (gridDim.y == 1) red_emit<Op>(p, toff + j * p.outer.tstride[0], st[j]);
            else reinterpret_cast<State *>(p.partials)[(o0 + j) * gridDim.y + blockIdx.y] = st[j];
        }
    }
    peer_exit(p.peer);
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
struct RedPlan {
    const char *src;
    char *dst;
    int64_t len, lstride_elems;
    int nouter;
    int64_t oshape[DN_MAX_DIMS], ostride_s[DN_MAX_DIMS], ostride_t[DN_MAX_DIMS];  // elements, innermost-first
    int64_t nrows;
    int in_size, out_size;
    char *dst2 = nullptr;  // fused Min/Max + Arg only
};

// Validates shapes (a: [..., L], t: [...]) and canonicalises the outer dims (drop size-1, sort by source stride,
// merge). nrows == 0 means nothing to do.
dn_status red_make_plan(RedPlan &plan, const dn_tensor *t, const dn_tensor *a, const char *what);
void red_fill_outer(RedOuter &o, const RedPlan &plan);
// True when a thread of the cols family may own `vec` adjacent outputs and read them with one vector load.
bool red_cols_can_vectorize(const RedPlan &plan, int vec);

template <class Op>
dn_status red_run(const RedPlan &plan, const Op &op) {
    using State = typename Op::State;
    using Out = typename Op::Out;
    if (plan.nrows == 0) return DN_OK;
    const int sms = sm_count();
    const int64_t L = plan.len;
    // rows family needs a contiguous reduced axis; tiny axes are better served by the cols family
    const bool rows_family = (plan.lstride_elems == 1 || L <= 1) && L >= 64;
    if (plan.nrows >= ((int64_t)1 << 31))
        return set_error(DN_ERR_UNSUPPORTED, "reduction with more than 2^31-1 outputs is not supported");
    {
        const int64_t rc = plan.nrows;
        RedParams p;
        red_fill_outer(p.outer, plan);
        p.src = plan.src;
        p.dst = plan.dst;
        p.nrows = (uint32_t)rc;
        p.len = L;
        p.lstride = plan.lstride_elems * plan.in_size;
        p.partials = nullptr;
        p.ctas_per_row = 1;
        void *scratch = nullptr;
        // sharded launch (shard.cu): the kernel that STORES the outputs also stores them into the peers' buffers
        // and runs the exit barrier; a kernel that only produces partials runs plain
        PeerSync sync = {};
        peer_take(sync);
        p.peer = PeerSync{};
        p.dst2 = plan.dst2;
        p.dst2_scale = plan.in_size ? 8 / plan.in_size : 1;
        if (rows_family) {
            const int64_t warps_wanted = (int64_t)sms * 32;  // enough resident warps to cover HBM latency
            const int64_t bytes = L * plan.in_size;
            // short rows: a warp per row (one or two load groups cover the row). Long rows: a CTA per row, so that
            // a row's latency chain is 1/8 as long and small row counts still fill the machine.
            // (measured: with >= one warp per resident slot and rows <= 64 KiB, the state-heavy ops — ordered
            // float Min/Max, (value, index) pairs — run faster warp-per-row; plain folds prefer CTA-per-row)
            // Warp-per-row leaves the machine partly idle in its last wave of rows: with long rows it needs at
            // least ~4 waves of resident warps (64 per SM) to keep that tail below ~10 % (measured: [16384,16384]
            // f32 Max = 1.7 waves ran at 86 % of peak).
            const bool heavy = Op::ordered || sizeof(State) > 8;
            const bool enough_waves = rc >= (int64_t)sms * 64 * 4 || bytes <= 16384;
            // (bool rows of 16 KiB — All / Any / CountTrue over [16384,16384] — run at 57-65 % of peak CTA-per-row;
            // warp-per-row was tried for them and is worse, 35-39 %: too few warps with too long a chain each)
            if (bytes <= 4096 || (heavy && rc >= warps_wanted && bytes <= 65536 && enough_waves)) {
                p.parts = 1;
                p.part_len = L;
                int64_t ctas = (rc + kRedWarps - 1) / kRedWarps;
                int64_t cap = (int64_t)sms * 16;
                if (sync.nflags > 0) {
                    // sharded launch: ONE wave of CTAs. Every CTA ends with a system-scope fence that waits for its
                    // peer stores to be acknowledged across NVLink (microseconds); with several waves that latency is
                    // paid once per wave (measured at 8 GPUs: 48 us on top of a 21 us kernel with 4 waves).
                    static const int occ = [] {
                        int nb = 0;
                        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, reduce_rows_kernel<Op>, kRedThreads, 0) != cudaSuccess || nb < 1) nb = 2;
                        return nb;
                    }();
                    cap = (int64_t)sms * occ;
                }
                if (ctas > cap) ctas = cap;
                p.peer = sync;
                DN_LAUNCH((reduce_rows_kernel<Op>), (unsigned)ctas, kRedThreads, 0, p, op);
            } else {
                // S CTAs per row so that rc*S CTAs fill the machine; every part stays >= 4 KiB
                int64_t S = (6 * (int64_t)sms + rc - 1) / rc;
                const int64_t max_s = bytes / (kRedWarps * 4096);
                if (S > max_s) S = max_s;
                if (S < 1) S = 1;
                if (S > 1024) S = 1024;
                const int64_t nparts = S * kRedWarps;
                constexpr int64_t kAlignElems = 512;  // keep part boundaries 16-byte aligned relative to the row
                int64_t part_len = (L + nparts - 1) / nparts;
                part_len = (part_len + kAlignElems - 1) / kAlignElems * kAlignElems;
                p.parts = (int32_t)nparts;
                p.part_len = part_len;
                p.ctas_per_row = (int32_t)S;
                if (S > 1) {
                    dn_status st = scratch_alloc((size_t)rc * S * sizeof(State), &scratch);
                    if (st != DN_OK) return st;
                    p.partials = scratch;
                }
                int64_t ctas = rc * S;
                const int64_t cap = (int64_t)sms * 16;
                if (ctas > cap) ctas = cap;
                if (S == 1) p.peer = sync;
                DN_LAUNCH((reduce_rows_kernel<Op>), (unsigned)ctas, kRedThreads, 0, p, op);
                p.peer = sync;
                if (S >= 32) {
                    DN_LAUNCH((reduce_finalize_warp_kernel<Op>), (unsigned)((rc + kRedWarps - 1) / kRedWarps), kRedThreads, 0, p);
                } else if (S > 1) {
                    DN_LAUNCH((reduce_finalize_kernel<Op>), (unsigned)((rc + kRedThreads - 1) / kRedThreads),
                              kRedThreads, 0, p);
                }
            }
        } else {
            // cols family: TX threads along outputs; TY*GY chunks along the axis when outputs alone cannot fill
            // the machine.
            constexpr int kColsVec = 16 / (int)sizeof(typename Op::In) > 4 ? 4 : 16 / (int)sizeof(typename Op::In);
            const int vec = red_cols_can_vectorize(plan, kColsVec) ? kColsVec : 1;
            int tx = kColsTX, ty = kRedThreads / kColsTX;
            const int64_t threads_x = (rc + vec - 1) / vec;
            int64_t gx = (threads_x + tx - 1) / tx;
            int64_t gy = 1;
            if (L < 64) {  // short axis: no point in splitting it
                tx = kRedThreads;
                ty = 1;
                gx = (threads_x + tx - 1) / tx;
            } else if (gx < 4 * sms) {
                gy = (4 * (int64_t)sms + gx - 1) / gx;
                const int64_t max_gy = L / (ty * 64) > 0 ? L / (ty * 64) : 1;
                if (gy > max_gy) gy = max_gy;
                if (gy > 65535) gy = 65535;
            }
            const int64_t chunks = (int64_t)ty * gy;
            p.parts = (int32_t)chunks;
            p.part_len = (L + chunks - 1) / chunks;
            if (p.part_len < 1) p.part_len = 1;
            p.ctas_per_row = (int32_t)gy;
            if (gy > 1) {
                dn_status st = scratch_alloc((size_t)rc * gy * sizeof(State), &scratch);
                if (st != DN_OK) return st;
                p.partials = scratch;
            }
            const size_t smem = ty > 1 ? (size_t)tx * ty * sizeof(State) * vec : 0;
            if (gy == 1) p.peer = sync;
            if (vec > 1)
                DN_LAUNCH((reduce_cols_kernel<Op, kColsVec>), dim3((unsigned)gx, (unsigned)gy), dim3(tx, ty), smem, p, op);
            else
                DN_LAUNCH((reduce_cols_kernel<Op, 1>), dim3((unsigned)gx, (unsigned)gy), dim3(tx, ty), smem, p, op);
            if (gy > 1) {
                p.peer = sync;
                DN_LAUNCH((reduce_finalize_kernel<Op>), (unsigned)((rc + kRedThreads - 1) / kRedThreads),
                          kRedThreads, 0, p);
            }
        }
        scratch_free(scratch);
        (void)sizeof(Out);
    }
    return launch_status("reduction kernel");
}

}  // namespace dn
