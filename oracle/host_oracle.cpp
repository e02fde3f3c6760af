// host_oracle.cpp — CPU restatement of Deep.Net's HostTensor backend.  TEST INFRASTRUCTURE ONLY.
//
// This file is the parity oracle for libdeepnet_b200.so. It is NOT part of the product: only tests/, the smoke
// check in __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs may load it. The product path
// (deepnet_b200/) never links, imports or calls anything in oracle/.
//
// What it restates (all citations relative to the reference tree /root/reference):
//   dispatch + layout swap ........ Tensor/Tensor/Host/HostBackend.fs:139-153,182-461
//   loop drivers + threading ...... Tensor/Tensor/Host/ScalarOps.fs:15-361   (Parallel.For over dim 0 iff nd > 1)
//   operator bodies ............... Tensor/Tensor/Host/ScalarOps.fs:363-707
//   SIMD path selection ........... Tensor/Tensor/Host/VectorOps.fs:34-240,287-301 (single-threaded)
//   addressing / iteration order .. Tensor/Tensor/Host/FastAccess.fs:20-123 (PosIter32 = logical row-major order)
//   scalar semantics .............. Tensor/Tensor/ScalarPrimitives.fs:50-246, Tensor/Tensor/Sgn.fs:8-31
//   initial values ................ Tensor/Tensor/Utils.fs:226-299, NotFound Tensor/Tensor/TensorRng.fs:24
//
// The reference itself is F#/.NET and cannot be compiled or run in this image (no dotnet/mono/fsharpc), so this is
// a restatement, pinned against the reference's documented known answers (tests/golden/doc_kats.json, transcribed
// from the XML-doc examples in Tensor/Tensor/Tensor.fs and Tensor.Docs/articles/Guide-*.md).
// PARITY UNPINNED for: the transcendental functions' last-ulp behaviour (CoreCLR System.Math vs glibc libm — only the
// 4-digit print at Guide-Operations.md:89 pins them) and GEMM rounding (MKL blob missing from the reference tree).
//
// Arithmetic rules restated here that differ from a naive C port:
//   1. f32 Log..Tanh, Power, Ceiling/Floor/Truncate/Round are evaluated in double and rounded to float
//      (FSharp.Core float32 operators call System.Math on the widened value; ScalarPrimitives.fs:72-155,177-181).
//   2. Round is half-to-even (Math.Round; ScalarPrimitives.fs:147-150).
//   3. Sgn(NaN) = 0 (Sgn.fs:14-15); Sgn exists only for i16/i32/i64/f32/f64.
//   4. ArgMin/ArgMax: strict compare, first occurrence, initial (NotFound, min/maxValue) (ScalarOps.fs:638-654).
//   5. Min/MaxLastAxis: `if res > v then res else v` from the FINITE initial value (ScalarOps.fs:620-628); a NaN
//      replaces the state and is replaced by the next element. Max/MinElemwise: (a>b)?a:b / (a<b)?a:b.
//   6. Integer ops wrap; integer / and % truncate; x/0 and MinValue/-1 throw on the host and are outside the
//      parity domain (this oracle returns 0 / wraps instead of trapping). Contiguous integer Abs wraps at MinValue
//      (Vector.Abs; VectorOps.fs:209-211); UnaryMinus on unsigned wraps (Vector.Negate on the SIMD path).
//   7. Convert is an unchecked cast (ScalarPrimitives.fs:50-52); float->int truncates; out-of-range is outside
//      the parity domain. bool<->number follows the reference CUDA kernel's static_cast (host LINQ has no such cast).
//   8. Sum/Product fold strictly left to right in the element type (ScalarOps.fs:610-618).
//   9. Scatter adds in logical row-major source order, single-threaded (ScalarOps.fs:593-604); MaskedGet/Set and
//      TrueIndices walk in logical row-major order (ScalarOps.fs:667-707).
//
// Build: see oracle/Makefile (g++ -O2 -fopenmp -ffp-contract=off; no -ffast-math).

#include "../include/dn_tensor.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <omp.h>
#include <limits>
#include <type_traits>
#include <vector>

namespace {

thread_local char g_err[512] = "";

dn_status fail(dn_status st, const char *msg) {
    std::snprintf(g_err, sizeof g_err, "%s", msg);
    return st;
}

constexpr int64_t kNotFound = INT64_MIN + 4;

// Threading policy switch: 1 = restate the reference's policy (default), 0 = force single thread.
int g_threads_enabled = 1;

// ---------------------------------------------------------------------------------------------------------------
// Typed strided view over host memory (FastLayout32 + data array; FastAccess.fs:20-61).
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
struct View {
    T *p;  // already offset
    int nd;
    int64_t shape[DN_MAX_DIMS];
    int64_t stride[DN_MAX_DIMS];
    int64_t nelems() const {
        int64_t n = 1;
        for (int d = 0; d < nd; ++d) n *= shape[d];
        return n;
    }
};

template <typename T>
View<T> view_of(const dn_tensor *t) {
    View<T> v;
    v.p = static_cast<T *>(t->base) + t->offset;
    v.nd = t->ndims;
    for (int d = 0; d < DN_MAX_DIMS; ++d) {
        v.shape[d] = d < t->ndims ? t->shape[d] : 1;
        v.stride[d] = d < t->ndims ? t->stride[d] : 0;
    }
    return v;
}

bool same_shape(const dn_tensor *a, const dn_tensor *b) {
    if (a->ndims != b->ndims) return false;
    for (int d = 0; d < a->ndims; ++d)
        if (a->shape[d] != b->shape[d]) return false;
    return true;
}

bool valid(const dn_tensor *t) {
    if (!t || t->ndims < 0 || t->ndims > DN_MAX_DIMS) return false;
    if (t->dtype < 0 || t->dtype >= DN_DTYPE_COUNT) return false;
    for (int d = 0; d < t->ndims; ++d)
        if (t->shape[d] < 0) return false;
    return true;
}

// HostBackend.ElemwiseLayouts (HostBackend.fs:139-153): pick the largest dim whose target stride is 1 and whose
// source strides are all 0/1 and swap it to the last position. Returns the dim chosen or -1.
int elemwise_best_last_dim(int nd, const int64_t *tshape, const int64_t *tstride, const int64_t *const *sstrides,
                           int nsrc) {
    int best = -1;
    int64_t best_size = -1;
    for (int d = 0; d < nd; ++d) {
        bool good = tstride[d] == 1;
        for (int s = 0; s < nsrc && good; ++s) good = sstrides[s][d] == 1 || sstrides[s][d] == 0;
        // List.maxBy keeps the FIRST maximal element.
        if (good && tshape[d] > best_size) {
            best = d;
            best_size = tshape[d];
        }
    }
    return best;
}

template <typename V>
void swap_dims(V &v, int a, int b) {
    std::swap(v.shape[a], v.shape[b]);
    std::swap(v.stride[a], v.stride[b]);
}

// ---------------------------------------------------------------------------------------------------------------
// Loop drivers.  ScalarOps.Apply*Op (ScalarOps.fs:15-230): logical row-major walk; Parallel.For over dim 0 only
// when useThreads && nd > 1.  `body(pos, offsets...)` is called once per element.
// ---------------------------------------------------------------------------------------------------------------
template <int N, typename Body>
void walk_rows(int nd, const int64_t *shape, const int64_t *const (&strides)[N], int64_t dim0_from, int64_t dim0_to,
               Body &&body) {
    // iterate positions with dim 0 restricted to [dim0_from, dim0_to)
    int64_t pos[DN_MAX_DIMS] = {0};
    int64_t off[N];
    if (nd == 0) {
        for (int k = 0; k < N; ++k) off[k] = 0;
        body(pos, off);
        return;
    }
    for (int d = 0; d < nd; ++d)
        if (shape[d] == 0) return;
    if (dim0_from >= dim0_to) return;
    pos[0] = dim0_from;
    for (int k = 0; k < N; ++k) off[k] = dim0_from * strides[k][0];
    const int64_t inner = shape[nd - 1];
    while (true) {
        // inner loop over last dim (when nd == 1 the last dim IS dim 0 and is bounded by [from,to))
        int64_t i0 = nd == 1 ? dim0_from : 0, i1 = nd == 1 ? dim0_to : inner;
        int64_t o[N];
        for (int k = 0; k < N; ++k) o[k] = nd == 1 ? i0 * strides[k][0] : off[k];
        for (int64_t i = i0; i < i1; ++i) {
            pos[nd - 1] = i;
            body(pos, o);
            for (int k = 0; k < N; ++k) o[k] += strides[k][nd - 1];
        }
        if (nd == 1) return;
        // advance dims nd-2 .. 0 (dim 0 bounded by dim0_to)
        int d = nd - 2;
        while (d >= 0) {
            int64_t lim = d == 0 ? dim0_to : shape[d];
            int64_t base = d == 0 ? dim0_from : 0;
            if (pos[d] + 1 < lim) {
                pos[d] += 1;
                for (int k = 0; k < N; ++k) off[k] += strides[k][d];
                break;
            }
            for (int k = 0; k < N; ++k) off[k] -= (pos[d] - base) * strides[k][d];
            pos[d] = base;
            --d;
        }
        if (d < 0) return;
    }
}

template <int N, typename Body>
void apply_scalar_path(int nd, const int64_t *shape, const int64_t *const (&strides)[N], bool use_threads,
                       Body &&body) {
    if (use_threads && g_threads_enabled && nd > 1 && shape[0] > 1) {
        const int64_t n0 = shape[0];
#pragma omp parallel for schedule(static)
        for (int64_t r = 0; r < n0; ++r) walk_rows<N>(nd, shape, strides, r, r + 1, body);
    } else {
        walk_rows<N>(nd, shape, strides, 0, nd > 0 ? shape[0] : 1, body);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Scalar primitives (ScalarPrimitives.fs).  f32 math goes through double (rule 1).
// ---------------------------------------------------------------------------------------------------------------
template <typename T> constexpr bool is_fp = std::is_floating_point<T>::value;
template <typename T> constexpr bool is_bool_t = std::is_same<T, bool>::value;
template <typename T> constexpr bool is_sint = std::is_integral<T>::value && std::is_signed<T>::value && !is_bool_t<T>;
template <typename T> constexpr bool is_uint = std::is_integral<T>::value && std::is_unsigned<T>::value && !is_bool_t<T>;
template <typename T> constexpr bool is_int = is_sint<T> || is_uint<T>;

template <typename T> using UnsignedOf = typename std::make_unsigned<T>::type;

template <typename T> T wrap_add(T a, T b) {
    if constexpr (is_int<T>) return (T)((UnsignedOf<T>)a + (UnsignedOf<T>)b);
    else return a + b;
}
template <typename T> T wrap_sub(T a, T b) {
    if constexpr (is_int<T>) return (T)((UnsignedOf<T>)a - (UnsignedOf<T>)b);
    else return a - b;
}
template <typename T> T wrap_mul(T a, T b) {
    if constexpr (std::is_same<T, int16_t>::value || std::is_same<T, uint16_t>::value ||
                  std::is_same<T, int8_t>::value || std::is_same<T, uint8_t>::value)
        return (T)((uint32_t)(UnsignedOf<T>)a * (uint32_t)(UnsignedOf<T>)b);
    else if constexpr (is_int<T>) return (T)((UnsignedOf<T>)a * (UnsignedOf<T>)b);
    else return a * b;
}
template <typename T> T op_div(T a, T b) {
    if constexpr (is_int<T>) {
        if (b == 0) return 0;  // host throws DivideByZeroException: outside the parity domain
        if constexpr (is_sint<T>)
            if (a == std::numeric_limits<T>::min() && b == (T)-1) return a;  // host throws OverflowException
        return (T)(a / b);
    } else return a / b;
}
template <typename T> T op_mod(T a, T b) {
    if constexpr (is_int<T>) {
        if (b == 0) return 0;
        if constexpr (is_sint<T>)
            if (a == std::numeric_limits<T>::min() && b == (T)-1) return 0;
        return (T)(a % b);
    } else if constexpr (std::is_same<T, float>::value) return std::fmod(a, b);
    else return std::fmod(a, b);
}

template <typename T, typename F> T via_double(T x, F f) {
    if constexpr (std::is_same<T, float>::value) return (float)f((double)x);
    else return f(x);
}

template <typename T> T round_half_even(T x) {
    // Math.Round(double) — banker's rounding; nearbyint under the default FE_TONEAREST mode.
    return (T)std::nearbyint((double)x);
}

// ---- unary op bodies -----------------------------------------------------------------------------------------
template <typename T> bool unary_supported(int op) {
    if constexpr (is_bool_t<T>) return op == DN_NEGATE;
    else if constexpr (is_fp<T>) return op != DN_NEGATE;
    else {
        switch (op) {
        case DN_UNARY_PLUS: case DN_UNARY_MINUS: case DN_ABS: return true;
        case DN_SGN: return std::is_same<T, int16_t>::value || std::is_same<T, int32_t>::value ||
                            std::is_same<T, int64_t>::value;
        default: return false;
        }
    }
}

template <typename T> T unary_eval(int op, T x) {
    if constexpr (is_bool_t<T>) {
        return !x;
    } else if constexpr (is_fp<T>) {
        switch (op) {
        case DN_UNARY_PLUS: return x;
        case DN_UNARY_MINUS: return -x;
        case DN_ABS: return std::fabs(x);
        case DN_SGN: return x < 0 ? (T)-1 : (x > 0 ? (T)1 : (T)0);
        case DN_LOG: return via_double(x, [](double v) { return std::log(v); });
        case DN_LOG10: return via_double(x, [](double v) { return std::log10(v); });
        case DN_EXP: return via_double(x, [](double v) { return std::exp(v); });
        case DN_SIN: return via_double(x, [](double v) { return std::sin(v); });
        case DN_COS: return via_double(x, [](double v) { return std::cos(v); });
        case DN_TAN: return via_double(x, [](double v) { return std::tan(v); });
        case DN_ASIN: return via_double(x, [](double v) { return std::asin(v); });
        case DN_ACOS: return via_double(x, [](double v) { return std::acos(v); });
        case DN_ATAN: return via_double(x, [](double v) { return std::atan(v); });
        case DN_SINH: return via_double(x, [](double v) { return std::sinh(v); });
        case DN_COSH: return via_double(x, [](double v) { return std::cosh(v); });
        case DN_TANH: return via_double(x, [](double v) { return std::tanh(v); });
        case DN_SQRT: return std::sqrt(x);
        case DN_CEILING: return via_double(x, [](double v) { return std::ceil(v); });
        case DN_FLOOR: return via_double(x, [](double v) { return std::floor(v); });
        case DN_ROUND: return round_half_even(x);
        case DN_TRUNCATE: return via_double(x, [](double v) { return std::trunc(v); });
        default: return x;
        }
    } else {
        switch (op) {
        case DN_UNARY_PLUS: return x;
        case DN_UNARY_MINUS: return (T)((UnsignedOf<T>)0 - (UnsignedOf<T>)x);
        case DN_ABS:
            if constexpr (is_sint<T>) return x < 0 ? (T)((UnsignedOf<T>)0 - (UnsignedOf<T>)x) : x;
            else return x;
        case DN_SGN: return x < 0 ? (T)-1 : (x > 0 ? (T)1 : (T)0);
        default: return x;
        }
    }
}

// ---- binary op bodies ----------------------------------------------------------------------------------------
template <typename T> bool binary_supported(int op) {
    if constexpr (is_bool_t<T>) return op == DN_AND || op == DN_OR || op == DN_XOR;
    else if constexpr (is_fp<T>) return op <= DN_MIN_ELEMWISE;
    else return op <= DN_MIN_ELEMWISE && op != DN_POWER;
}

template <typename T> T binary_eval(int op, T a, T b) {
    if constexpr (is_bool_t<T>) {
        switch (op) {
        case DN_AND: return a && b;
        case DN_OR: return a || b;
        default: return a != b;
        }
    } else {
        switch (op) {
        case DN_ADD: return wrap_add(a, b);
        case DN_SUBTRACT: return wrap_sub(a, b);
        case DN_MULTIPLY: return wrap_mul(a, b);
        case DN_DIVIDE: return op_div(a, b);
        case DN_MODULO: return op_mod(a, b);
        case DN_POWER:
            if constexpr (std::is_same<T, float>::value) return (float)std::pow((double)a, (double)b);
            else if constexpr (std::is_same<T, double>::value) return std::pow(a, b);
            else return a;
        case DN_MAX_ELEMWISE: return a > b ? a : b;
        case DN_MIN_ELEMWISE: return a < b ? a : b;
        default: return a;
        }
    }
}

template <typename T> bool compare_eval(int op, T a, T b) {
    switch (op) {
    case DN_EQUAL: return a == b;
    case DN_NOT_EQUAL: return a != b;
    case DN_LESS: return a < b;
    case DN_LESS_OR_EQUAL: return a <= b;
    case DN_GREATER: return a > b;
    default: return a >= b;
    }
}

// VectorOps.CanUse (VectorOps.fs:287-301) on the swapped layout: primitive non-bool type, nd > 0, target last
// stride 1, source last strides 0/1.
template <typename T>
bool simd_path(int nd, const int64_t *tstride, const int64_t *const *sstrides, int nsrc) {
    if (nd == 0 || is_bool_t<T>) return false;
    if (tstride[nd - 1] != 1) return false;
    for (int s = 0; s < nsrc; ++s)
        if (sstrides[s][nd - 1] != 0 && sstrides[s][nd - 1] != 1) return false;
    return true;
}

// Which unary/binary ops have a VectorOps implementation (HostBackend.fs:221-229,283-286,312-348).
bool unary_has_simd(int op) { return op == DN_UNARY_MINUS || op == DN_ABS || op == DN_SQRT; }
bool binary_has_simd(int op) {
    return op == DN_ADD || op == DN_SUBTRACT || op == DN_MULTIPLY || op == DN_DIVIDE || op == DN_MAX_ELEMWISE ||
           op == DN_MIN_ELEMWISE;
}

// Generic elementwise driver: applies the HostBackend layout swap, then either the single-threaded SIMD-path
// loop order or the dim-0-parallel scalar path. NSRC sources of possibly different types are passed as raw
// pointers + strides; `body(tptr, srcptrs)` computes one element.
struct Operand {
    char *p;
    int64_t esize;
    int64_t stride[DN_MAX_DIMS];
};

template <int N, typename Body>
void elemwise_drive(int nd, const int64_t *shape_in, Operand (&ops)[N], bool simd_capable_op, bool simd_type,
                    Body &&body) {
    int64_t shape[DN_MAX_DIMS];
    for (int d = 0; d < DN_MAX_DIMS; ++d) shape[d] = d < nd ? shape_in[d] : 1;
    // layout swap
    const int64_t *sstr[N > 1 ? N - 1 : 1];
    for (int k = 1; k < N; ++k) sstr[k - 1] = ops[k].stride;
    int best = elemwise_best_last_dim(nd, shape, ops[0].stride, sstr, N - 1);
    if (best >= 0 && best != nd - 1) {
        std::swap(shape[best], shape[nd - 1]);
        for (int k = 0; k < N; ++k) std::swap(ops[k].stride[best], ops[k].stride[nd - 1]);
    }
    bool simd = simd_capable_op && simd_type && nd > 0 && ops[0].stride[nd - 1] == 1;
    for (int k = 1; k < N && simd; ++k) simd = ops[k].stride[nd - 1] == 0 || ops[k].stride[nd - 1] == 1;
    const int64_t *strides[N];
    for (int k = 0; k < N; ++k) strides[k] = ops[k].stride;
    char *base[N];
    int64_t es[N];
    for (int k = 0; k < N; ++k) {
        base[k] = ops[k].p;
        es[k] = ops[k].esize;
    }
    auto elem = [&](const int64_t *, const int64_t *off) {
        char *ptrs[N];
        for (int k = 0; k < N; ++k) ptrs[k] = base[k] + off[k] * es[k];
        body(ptrs);
    };
    // SIMD path (VectorOps) is single-threaded; scalar path threads over dim 0 when nd > 1.
    apply_scalar_path<N>(nd, shape, strides, /*use_threads=*/!simd, elem);
}

template <typename T> Operand operand_of(const dn_tensor *t) {
    Operand o;
    o.p = (char *)((T *)t->base + t->offset);
    o.esize = sizeof(T);
    for (int d = 0; d < DN_MAX_DIMS; ++d) o.stride[d] = d < t->ndims ? t->stride[d] : 0;
    return o;
}

// ---------------------------------------------------------------------------------------------------------------
// dtype dispatch
// ---------------------------------------------------------------------------------------------------------------
#define DN_DISPATCH(DT, ...)                                                   \
    switch (DT) {                                                              \
    case DN_F32: { using T = float; __VA_ARGS__; } break;                      \
    case DN_F64: { using T = double; __VA_ARGS__; } break;                     \
    case DN_I8: { using T = int8_t; __VA_ARGS__; } break;                      \
    case DN_U8: { using T = uint8_t; __VA_ARGS__; } break;                     \
    case DN_I16: { using T = int16_t; __VA_ARGS__; } break;                    \
    case DN_U16: { using T = uint16_t; __VA_ARGS__; } break;                   \
    case DN_I32: { using T = int32_t; __VA_ARGS__; } break;                    \
    case DN_U32: { using T = uint32_t; __VA_ARGS__; } break;                   \
    case DN_I64: { using T = int64_t; __VA_ARGS__; } break;                    \
    case DN_U64: { using T = uint64_t; __VA_ARGS__; } break;                   \
    case DN_BOOL: { using T = bool; __VA_ARGS__; } break;                      \
    default: return fail(DN_ERR_INVALID_ARG, "bad dtype");                     \
    }

#define DN_DISPATCH2(DT, ...)                                                  \
    switch (DT) {                                                              \
    case DN_F32: { using S = float; __VA_ARGS__; } break;                      \
    case DN_F64: { using S = double; __VA_ARGS__; } break;                     \
    case DN_I8: { using S = int8_t; __VA_ARGS__; } break;                      \
    case DN_U8: { using S = uint8_t; __VA_ARGS__; } break;                     \
    case DN_I16: { using S = int16_t; __VA_ARGS__; } break;                    \
    case DN_U16: { using S = uint16_t; __VA_ARGS__; } break;                   \
    case DN_I32: { using S = int32_t; __VA_ARGS__; } break;                    \
    case DN_U32: { using S = uint32_t; __VA_ARGS__; } break;                   \
    case DN_I64: { using S = int64_t; __VA_ARGS__; } break;                    \
    case DN_U64: { using S = uint64_t; __VA_ARGS__; } break;                   \
    case DN_BOOL: { using S = bool; __VA_ARGS__; } break;                      \
    default: return fail(DN_ERR_INVALID_ARG, "bad dtype");                     \
    }

// Unchecked conversion (rule 7). Out-of-range float->int yields the x64 "integer indefinite" value for 32/64-bit
// targets and its truncation for narrower ones; this is outside the parity domain and only avoids C++ UB here.
template <typename Tt, typename Ts> Tt convert_eval(Ts v) {
    if constexpr (is_bool_t<Tt>) return v != (Ts)0;
    else if constexpr (is_bool_t<Ts>) return (Tt)(v ? 1 : 0);
    else if constexpr (is_fp<Ts> && is_int<Tt>) {
        double d = (double)v;
        if constexpr (std::is_same<Tt, uint64_t>::value) {
            if (!(d > -1.0 && d < 18446744073709551616.0)) return (Tt)0x8000000000000000ull;
            return (Tt)d;
        } else if constexpr (std::is_same<Tt, int64_t>::value) {
            if (!(d > -9223372036854777856.0 && d < 9223372036854775808.0)) return std::numeric_limits<int64_t>::min();
            return (Tt)d;
        } else if constexpr (std::is_same<Tt, uint32_t>::value) {
            if (!(d > -9223372036854777856.0 && d < 9223372036854775808.0)) return 0;
            return (Tt)(int64_t)d;
        } else {
            if (!(d > -2147483649.0 && d < 2147483648.0)) return (Tt)std::numeric_limits<int32_t>::min();
            return (Tt)(int32_t)d;
        }
    } else return (Tt)v;
}

}  // namespace

// =================================================================================================================
// Exported C API (dno_ = "Deep.Net oracle").  Same descriptors and op codes as include/dn_tensor.h, host pointers.
// =================================================================================================================
extern "C" {

const char *dno_last_error(void) { return g_err; }
void dno_set_threads_enabled(int enabled) { g_threads_enabled = enabled; }
// Number of worker threads of the Parallel.For restatement (0 = one per core). Launchers such as torchrun export
// OMP_NUM_THREADS=1; the CPU arm of bench.py overrides that so that it runs on all cores whatever started it.
void dno_set_num_threads(int n) { omp_set_num_threads(n > 0 ? n : omp_get_num_procs()); }
int dno_get_num_threads(void) { return omp_get_max_threads(); }

// HostBackend.FillConst (HostBackend.fs:187-190) -> VectorOps.Fill / ScalarOps.Fill (ScalarOps.fs:363-365).
dn_status dno_fill_const(const dn_tensor *t, const void *value) {
    if (!valid(t) || !value) return fail(DN_ERR_INVALID_ARG, "fill_const: bad argument");
    DN_DISPATCH(t->dtype, {
        T v;
        std::memcpy(&v, value, sizeof(T));
        Operand ops[1] = {operand_of<T>(t)};
        elemwise_drive<1>(t->ndims, t->shape, ops, true, !is_bool_t<T>, [&](char **p) { *(T *)p[0] = v; });
    });
    return DN_OK;
}

// HostBackend.FillIncrementing (HostBackend.fs:192-194) -> ScalarOps.FillIncrementing (ScalarOps.fs:367-370):
// start + incr * conv(pos[0]) evaluated in 'T. NOTE pos[0] is dim 0 of the SWAPPED layout (HostBackend applies
// ElemwiseDataAndLayout first), restated faithfully: for a 1-D target there is nothing to swap.
dn_status dno_fill_incrementing(const dn_tensor *t, const void *start, const void *incr) {
    if (!valid(t) || !start || !incr) return fail(DN_ERR_INVALID_ARG, "fill_incrementing: bad argument");
    if (t->dtype == DN_BOOL) return fail(DN_ERR_UNSUPPORTED, "fill_incrementing: bool");
    DN_DISPATCH(t->dtype, {
        if constexpr (!is_bool_t<T>) {
            T s, i;
            std::memcpy(&s, start, sizeof(T));
            std::memcpy(&i, incr, sizeof(T));
            View<T> v = view_of<T>(t);
            const int64_t *none[1] = {nullptr};
            int best = elemwise_best_last_dim(v.nd, v.shape, v.stride, none, 0);
            if (best >= 0 && best != v.nd - 1) swap_dims(v, best, v.nd - 1);
            const int64_t *strides[1] = {v.stride};
            T *base = v.p;
            apply_scalar_path<1>(v.nd, v.shape, strides, true, [&](const int64_t *pos, const int64_t *off) {
                T p0 = v.nd > 0 ? convert_eval<T, int64_t>(pos[0]) : (T)0;
                base[off[0]] = wrap_add<T>(s, wrap_mul<T>(i, p0));
            });
        }
    });
    return DN_OK;
}

// HostBackend.Copy (HostBackend.fs:196-208).
dn_status dno_copy(const dn_tensor *t, const dn_tensor *a) {
    if (!valid(t) || !valid(a)) return fail(DN_ERR_INVALID_ARG, "copy: bad argument");
    if (t->dtype != a->dtype) return fail(DN_ERR_INVALID_ARG, "copy: dtype mismatch");
    if (!same_shape(t, a)) return fail(DN_ERR_SHAPE_MISMATCH, "copy: shape mismatch");
    DN_DISPATCH(t->dtype, {
        Operand ops[2] = {operand_of<T>(t), operand_of<T>(a)};
        elemwise_drive<2>(t->ndims, t->shape, ops, true, !is_bool_t<T>,
                          [&](char **p) { *(T *)p[0] = *(const T *)p[1]; });
    });
    return DN_OK;
}

// HostBackend.Convert (HostBackend.fs:213-215) -> ScalarOps.Convert (ScalarOps.fs:376-379).
dn_status dno_convert(const dn_tensor *t, const dn_tensor *a) {
    if (!valid(t) || !valid(a)) return fail(DN_ERR_INVALID_ARG, "convert: bad argument");
    if (!same_shape(t, a)) return fail(DN_ERR_SHAPE_MISMATCH, "convert: shape mismatch");
    DN_DISPATCH(t->dtype, {
        DN_DISPATCH2(a->dtype, {
            Operand ops[2] = {operand_of<T>(t), operand_of<S>(a)};
            elemwise_drive<2>(t->ndims, t->shape, ops, false, false,
                              [&](char **p) { *(T *)p[0] = convert_eval<T, S>(*(const S *)p[1]); });
        });
    });
    return DN_OK;
}

// HostBackend unary members (HostBackend.fs:217-310) -> ScalarOps.fs:381-493 / VectorOps.fs:205-215.
dn_status dno_unary(int32_t op, const dn_tensor *t, const dn_tensor *a) {
    if (!valid(t) || !valid(a) || op < 0 || op >= DN_UNARY_OP_COUNT)
        return fail(DN_ERR_INVALID_ARG, "unary: bad argument");
    if (t->dtype != a->dtype) return fail(DN_ERR_INVALID_ARG, "unary: dtype mismatch");
    if (!same_shape(t, a)) return fail(DN_ERR_SHAPE_MISMATCH, "unary: shape mismatch");
    DN_DISPATCH(t->dtype, {
        if (!unary_supported<T>(op)) return fail(DN_ERR_UNSUPPORTED, "unary: op not defined for dtype");
        Operand ops[2] = {operand_of<T>(t), operand_of<T>(a)};
        elemwise_drive<2>(t->ndims, t->shape, ops, unary_has_simd(op), !is_bool_t<T>,
                          [&](char **p) { *(T *)p[0] = unary_eval<T>(op, *(const T *)p[1]); });
    });
    return DN_OK;
}

// HostBackend binary members (HostBackend.fs:312-384) -> ScalarOps.fs:495-575 / VectorOps.fs:217-240.
dn_status dno_binary(int32_t op, const dn_tensor *t, const dn_tensor *a, const dn_tensor *b) {
    if (!valid(t) || !valid(a) || !valid(b) || op < 0 || op >= DN_BINARY_OP_COUNT)
        return fail(DN_ERR_INVALID_ARG, "binary: bad argument");
    if (t->dtype != a->dtype || t->dtype != b->dtype) return fail(DN_ERR_INVALID_ARG, "binary: dtype mismatch");
    if (!same_shape(t, a) || !same_shape(t, b)) return fail(DN_ERR_SHAPE_MISMATCH, "binary: shape mismatch");
    DN_DISPATCH(t->dtype, {
        if (!binary_supported<T>(op)) return fail(DN_ERR_UNSUPPORTED, "binary: op not defined for dtype");
        Operand ops[3] = {operand_of<T>(t), operand_of<T>(a), operand_of<T>(b)};
        elemwise_drive<3>(t->ndims, t->shape, ops, binary_has_simd(op), !is_bool_t<T>, [&](char **p) {
            *(T *)p[0] = binary_eval<T>(op, *(const T *)p[1], *(const T *)p[2]);
        });
    });
    return DN_OK;
}

// HostBackend comparison members (HostBackend.fs:350-372) -> ScalarOps.fs:535-563.
dn_status dno_compare(int32_t op, const dn_tensor *t, const dn_tensor *a, const dn_tensor *b) {
    if (!valid(t) || !valid(a) || !valid(b) || op < 0 || op >= DN_COMPARE_OP_COUNT)
        return fail(DN_ERR_INVALID_ARG, "compare: bad argument");
    if (t->dtype != DN_BOOL || a->dtype != b->dtype) return fail(DN_ERR_INVALID_ARG, "compare: dtype mismatch");
    if (!same_shape(t, a) || !same_shape(t, b)) return fail(DN_ERR_SHAPE_MISMATCH, "compare: shape mismatch");
    DN_DISPATCH(a->dtype, {
        Operand ops[3] = {operand_of<bool>(t), operand_of<T>(a), operand_of<T>(b)};
        elemwise_drive<3>(t->ndims, t->shape, ops, false, false, [&](char **p) {
            *(bool *)p[0] = compare_eval<T>(op, *(const T *)p[1], *(const T *)p[2]);
        });
    });
    return DN_OK;
}

// HostBackend.IsFinite (HostBackend.fs:304-306) -> ScalarPrimitives.fs:177-184.
dn_status dno_is_finite(const dn_tensor *t, const dn_tensor *a) {
    if (!valid(t) || !valid(a)) return fail(DN_ERR_INVALID_ARG, "is_finite: bad argument");
    if (t->dtype != DN_BOOL) return fail(DN_ERR_INVALID_ARG, "is_finite: target must be bool");
    if (!same_shape(t, a)) return fail(DN_ERR_SHAPE_MISMATCH, "is_finite: shape mismatch");
    DN_DISPATCH(a->dtype, {
        Operand ops[2] = {operand_of<bool>(t), operand_of<T>(a)};
        elemwise_drive<2>(t->ndims, t->shape, ops, false, false, [&](char **p) {
            if constexpr (is_fp<T>) *(bool *)p[0] = std::isfinite(*(const T *)p[1]);
            else *(bool *)p[0] = true;
        });
    });
    return DN_OK;
}

// HostBackend.IfThenElse (HostBackend.fs:386-389) -> ScalarOps.fs:577-581.
dn_status dno_if_then_else(const dn_tensor *t, const dn_tensor *c, const dn_tensor *a, const dn_tensor *b) {
    if (!valid(t) || !valid(c) || !valid(a) || !valid(b)) return fail(DN_ERR_INVALID_ARG, "if_then_else: bad argument");
    if (c->dtype != DN_BOOL || t->dtype != a->dtype || t->dtype != b->dtype)
        return fail(DN_ERR_INVALID_ARG, "if_then_else: dtype mismatch");
    if (!same_shape(t, c) || !same_shape(t, a) || !same_shape(t, b))
        return fail(DN_ERR_SHAPE_MISMATCH, "if_then_else: shape mismatch");
    DN_DISPATCH(t->dtype, {
        Operand ops[4] = {operand_of<T>(t), operand_of<bool>(c), operand_of<T>(a), operand_of<T>(b)};
        elemwise_drive<4>(t->ndims, t->shape, ops, false, false, [&](char **p) {
            *(T *)p[0] = *(const bool *)p[1] ? *(const T *)p[2] : *(const T *)p[3];
        });
    });
    return DN_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// Axis folds.  ScalarOps.ApplyAxisFold (ScalarOps.fs:299-361): for every target position (row-major, dim 0 in
// parallel when the SOURCE rank > 1) fold the source's last axis sequentially from the initial state.
// No layout swap here: HostBackend uses GetDataAndLayout for reductions (HostBackend.fs:423-461).
// ---------------------------------------------------------------------------------------------------------------
namespace {

template <typename Tt, typename Ts, typename Fold>
void axis_fold(const dn_tensor *t, const dn_tensor *a, Fold &&fold) {
    View<Tt> tv = view_of<Tt>(t);
    View<Ts> av = view_of<Ts>(a);
    const int nd = av.nd;  // source rank; target rank nd-1
    const int64_t L = av.shape[nd - 1];
    const int64_t sL = av.stride[nd - 1];
    const int tnd = nd - 1;
    int64_t shape[DN_MAX_DIMS];
    for (int d = 0; d < DN_MAX_DIMS; ++d) shape[d] = d < tnd ? av.shape[d] : 1;
    for (int d = 0; d < tnd; ++d)
        if (shape[d] == 0) return;
    auto row = [&](const int64_t *, const int64_t *off) {
        tv.p[off[0]] = fold(av.p + off[1], L, sL);
    };
    const int64_t *strides[2] = {tv.stride, av.stride};
    if (tnd == 0) {
        int64_t pos[1] = {0}, off[2] = {0, 0};
        row(pos, off);
        return;
    }
    // ApplyAxisFold threads when src nd > 1, i.e. always when the target has at least one dim.
    if (g_threads_enabled && shape[0] > 1) {
        const int64_t n0 = shape[0];
#pragma omp parallel for schedule(static)
        for (int64_t r = 0; r < n0; ++r) walk_rows<2>(tnd, shape, strides, r, r + 1, row);
    } else {
        walk_rows<2>(tnd, shape, strides, 0, shape[0], row);
    }
}

dn_status check_reduce_shapes(const dn_tensor *t, const dn_tensor *a, const char *what) {
    if (!valid(t) || !valid(a)) return fail(DN_ERR_INVALID_ARG, what);
    if (a->ndims < 1 || t->ndims != a->ndims - 1) return fail(DN_ERR_SHAPE_MISMATCH, what);
    for (int d = 0; d < t->ndims; ++d)
        if (t->shape[d] != a->shape[d]) return fail(DN_ERR_SHAPE_MISMATCH, what);
    return DN_OK;
}

}  // namespace

extern "C" {

// HostBackend.*LastAxis (HostBackend.fs:423-449) -> ScalarOps.fs:606-636.
dn_status dno_reduce_last_axis(int32_t op, const dn_tensor *t, const dn_tensor *a) {
    dn_status st = check_reduce_shapes(t, a, "reduce_last_axis: bad shapes");
    if (st != DN_OK) return st;
    if (op < 0 || op >= DN_REDUCE_OP_COUNT) return fail(DN_ERR_INVALID_ARG, "reduce_last_axis: bad op");
    if (op == DN_COUNT_TRUE) {
        if (a->dtype != DN_BOOL || t->dtype != DN_I64) return fail(DN_ERR_INVALID_ARG, "count_true: dtypes");
        axis_fold<int64_t, bool>(t, a, [](const bool *p, int64_t L, int64_t s) {
            int64_t r = 0;
            for (int64_t i = 0; i < L; ++i, p += s) r = *p ? r + 1 : r;
            return r;
        });
        return DN_OK;
    }
    if (op == DN_ALL || op == DN_ANY) {
        if (a->dtype != DN_BOOL || t->dtype != DN_BOOL) return fail(DN_ERR_INVALID_ARG, "all/any: dtypes");
        const bool is_all = op == DN_ALL;
        axis_fold<bool, bool>(t, a, [is_all](const bool *p, int64_t L, int64_t s) {
            bool r = is_all;
            for (int64_t i = 0; i < L; ++i, p += s) r = is_all ? (r && *p) : (r || *p);
            return r;
        });
        return DN_OK;
    }
    if (t->dtype != a->dtype) return fail(DN_ERR_INVALID_ARG, "reduce_last_axis: dtype mismatch");
    if (a->dtype == DN_BOOL) return fail(DN_ERR_UNSUPPORTED, "reduce_last_axis: numeric fold on bool");
    DN_DISPATCH(a->dtype, {
        if constexpr (!is_bool_t<T>) {
            switch (op) {
            case DN_SUM:
                axis_fold<T, T>(t, a, [](const T *p, int64_t L, int64_t s) {
                    T r = (T)0;
                    for (int64_t i = 0; i < L; ++i, p += s) r = wrap_add<T>(r, *p);
                    return r;
                });
                break;
            case DN_PRODUCT:
                axis_fold<T, T>(t, a, [](const T *p, int64_t L, int64_t s) {
                    T r = (T)1;
                    for (int64_t i = 0; i < L; ++i, p += s) r = wrap_mul<T>(r, *p);
                    return r;
                });
                break;
            case DN_MAX:
                axis_fold<T, T>(t, a, [](const T *p, int64_t L, int64_t s) {
                    T r = std::numeric_limits<T>::lowest();  // minValue<'T>: FINITE for floats (Utils.fs:264-281)
                    for (int64_t i = 0; i < L; ++i, p += s) r = (r > *p) ? r : *p;
                    return r;
                });
                break;
            default:  // DN_MIN
                axis_fold<T, T>(t, a, [](const T *p, int64_t L, int64_t s) {
                    T r = std::numeric_limits<T>::max();
                    for (int64_t i = 0; i < L; ++i, p += s) r = (r < *p) ? r : *p;
                    return r;
                });
                break;
            }
        }
    });
    return DN_OK;
}

// HostBackend.ArgMin/ArgMaxLastAxis (HostBackend.fs:451-457) -> ScalarOps.fs:638-654.
dn_status dno_arg_reduce_last_axis(int32_t op, const dn_tensor *t, const dn_tensor *a) {
    dn_status st = check_reduce_shapes(t, a, "arg_reduce_last_axis: bad shapes");
    if (st != DN_OK) return st;
    if (t->dtype != DN_I64) return fail(DN_ERR_INVALID_ARG, "arg_reduce_last_axis: target must be int64");
    if (op != DN_ARG_MIN && op != DN_ARG_MAX) return fail(DN_ERR_INVALID_ARG, "arg_reduce_last_axis: bad op");
    if (a->dtype == DN_BOOL) return fail(DN_ERR_UNSUPPORTED, "arg_reduce_last_axis: bool");
    DN_DISPATCH(a->dtype, {
        if constexpr (!is_bool_t<T>) {
            if (op == DN_ARG_MAX)
                axis_fold<int64_t, T>(t, a, [](const T *p, int64_t L, int64_t s) {
                    int64_t pos = kNotFound;
                    T best = std::numeric_limits<T>::lowest();
                    for (int64_t i = 0; i < L; ++i, p += s)
                        if (*p > best) { pos = i; best = *p; }
                    return pos;
                });
            else
                axis_fold<int64_t, T>(t, a, [](const T *p, int64_t L, int64_t s) {
                    int64_t pos = kNotFound;
                    T best = std::numeric_limits<T>::max();
                    for (int64_t i = 0; i < L; ++i, p += s)
                        if (*p < best) { pos = i; best = *p; }
                    return pos;
                });
        }
    });
    return DN_OK;
}

// HostBackend.FindLastAxis (HostBackend.fs:459-461) -> ScalarOps.fs:656-665.
dn_status dno_find_last_axis(const void *value, const dn_tensor *t, const dn_tensor *a) {
    dn_status st = check_reduce_shapes(t, a, "find_last_axis: bad shapes");
    if (st != DN_OK) return st;
    if (t->dtype != DN_I64 || !value) return fail(DN_ERR_INVALID_ARG, "find_last_axis: bad argument");
    DN_DISPATCH(a->dtype, {
        T v;
        std::memcpy(&v, value, sizeof(T));
        axis_fold<int64_t, T>(t, a, [v](const T *p, int64_t L, int64_t s) {
            for (int64_t i = 0; i < L; ++i, p += s)
                if (*p == v) return i;
            return kNotFound;
        });
    });
    return DN_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// Indexing.
// ---------------------------------------------------------------------------------------------------------------
namespace {

// FastLayout32.Addr (FastAccess.fs:41-51): bounds-checked address; returns false when out of range.
template <typename V>
bool checked_addr(const V &v, const int64_t *pos, int64_t &addr) {
    addr = 0;
    for (int d = 0; d < v.nd; ++d) {
        if (pos[d] < 0 || pos[d] >= v.shape[d]) return false;
        addr += pos[d] * v.stride[d];
    }
    return true;
}

}  // namespace

extern "C" {

// HostBackend.Gather (HostBackend.fs:391-397) -> ScalarOps.Gather (ScalarOps.fs:583-591). Threads over dim 0 of
// the target when its rank > 1.
dn_status dno_gather(const dn_tensor *t, const dn_tensor *const *idxs, int32_t nidxs, const dn_tensor *a) {
    if (!valid(t) || !valid(a) || !idxs) return fail(DN_ERR_INVALID_ARG, "gather: bad argument");
    if (t->dtype != a->dtype) return fail(DN_ERR_INVALID_ARG, "gather: dtype mismatch");
    if (nidxs != a->ndims) return fail(DN_ERR_INVALID_ARG, "gather: one index tensor per source dim required");
    for (int d = 0; d < nidxs; ++d) {
        if (idxs[d]) {
            if (!valid(idxs[d]) || idxs[d]->dtype != DN_I64 || !same_shape(idxs[d], t))
                return fail(DN_ERR_INVALID_ARG, "gather: index tensors must be int64 with the target's shape");
        } else if (d >= t->ndims)
            return fail(DN_ERR_INVALID_ARG, "gather: None index beyond target rank");
    }
    bool oob = false;
    DN_DISPATCH(t->dtype, {
        View<T> tv = view_of<T>(t), av = view_of<T>(a);
        View<int64_t> iv[DN_MAX_DIMS];
        for (int d = 0; d < nidxs; ++d)
            if (idxs[d]) iv[d] = view_of<int64_t>(idxs[d]);
        const int64_t *strides[1] = {tv.stride};
        apply_scalar_path<1>(tv.nd, tv.shape, strides, true, [&](const int64_t *pos, const int64_t *off) {
            int64_t spos[DN_MAX_DIMS];
            for (int d = 0; d < av.nd; ++d) {
                if (idxs[d]) {
                    int64_t ia = 0;
                    for (int k = 0; k < tv.nd; ++k) ia += pos[k] * iv[d].stride[k];
                    spos[d] = iv[d].p[ia];
                } else spos[d] = pos[d];
            }
            int64_t addr;
            if (!checked_addr(av, spos, addr)) { oob = true; return; }
            tv.p[off[0]] = av.p[addr];
        });
    });
    if (oob) return fail(DN_ERR_INDEX_OUT_OF_RANGE, "invalid index during gather or scatter");
    return DN_OK;
}

// HostBackend.Scatter (HostBackend.fs:399-405) -> ScalarOps.Scatter (ScalarOps.fs:593-604). Single-threaded, adds
// in logical row-major source order. The zero-fill the frontend does first (Tensor.fs:2159) is included here, as
// it is in the CUDA backend (CudaBackend.fs:379), so that the call is self-contained like dn_scatter.
dn_status dno_scatter(const dn_tensor *t, const dn_tensor *const *idxs, int32_t nidxs, const dn_tensor *a) {
    if (!valid(t) || !valid(a) || !idxs) return fail(DN_ERR_INVALID_ARG, "scatter: bad argument");
    if (t->dtype != a->dtype) return fail(DN_ERR_INVALID_ARG, "scatter: dtype mismatch");
    if (t->dtype == DN_BOOL) return fail(DN_ERR_UNSUPPORTED, "scatter: bool has no addition");
    if (nidxs != t->ndims) return fail(DN_ERR_INVALID_ARG, "scatter: one index tensor per target dim required");
    for (int d = 0; d < nidxs; ++d) {
        if (idxs[d]) {
            if (!valid(idxs[d]) || idxs[d]->dtype != DN_I64 || !same_shape(idxs[d], a))
                return fail(DN_ERR_INVALID_ARG, "scatter: index tensors must be int64 with the source's shape");
        } else if (d >= a->ndims)
            return fail(DN_ERR_INVALID_ARG, "scatter: None index beyond source rank");
    }
    bool oob = false;
    DN_DISPATCH(t->dtype, {
        if constexpr (!is_bool_t<T>) {
            View<T> tv = view_of<T>(t), av = view_of<T>(a);
            {
                const int64_t *ts[1] = {tv.stride};
                walk_rows<1>(tv.nd, tv.shape, ts, 0, tv.nd > 0 ? tv.shape[0] : 1,
                             [&](const int64_t *, const int64_t *off) { tv.p[off[0]] = (T)0; });
            }
            View<int64_t> iv[DN_MAX_DIMS];
            for (int d = 0; d < nidxs; ++d)
                if (idxs[d]) iv[d] = view_of<int64_t>(idxs[d]);
            const int64_t *strides[1] = {av.stride};
            walk_rows<1>(av.nd, av.shape, strides, 0, av.nd > 0 ? av.shape[0] : 1,
                         [&](const int64_t *pos, const int64_t *off) {
                             if (oob) return;
                             int64_t tpos[DN_MAX_DIMS];
                             for (int d = 0; d < tv.nd; ++d) {
                                 if (idxs[d]) {
                                     int64_t ia = 0;
                                     for (int k = 0; k < av.nd; ++k) ia += pos[k] * iv[d].stride[k];
                                     tpos[d] = iv[d].p[ia];
                                 } else tpos[d] = pos[d];
                             }
                             int64_t addr;
                             if (!checked_addr(tv, tpos, addr)) { oob = true; return; }
                             tv.p[addr] = wrap_add<T>(tv.p[addr], av.p[off[0]]);
                         });
        }
    });
    if (oob) return fail(DN_ERR_INDEX_OUT_OF_RANGE, "invalid index during gather or scatter");
    return DN_OK;
}

// Tensor.countTrue (Tensor.fs:2232-2233): flatten -> countTrueAxis 0 -> value.
dn_status dno_count_true(const dn_tensor *a, int64_t *count) {
    if (!valid(a) || !count || a->dtype != DN_BOOL) return fail(DN_ERR_INVALID_ARG, "count_true: bad argument");
    View<bool> av = view_of<bool>(a);
    int64_t n = 0;
    const int64_t *strides[1] = {av.stride};
    walk_rows<1>(av.nd, av.shape, strides, 0, av.nd > 0 ? av.shape[0] : 1,
                 [&](const int64_t *, const int64_t *off) { n += av.p[off[0]] ? 1 : 0; });
    *count = n;
    return DN_OK;
}

static dn_status check_masks(const dn_tensor *full, const dn_tensor *const *masks, int32_t nmasks, const char *what) {
    if (!masks || nmasks != full->ndims) return fail(DN_ERR_INVALID_ARG, what);
    for (int d = 0; d < nmasks; ++d)
        if (masks[d]) {
            if (!valid(masks[d]) || masks[d]->dtype != DN_BOOL || masks[d]->ndims != 1 ||
                masks[d]->shape[0] != full->shape[d])
                return fail(DN_ERR_INVALID_ARG, what);
        }
    return DN_OK;
}

// HostBackend.MaskedGet (HostBackend.fs:407-411) -> ScalarOps.MaskedGet (ScalarOps.fs:667-681): walk the source
// in row-major order, copy elements whose per-dim masks are all true into the next target position; stops when
// the target is full.
dn_status dno_masked_get(const dn_tensor *t, const dn_tensor *a, const dn_tensor *const *masks, int32_t nmasks) {
    if (!valid(t) || !valid(a)) return fail(DN_ERR_INVALID_ARG, "masked_get: bad argument");
    if (t->dtype != a->dtype || t->ndims != a->ndims) return fail(DN_ERR_INVALID_ARG, "masked_get: dtype/rank mismatch");
    dn_status st = check_masks(a, masks, nmasks, "masked_get: masks must be 1-D bool of the source's dim sizes");
    if (st != DN_OK) return st;
    bool overflow = false;
    DN_DISPATCH(t->dtype, {
        View<T> tv = view_of<T>(t), av = view_of<T>(a);
        std::vector<T *> slots;  // target addresses in row-major order
        const int64_t *ts[1] = {tv.stride};
        walk_rows<1>(tv.nd, tv.shape, ts, 0, tv.nd > 0 ? tv.shape[0] : 1,
                     [&](const int64_t *, const int64_t *off) { slots.push_back(tv.p + off[0]); });
        size_t next = 0;
        const int64_t *as[1] = {av.stride};
        walk_rows<1>(av.nd, av.shape, as, 0, av.nd > 0 ? av.shape[0] : 1,
                     [&](const int64_t *pos, const int64_t *off) {
                         bool m = true;
                         for (int d = 0; d < av.nd; ++d)
                             if (masks[d]) {
                                 const bool *mp = (const bool *)masks[d]->base + masks[d]->offset;
                                 m = m && mp[pos[d] * masks[d]->stride[0]];
                             }
                         if (!m) return;
                         if (next < slots.size()) *slots[next] = av.p[off[0]];
                         else overflow = true;
                         ++next;
                     });
        if (next < slots.size()) overflow = true;
    });
    if (overflow) return fail(DN_ERR_SHAPE_MISMATCH, "masked_get: target size does not match the number of selected elements");
    return DN_OK;
}

// HostBackend.MaskedSet (HostBackend.fs:413-417) -> ScalarOps.MaskedSet (ScalarOps.fs:683-697).
dn_status dno_masked_set(const dn_tensor *t, const dn_tensor *const *masks, int32_t nmasks, const dn_tensor *a) {
    if (!valid(t) || !valid(a)) return fail(DN_ERR_INVALID_ARG, "masked_set: bad argument");
    if (t->dtype != a->dtype || t->ndims != a->ndims) return fail(DN_ERR_INVALID_ARG, "masked_set: dtype/rank mismatch");
    dn_status st = check_masks(t, masks, nmasks, "masked_set: masks must be 1-D bool of the target's dim sizes");
    if (st != DN_OK) return st;
    bool overflow = false;
    DN_DISPATCH(t->dtype, {
        View<T> tv = view_of<T>(t), av = view_of<T>(a);
        std::vector<const T *> vals;
        const int64_t *as[1] = {av.stride};
        walk_rows<1>(av.nd, av.shape, as, 0, av.nd > 0 ? av.shape[0] : 1,
                     [&](const int64_t *, const int64_t *off) { vals.push_back(av.p + off[0]); });
        size_t next = 0;
        const int64_t *ts[1] = {tv.stride};
        walk_rows<1>(tv.nd, tv.shape, ts, 0, tv.nd > 0 ? tv.shape[0] : 1,
                     [&](const int64_t *pos, const int64_t *off) {
                         bool m = true;
                         for (int d = 0; d < tv.nd; ++d)
                             if (masks[d]) {
                                 const bool *mp = (const bool *)masks[d]->base + masks[d]->offset;
                                 m = m && mp[pos[d] * masks[d]->stride[0]];
                             }
                         if (!m) return;
                         if (next < vals.size()) tv.p[off[0]] = *vals[next];
                         else overflow = true;
                         ++next;
                     });
        if (next < vals.size()) overflow = true;
    });
    if (overflow) return fail(DN_ERR_SHAPE_MISMATCH, "masked_set: value size does not match the number of selected elements");
    return DN_OK;
}

// HostBackend.TrueIndices (HostBackend.fs:419-421) -> ScalarOps.TrueIndices (ScalarOps.fs:699-707).
dn_status dno_true_indices(const dn_tensor *t, const dn_tensor *a) {
    if (!valid(t) || !valid(a)) return fail(DN_ERR_INVALID_ARG, "true_indices: bad argument");
    if (t->dtype != DN_I64 || a->dtype != DN_BOOL || t->ndims != 2 || t->shape[1] != a->ndims)
        return fail(DN_ERR_INVALID_ARG, "true_indices: target must be int64 [nTrue, ndims]");
    View<int64_t> tv = view_of<int64_t>(t);
    View<bool> av = view_of<bool>(a);
    int64_t row = 0;
    bool overflow = false;
    const int64_t *as[1] = {av.stride};
    walk_rows<1>(av.nd, av.shape, as, 0, av.nd > 0 ? av.shape[0] : 1, [&](const int64_t *pos, const int64_t *off) {
        if (!av.p[off[0]]) return;
        if (row < tv.shape[0])
            for (int d = 0; d < av.nd; ++d) tv.p[row * tv.stride[0] + d * tv.stride[1]] = pos[d];
        else overflow = true;
        ++row;
    });
    if (overflow || row != tv.shape[0])
        return fail(DN_ERR_SHAPE_MISMATCH, "true_indices: target rows do not match the number of true elements");
    return DN_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Dense contractions.  Host uses MKL (blob missing): restated as an fp64-accumulated reference rounded to 'T,
// which is what a correct BLAS agrees with to ~1e-6 rel; tolerance for the tf32 GPU GEMM is rel 1e-2 (north_star).
// HostBackend.fs:463-546.
// ---------------------------------------------------------------------------------------------------------------
static dn_status matmul_impl(const dn_tensor *t, const dn_tensor *a, const dn_tensor *b, int nbatch_dims) {
    if (!valid(t) || !valid(a) || !valid(b)) return fail(DN_ERR_INVALID_ARG, "mat_mat_dot: bad argument");
    if (t->dtype != a->dtype || t->dtype != b->dtype) return fail(DN_ERR_INVALID_ARG, "mat_mat_dot: dtype mismatch");
    if (t->dtype != DN_F32 && t->dtype != DN_F64) return fail(DN_ERR_UNSUPPORTED, "mat_mat_dot: f32/f64 only");
    const int nd = nbatch_dims + 2;
    if (t->ndims != nd || a->ndims != nd || b->ndims != nd) return fail(DN_ERR_SHAPE_MISMATCH, "mat_mat_dot: rank");
    const int64_t M = a->shape[nd - 2], K = a->shape[nd - 1], N = b->shape[nd - 1];
    if (b->shape[nd - 2] != K || t->shape[nd - 2] != M || t->shape[nd - 1] != N)
        return fail(DN_ERR_SHAPE_MISMATCH, "mat_mat_dot: inner dimensions");
    int64_t nbatch = 1;
    for (int d = 0; d < nbatch_dims; ++d) {
        if (a->shape[d] != t->shape[d] || b->shape[d] != t->shape[d])
            return fail(DN_ERR_SHAPE_MISMATCH, "mat_mat_dot: batch dimensions");
        nbatch *= t->shape[d];
    }
    if (nbatch == 0 || M == 0 || N == 0) return DN_OK;
    auto run = [&](auto zero) {
        using T = decltype(zero);
        const T *ap = (const T *)a->base + a->offset;
        const T *bp = (const T *)b->base + b->offset;
        T *tp = (T *)t->base + t->offset;
#pragma omp parallel for schedule(static) collapse(2)
        for (int64_t bi = 0; bi < nbatch; ++bi)
            for (int64_t m = 0; m < M; ++m) {
                int64_t rem = bi, ao = 0, bo = 0, to = 0;
                for (int d = nbatch_dims - 1; d >= 0; --d) {
                    int64_t p = rem % t->shape[d];
                    rem /= t->shape[d];
                    ao += p * a->stride[d];
                    bo += p * b->stride[d];
                    to += p * t->stride[d];
                }
                std::vector<double> acc((size_t)N, 0.0);
                for (int64_t k = 0; k < K; ++k) {
                    const double av = (double)ap[ao + m * a->stride[nd - 2] + k * a->stride[nd - 1]];
                    const T *brow = bp + bo + k * b->stride[nd - 2];
                    const int64_t bs = b->stride[nd - 1];
                    for (int64_t n = 0; n < N; ++n) acc[(size_t)n] += av * (double)brow[n * bs];
                }
                for (int64_t n = 0; n < N; ++n)
                    tp[to + m * t->stride[nd - 2] + n * t->stride[nd - 1]] = (T)acc[(size_t)n];
            }
    };
    if (t->dtype == DN_F32) run(0.0f);
    else run(0.0);
    return DN_OK;
}

dn_status dno_mat_mat_dot(const dn_tensor *t, const dn_tensor *a, const dn_tensor *b) {
    return matmul_impl(t, a, b, 0);
}
dn_status dno_batched_mat_mat_dot(const dn_tensor *t, const dn_tensor *a, const dn_tensor *b) {
    if (!valid(t) || t->ndims < 2) return fail(DN_ERR_INVALID_ARG, "batched_mat_mat_dot: bad argument");
    return matmul_impl(t, a, b, t->ndims - 2);
}

// VecVecDot / MatVecDot (HostBackend.fs:463-493): fp64-accumulated reference.
dn_status dno_vec_vec_dot(const dn_tensor *t, const dn_tensor *a, const dn_tensor *b) {
    if (!valid(t) || !valid(a) || !valid(b)) return fail(DN_ERR_INVALID_ARG, "vec_vec_dot: bad argument");
    if (t->ndims != 0 || a->ndims != 1 || b->ndims != 1 || a->shape[0] != b->shape[0])
        return fail(DN_ERR_SHAPE_MISMATCH, "vec_vec_dot: shapes");
    if (t->dtype != a->dtype || t->dtype != b->dtype || (t->dtype != DN_F32 && t->dtype != DN_F64))
        return fail(DN_ERR_UNSUPPORTED, "vec_vec_dot: f32/f64 only");
    double acc = 0;
    for (int64_t i = 0; i < a->shape[0]; ++i) {
        if (t->dtype == DN_F32)
            acc += (double)((const float *)a->base)[a->offset + i * a->stride[0]] *
                   (double)((const float *)b->base)[b->offset + i * b->stride[0]];
        else
            acc += ((const double *)a->base)[a->offset + i * a->stride[0]] *
                   ((const double *)b->base)[b->offset + i * b->stride[0]];
    }
    if (t->dtype == DN_F32) ((float *)t->base)[t->offset] = (float)acc;
    else ((double *)t->base)[t->offset] = acc;
    return DN_OK;
}

dn_status dno_mat_vec_dot(const dn_tensor *t, const dn_tensor *a, const dn_tensor *b) {
    if (!valid(t) || !valid(a) || !valid(b)) return fail(DN_ERR_INVALID_ARG, "mat_vec_dot: bad argument");
    if (t->ndims != 1 || a->ndims != 2 || b->ndims != 1 || a->shape[1] != b->shape[0] || t->shape[0] != a->shape[0])
        return fail(DN_ERR_SHAPE_MISMATCH, "mat_vec_dot: shapes");
    if (t->dtype != a->dtype || t->dtype != b->dtype || (t->dtype != DN_F32 && t->dtype != DN_F64))
        return fail(DN_ERR_UNSUPPORTED, "mat_vec_dot: f32/f64 only");
    for (int64_t m = 0; m < a->shape[0]; ++m) {
        double acc = 0;
        for (int64_t k = 0; k < a->shape[1]; ++k) {
            if (t->dtype == DN_F32)
                acc += (double)((const float *)a->base)[a->offset + m * a->stride[0] + k * a->stride[1]] *
                       (double)((const float *)b->base)[b->offset + k * b->stride[0]];
            else
                acc += ((const double *)a->base)[a->offset + m * a->stride[0] + k * a->stride[1]] *
                       ((const double *)b->base)[b->offset + k * b->stride[0]];
        }
        if (t->dtype == DN_F32) ((float *)t->base)[t->offset + m * t->stride[0]] = (float)acc;
        else ((double *)t->base)[t->offset + m * t->stride[0]] = acc;
    }
    return DN_OK;
}

}  // extern "C"

// BatchedInvert (HostBackend.fs:548-577): copy the source into the target, then per matrix LAPACKE_?getrf (LU with
// partial pivoting: largest magnitude in the column, first occurrence — LAPACK dgetf2 / idamax) followed by
// LAPACKE_?getri. MKL itself is not in the reference tree (SURVEY.md §8c "third-party arithmetic"); its published
// algorithm is restated: unblocked right-looking LU, then A^-1 from  U x = L^-1 P e_j  column by column. info > 0
// (an exactly zero pivot) = SingularMatrixException. Computed in the element type, like LAPACK.
template <class T>
static dn_status invert_impl(const dn_tensor *t, const dn_tensor *a) {
    const int nd = t->ndims;
    const int64_t n = t->shape[nd - 1];
    int64_t batch = 1;
    for (int d = 0; d < nd - 2; ++d) batch *= t->shape[d];
    const T *ap = (const T *)a->base + a->offset;
    T *tp = (T *)t->base + t->offset;
    std::vector<T> lu((size_t)(n * n)), inv((size_t)(n * n)), col((size_t)n);
    std::vector<int64_t> piv((size_t)n);
    for (int64_t bidx = 0; bidx < batch; ++bidx) {
        int64_t rem = bidx, ao = 0, to = 0;
        for (int d = nd - 3; d >= 0; --d) {
            const int64_t p = rem % t->shape[d];
            rem /= t->shape[d];
            ao += p * a->stride[d];
            to += p * t->stride[d];
        }
        for (int64_t i = 0; i < n; ++i)
            for (int64_t j = 0; j < n; ++j) lu[(size_t)(i * n + j)] = ap[ao + i * a->stride[nd - 2] + j * a->stride[nd - 1]];
        // getrf
        for (int64_t k = 0; k < n; ++k) {
            int64_t p = k;
            T best = std::fabs(lu[(size_t)(k * n + k)]);
            for (int64_t i = k + 1; i < n; ++i) {
                const T v = std::fabs(lu[(size_t)(i * n + k)]);
                if (v > best) { best = v; p = i; }
            }
            piv[(size_t)k] = p;
            if (!(best > T(0))) return fail(DN_ERR_SINGULAR_MATRIX, "cannot invert singular matrix");
            if (p != k)
                for (int64_t j = 0; j < n; ++j) std::swap(lu[(size_t)(k * n + j)], lu[(size_t)(p * n + j)]);
            const T pivot = lu[(size_t)(k * n + k)];
            for (int64_t i = k + 1; i < n; ++i) {
                const T f = lu[(size_t)(i * n + k)] / pivot;
                lu[(size_t)(i * n + k)] = f;
                for (int64_t j = k + 1; j < n; ++j) lu[(size_t)(i * n + j)] -= f * lu[(size_t)(k * n + j)];
            }
        }
        // getri: column j of the inverse solves L U x = P e_j
        for (int64_t j = 0; j < n; ++j) {
            for (int64_t i = 0; i < n; ++i) col[(size_t)i] = T(i == j ? 1 : 0);
            for (int64_t k = 0; k < n; ++k) std::swap(col[(size_t)k], col[(size_t)piv[(size_t)k]]);
            for (int64_t i = 0; i < n; ++i) {
                T acc = col[(size_t)i];
                for (int64_t k = 0; k < i; ++k) acc -= lu[(size_t)(i * n + k)] * col[(size_t)k];
                col[(size_t)i] = acc;
            }
            for (int64_t i = n - 1; i >= 0; --i) {
                T acc = col[(size_t)i];
                for (int64_t k = i + 1; k < n; ++k) acc -= lu[(size_t)(i * n + k)] * col[(size_t)k];
                col[(size_t)i] = acc / lu[(size_t)(i * n + i)];
            }
            for (int64_t i = 0; i < n; ++i) inv[(size_t)(i * n + j)] = col[(size_t)i];
        }
        for (int64_t i = 0; i < n; ++i)
            for (int64_t j = 0; j < n; ++j) tp[to + i * t->stride[nd - 2] + j * t->stride[nd - 1]] = inv[(size_t)(i * n + j)];
    }
    return DN_OK;
}

extern "C" {

dn_status dno_batched_invert(const dn_tensor *t, const dn_tensor *a) {
    if (!valid(t) || !valid(a)) return fail(DN_ERR_INVALID_ARG, "batched_invert: bad argument");
    if (t->dtype != a->dtype || (t->dtype != DN_F32 && t->dtype != DN_F64))
        return fail(DN_ERR_UNSUPPORTED, "this operation is only supported for floating point numbers");
    if (t->ndims < 2 || t->ndims != a->ndims || t->shape[t->ndims - 1] != t->shape[t->ndims - 2])
        return fail(DN_ERR_SHAPE_MISMATCH, "batched_invert: need [..., n, n]");
    for (int d = 0; d < t->ndims; ++d)
        if (t->shape[d] != a->shape[d]) return fail(DN_ERR_SHAPE_MISMATCH, "batched_invert: shapes differ");
    for (int d = 0; d < t->ndims; ++d)
        if (t->shape[d] == 0) return DN_OK;
    return t->dtype == DN_F32 ? invert_impl<float>(t, a) : invert_impl<double>(t, a);
}

}  // extern "C"

// Fused element-wise expression (dn_fused_elemwise, include/dn_tensor.h): the reference has no such member — the
// oracle defines it as EXACTLY the sequence of host operators it stands for: every instruction is unary_eval /
// binary_eval of the element type (ScalarOps.fs:381-533 semantics, intermediate results rounded to T).
extern "C" dn_status dno_fused_elemwise(const dn_tensor *t, const dn_tensor *const *srcs, int32_t nsrc,
                                        const dn_fused_instr *prog, int32_t ninstr) {
    if (!valid(t) || !srcs || !prog || nsrc < 1 || nsrc > DN_FUSED_MAX_SRCS || ninstr < 1 || ninstr > DN_FUSED_MAX_INSTRS)
        return fail(DN_ERR_INVALID_ARG, "fused_elemwise: bad argument");
    if (t->dtype != DN_F32 && t->dtype != DN_F64) return fail(DN_ERR_UNSUPPORTED, "fused_elemwise: f32/f64 only");
    for (int k = 0; k < nsrc; ++k) {
        if (!valid(srcs[k]) || srcs[k]->dtype != t->dtype) return fail(DN_ERR_INVALID_ARG, "fused_elemwise: dtype mismatch");
        if (!same_shape(t, srcs[k])) return fail(DN_ERR_SHAPE_MISMATCH, "fused_elemwise: shape mismatch");
    }
    uint32_t written = (1u << nsrc) - 1;
    for (int k = 0; k < ninstr; ++k) {
        const dn_fused_instr &in = prog[k];
        if (in.dst < 0 || in.dst >= DN_FUSED_REGS) return fail(DN_ERR_INVALID_ARG, "fused_elemwise: bad destination");
        const int nread = in.kind == DN_FUSED_CONST ? 0 : (in.kind == DN_FUSED_UNARY ? 1 : 2);
        if (in.kind == DN_FUSED_UNARY && (in.op < 0 || in.op > DN_TRUNCATE)) return fail(DN_ERR_UNSUPPORTED, "fused_elemwise: unary op");
        if (in.kind == DN_FUSED_BINARY && (in.op < 0 || in.op > DN_MIN_ELEMWISE)) return fail(DN_ERR_UNSUPPORTED, "fused_elemwise: binary op");
        if (in.kind < 0 || in.kind > DN_FUSED_CONST) return fail(DN_ERR_INVALID_ARG, "fused_elemwise: bad kind");
        const int regs[2] = {in.a, in.b};
        for (int q = 0; q < nread; ++q)
            if (regs[q] < 0 || regs[q] >= DN_FUSED_REGS || !((written >> regs[q]) & 1u))
                return fail(DN_ERR_INVALID_ARG, "fused_elemwise: register read before written");
        written |= 1u << in.dst;
    }
    auto run = [&](auto tag) {
        using T = decltype(tag);
        Operand ops[4] = {operand_of<T>(t), operand_of<T>(srcs[0]), operand_of<T>(srcs[nsrc > 1 ? 1 : 0]),
                          operand_of<T>(srcs[nsrc > 2 ? 2 : 0])};
        elemwise_drive<4>(t->ndims, t->shape, ops, false, true, [&](char **p) {
            T r[DN_FUSED_REGS] = {};
            for (int k = 0; k < nsrc; ++k) r[k] = *(const T *)p[1 + k];
            T z = T(0);
            for (int k = 0; k < ninstr; ++k) {
                const dn_fused_instr &in = prog[k];
                if (in.kind == DN_FUSED_CONST) z = (T)in.imm;
                else if (in.kind == DN_FUSED_UNARY) z = unary_eval<T>(in.op, r[in.a]);
                else z = binary_eval<T>(in.op, r[in.a], r[in.b]);
                r[in.dst] = z;
            }
            *(T *)p[0] = z;
        });
    };
    if (t->dtype == DN_F32) run(0.0f);
    else run(0.0);
    return DN_OK;
}
