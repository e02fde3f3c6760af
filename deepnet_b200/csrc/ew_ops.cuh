// ew_ops.cuh — element-wise functors. Semantics follow the reference's HOST backend wherever host and the
// reference CUDA kernels disagree: Tensor/Tensor/ScalarPrimitives.fs:50-246, Tensor/Tensor/Sgn.fs:8-31,
// Tensor/Tensor/Host/VectorOps.fs:201-240. (Reference device code being replaced: Kernels/Elemwise.cuh:70-308.)
//
// Compiled WITHOUT --use_fast_math: +,-,*,/ and sqrt are IEEE-rounded (bit-exact vs. host); transcendental
// functions use CUDA's float/double libm (<= 2-4 ulp), inside the rel 1e-5 tolerance north_star allows against
// the host's evaluate-in-double-then-round results.
#pragma once

#include "elemwise.cuh"

namespace dn {

template <class T> constexpr bool kIsFloat = std::is_floating_point<T>::value;
template <class T> constexpr bool kIsBool = std::is_same<T, bool8>::value;
template <class T> constexpr bool kIsSigned = std::is_integral<T>::value && std::is_signed<T>::value;
template <class T> constexpr bool kIsUnsigned = std::is_integral<T>::value && std::is_unsigned<T>::value;
template <class T> constexpr bool kIsInt = std::is_integral<T>::value;

template <class T> struct MakeUnsigned { using type = typename std::make_unsigned<T>::type; };
template <> struct MakeUnsigned<float> { using type = float; };
template <> struct MakeUnsigned<double> { using type = double; };
template <> struct MakeUnsigned<bool8> { using type = bool8; };
template <class T> using UnsignedT = typename MakeUnsigned<T>::type;

template <class T> struct Limits;
#define DN_LIMITS(T, LO, HI)                                        \
    template <> struct Limits<T> {                                  \
        __host__ __device__ static constexpr T lowest() { return LO; } \
        __host__ __device__ static constexpr T max() { return HI; }    \
    };
DN_LIMITS(float, -3.402823466e+38f, 3.402823466e+38f)
DN_LIMITS(double, -1.7976931348623157e+308, 1.7976931348623157e+308)
DN_LIMITS(int8_t, INT8_MIN, INT8_MAX)
DN_LIMITS(uint8_t, 0, UINT8_MAX)
DN_LIMITS(int16_t, INT16_MIN, INT16_MAX)
DN_LIMITS(uint16_t, 0, UINT16_MAX)
DN_LIMITS(int32_t, INT32_MIN, INT32_MAX)
DN_LIMITS(uint32_t, 0u, UINT32_MAX)
DN_LIMITS(int64_t, INT64_MIN, INT64_MAX)
DN_LIMITS(uint64_t, 0ull, UINT64_MAX)
#undef DN_LIMITS

// Four bool bytes per 32-bit word: 0x01 in every byte whose input byte is non-zero (bools are normalised on load,
// KernelCompiler.fs:190-193 marshals them as one byte but foreign memory may hold any non-zero value).
__device__ __forceinline__ uint32_t bool4_norm(uint32_t w) { return __vcmpne4(w, 0u) & 0x01010101u; }

// ---- unary ---------------------------------------------------------------------------------------------------
template <class T, int OP>
struct UnaryF : EwSig<T, T> {
    static constexpr bool Packed = kIsBool<T>;
    __device__ __forceinline__ uint32_t packed(uint32_t a) const { return __vcmpeq4(a, 0u) & 0x01010101u; }  // Negate
    __device__ __forceinline__ T operator()(T x) const {
        if constexpr (kIsBool<T>) {
            return bool8(!bool(x));  // Negate, ScalarOps.fs:491-493
        } else if constexpr (kIsFloat<T>) {
            if constexpr (OP == DN_UNARY_MINUS) return -x;
            else if constexpr (OP == DN_ABS) return fabs(x);
            else if constexpr (OP == DN_SGN) return x < T(0) ? T(-1) : (x > T(0) ? T(1) : T(0));  // Sgn(NaN)=0
            else if constexpr (OP == DN_LOG) return log(x);
            else if constexpr (OP == DN_LOG10) return log10(x);
            else if constexpr (OP == DN_EXP) return exp(x);
            else if constexpr (OP == DN_SIN) return sin(x);
            else if constexpr (OP == DN_COS) return cos(x);
            else if constexpr (OP == DN_TAN) return tan(x);
            else if constexpr (OP == DN_ASIN) return asin(x);
            else if constexpr (OP == DN_ACOS) return acos(x);
            else if constexpr (OP == DN_ATAN) return atan(x);
            else if constexpr (OP == DN_SINH) return sinh(x);
            else if constexpr (OP == DN_COSH) return cosh(x);
            else if constexpr (OP == DN_TANH) return tanh(x);
            else if constexpr (OP == DN_SQRT) return sqrt(x);
            else if constexpr (OP == DN_CEILING) return ceil(x);
            else if constexpr (OP == DN_FLOOR) return floor(x);
            else if constexpr (OP == DN_ROUND) return rint(x);  // Math.Round: half to even
            else if constexpr (OP == DN_TRUNCATE) return trunc(x);
            else return x;
        } else {
            using U = UnsignedT<T>;
            if constexpr (OP == DN_UNARY_MINUS) return (T)(U(0) - (U)x);  // wraps (Vector.Negate)
            else if constexpr (OP == DN_ABS) {
                if constexpr (kIsSigned<T>) return x < 0 ? (T)(U(0) - (U)x) : x;  // wraps at MinValue (Vector.Abs)
                else return x;
            } else if constexpr (OP == DN_SGN) return x < 0 ? T(-1) : (x > 0 ? T(1) : T(0));
            else return x;
        }
    }
};

// ---- binary --------------------------------------------------------------------------------------------------
template <class T, int OP>
struct BinaryF : EwSig<T, T, T> {
    static constexpr bool Packed = kIsBool<T>;
    __device__ __forceinline__ uint32_t packed(uint32_t a, uint32_t b) const {
        if constexpr (OP == DN_AND) return bool4_norm(a) & bool4_norm(b);
        else if constexpr (OP == DN_OR) return bool4_norm(a | b);
        else return bool4_norm(a) ^ bool4_norm(b);
    }
    __device__ __forceinline__ T operator()(T a, T b) const {
        if constexpr (kIsBool<T>) {
            if constexpr (OP == DN_AND) return bool8(bool(a) && bool(b));
            else if constexpr (OP == DN_OR) return bool8(bool(a) || bool(b));
            else return bool8(bool(a) != bool(b));
        } else if constexpr (kIsFloat<T>) {
            if constexpr (OP == DN_ADD) return a + b;
            else if constexpr (OP == DN_SUBTRACT) return a - b;
            else if constexpr (OP == DN_MULTIPLY) return a * b;
            else if constexpr (OP == DN_DIVIDE) return a / b;
            else if constexpr (OP == DN_MODULO) return fmod(a, b);
            else if constexpr (OP == DN_POWER) return pow(a, b);
            else if constexpr (OP == DN_MAX_ELEMWISE) return a > b ? a : b;  // ScalarOps.fs:525-528
            else return a < b ? a : b;
        } else {
            using U = UnsignedT<T>;
            if constexpr (OP == DN_ADD) return (T)((U)a + (U)b);
            else if constexpr (OP == DN_SUBTRACT) return (T)((U)a - (U)b);
            else if constexpr (OP == DN_MULTIPLY) {
                if constexpr (sizeof(T) < 4) return (T)((uint32_t)(U)a * (uint32_t)(U)b);
                else return (T)((U)a * (U)b);
            } else if constexpr (OP == DN_DIVIDE) {
                // host throws on x/0 and MinValue/-1 (outside the parity domain); defined here as 0 / wrap
                if (b == 0) return T(0);
                if constexpr (kIsSigned<T>)
                    if (a == Limits<T>::lowest() && b == T(-1)) return a;
                return (T)(a / b);
            } else if constexpr (OP == DN_MODULO) {
                if (b == 0) return T(0);
                if constexpr (kIsSigned<T>)
                    if (a == Limits<T>::lowest() && b == T(-1)) return T(0);
                return (T)(a % b);
            } else if constexpr (OP == DN_MAX_ELEMWISE) return a > b ? a : b;
            else return a < b ? a : b;
        }
    }
};

// ---- comparisons ---------------------------------------------------------------------------------------------
template <class T> __device__ __forceinline__ auto cmp_value(T v) {
    if constexpr (kIsBool<T>) return (int)(v.v != 0);
    else return v;
}

template <class T, int OP>
struct CompareF : EwSig<bool8, T, T> {
    __device__ __forceinline__ bool8 operator()(T a0, T b0) const {
        const auto a = cmp_value(a0), b = cmp_value(b0);
        if constexpr (OP == DN_EQUAL) return bool8(a == b);
        else if constexpr (OP == DN_NOT_EQUAL) return bool8(a != b);
        else if constexpr (OP == DN_LESS) return bool8(a < b);
        else if constexpr (OP == DN_LESS_OR_EQUAL) return bool8(a <= b);
        else if constexpr (OP == DN_GREATER) return bool8(a > b);
        else return bool8(a >= b);
    }
};

template <class T>
struct IsFiniteF : EwSig<bool8, T> {
    __device__ __forceinline__ bool8 operator()(T x) const {
        if constexpr (kIsFloat<T>) return bool8(isfinite(x));
        else return bool8(true);
    }
};

// ---- data movement (by element size: B is an unsigned integer of the element's width) --------------------------
template <class B>
struct CopyF : EwSig<B, B> {
    __device__ __forceinline__ B operator()(B x) const { return x; }
};

template <class B>
struct FillF : EwSig<B> {
    B value;
    __device__ __forceinline__ B operator()() const { return value; }
};

template <class B>
struct SelectF : EwSig<B, bool8, B, B> {
    __device__ __forceinline__ B operator()(bool8 c, B t, B f) const { return bool(c) ? t : f; }
};

// FillIncrementing (ScalarOps.fs:367-370): start + incr * conv(pos0), two roundings for floats (never fused).
template <class T>
struct FillIncrF : EwSig<T, IndexT> {
    static constexpr bool Tiled = false;
    T start, incr;
    __device__ __forceinline__ T operator()(int64_t pos) const {
        if constexpr (std::is_same<T, float>::value) return __fadd_rn(start, __fmul_rn(incr, (float)pos));
        else if constexpr (std::is_same<T, double>::value) return __dadd_rn(start, __dmul_rn(incr, (double)pos));
        else {
            using U = UnsignedT<T>;
            if constexpr (sizeof(T) < 4) return (T)((uint32_t)(U)start + (uint32_t)(U)incr * (uint32_t)(U)(T)pos);
            else return (T)((U)start + (U)incr * (U)(T)pos);
        }
    }
};

// Convert (ScalarPrimitives.fs:50-52, Elemwise.cuh:50-63): unchecked cast; float->narrow int goes through int32.
template <class Tt, class Ts>
struct ConvertF : EwSig<Tt, Ts> {
    static constexpr bool Tiled = sizeof(Tt) == sizeof(Ts) || (sizeof(Tt) >= 4 && sizeof(Ts) >= 4);
    __device__ __forceinline__ Tt operator()(Ts v) const {
        if constexpr (kIsBool<Tt>) {
            if constexpr (kIsBool<Ts>) return v;
            else return bool8(v != Ts(0));
        } else if constexpr (kIsBool<Ts>) {
            return (Tt)(bool(v) ? 1 : 0);
        } else if constexpr (kIsFloat<Ts> && kIsInt<Tt> && sizeof(Tt) < 4) {
            return (Tt)(int32_t)v;
        } else if constexpr (kIsFloat<Ts> && std::is_same<Tt, uint32_t>::value) {
            return (Tt)(int64_t)v;
        } else {
            return (Tt)v;
        }
    }
};

template <int BYTES> struct BitsOf;
template <> struct BitsOf<1> { using type = uint8_t; };
template <> struct BitsOf<2> { using type = uint16_t; };
template <> struct BitsOf<4> { using type = uint32_t; };
template <> struct BitsOf<8> { using type = uint64_t; };

// dtype dispatch helper: calls fn(T{}) with the C type of `dt`.
#define DN_SWITCH_DTYPE(DT, ...)                                                \
    switch (DT) {                                                               \
    case DN_F32: { using T = float; __VA_ARGS__; } break;                       \
    case DN_F64: { using T = double; __VA_ARGS__; } break;                      \
    case DN_I8: { using T = int8_t; __VA_ARGS__; } break;                       \
    case DN_U8: { using T = uint8_t; __VA_ARGS__; } break;                      \
    case DN_I16: { using T = int16_t; __VA_ARGS__; } break;                     \
    case DN_U16: { using T = uint16_t; __VA_ARGS__; } break;                    \
    case DN_I32: { using T = int32_t; __VA_ARGS__; } break;                     \
    case DN_U32: { using T = uint32_t; __VA_ARGS__; } break;                    \
    case DN_I64: { using T = int64_t; __VA_ARGS__; } break;                     \
    case DN_U64: { using T = uint64_t; __VA_ARGS__; } break;                    \
    case DN_BOOL: { using T = bool8; __VA_ARGS__; } break;                      \
    default: return set_error(DN_ERR_INVALID_ARG, "bad dtype %d", (int)(DT));   \
    }

#define DN_SWITCH_SIZE(BYTES, ...)                                              \
    switch (BYTES) {                                                            \
    case 1: { using B = uint8_t; __VA_ARGS__; } break;                          \
    case 2: { using B = uint16_t; __VA_ARGS__; } break;                         \
    case 4: { using B = uint32_t; __VA_ARGS__; } break;                         \
    case 8: { using B = uint64_t; __VA_ARGS__; } break;                         \
    default: return set_error(DN_ERR_INVALID_ARG, "bad element size");          \
    }

}  // namespace dn
