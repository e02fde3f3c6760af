// ew_unary.cu — dn_unary: UnaryPlus..Truncate, Negate (TensorBackend.fs:74-94,113).
// Replaces CudaBackend.fs:305-325,346 and Kernels/Elemwise.cuh:70-156. Support matrix follows the HOST backend
// (ScalarPrimitives.fs:58-155, Sgn.fs:8-31, VectorOps.fs:205-215): transcendental and rounding ops exist for
// f32/f64 only; Sgn for i16/i32/i64/f32/f64; UnaryMinus/Abs for every numeric type (wrapping); Negate for bool.
#include "ew_ops.cuh"

using namespace dn;

namespace {

template <class T, int OP>
dn_status run_unary(EwPlan &plan) {
    return ew_run(plan, UnaryF<T, OP>());
}

template <class T>
dn_status unary_float(int op, EwPlan &plan) {
    switch (op) {
    case DN_UNARY_MINUS: return run_unary<T, DN_UNARY_MINUS>(plan);
    case DN_ABS: return run_unary<T, DN_ABS>(plan);
    case DN_SGN: return run_unary<T, DN_SGN>(plan);
    case DN_LOG: return run_unary<T, DN_LOG>(plan);
    case DN_LOG10: return run_unary<T, DN_LOG10>(plan);
    case DN_EXP: return run_unary<T, DN_EXP>(plan);
    case DN_SIN: return run_unary<T, DN_SIN>(plan);
    case DN_COS: return run_unary<T, DN_COS>(plan);
    case DN_TAN: return run_unary<T, DN_TAN>(plan);
    case DN_ASIN: return run_unary<T, DN_ASIN>(plan);
    case DN_ACOS: return run_unary<T, DN_ACOS>(plan);
    case DN_ATAN: return run_unary<T, DN_ATAN>(plan);
    case DN_SINH: return run_unary<T, DN_SINH>(plan);
    case DN_COSH: return run_unary<T, DN_COSH>(plan);
    case DN_TANH: return run_unary<T, DN_TANH>(plan);
    case DN_SQRT: return run_unary<T, DN_SQRT>(plan);
    case DN_CEILING: return run_unary<T, DN_CEILING>(plan);
    case DN_FLOOR: return run_unary<T, DN_FLOOR>(plan);
    case DN_ROUND: return run_unary<T, DN_ROUND>(plan);
    case DN_TRUNCATE: return run_unary<T, DN_TRUNCATE>(plan);
    default: return set_error(DN_ERR_UNSUPPORTED, "unary op %d is not defined for floating point tensors", op);
    }
}

template <class T>
dn_status unary_int(int op, int dtype, EwPlan &plan) {
    switch (op) {
    case DN_UNARY_MINUS: return run_unary<T, DN_UNARY_MINUS>(plan);
    case DN_ABS:
        if constexpr (kIsSigned<T>) return run_unary<T, DN_ABS>(plan);
        else return ew_run(plan, CopyF<UnsignedT<T>>());  // |x| = x for unsigned (Vector.Abs)
    case DN_SGN:
        if constexpr (std::is_same<T, int16_t>::value || std::is_same<T, int32_t>::value ||
                      std::is_same<T, int64_t>::value)
            return run_unary<T, DN_SGN>(plan);
        else
            return set_error(DN_ERR_UNSUPPORTED, "Sgn is not defined for type %s", dtype_name(dtype));
    default:
        return set_error(DN_ERR_UNSUPPORTED, "unary op %d is not defined for type %s", op, dtype_name(dtype));
    }
}

}  // namespace

extern "C" dn_status dn_unary(int32_t op, const dn_tensor *t, const dn_tensor *a) {
    if (!tensor_valid(t) || !tensor_valid(a) || op < 0 || op >= DN_UNARY_OP_COUNT)
        return set_error(DN_ERR_INVALID_ARG, "unary: bad argument");
    if (t->dtype != a->dtype) return set_error(DN_ERR_INVALID_ARG, "unary: source and target types differ");
    if ((op == DN_NEGATE) != (t->dtype == DN_BOOL))
        return set_error(DN_ERR_UNSUPPORTED, "unary op %d is not defined for type %s", op, dtype_name(t->dtype));
    EwPlan plan;
    const dn_tensor *srcs[1] = {a};
    dn_status st = ew_make_plan(plan, t, srcs, 1);
    if (st != DN_OK || plan.n == 0) return st;
    if (op == DN_UNARY_PLUS) {  // identity: a copy by element size
        DN_SWITCH_SIZE(dtype_size(t->dtype), { return ew_run(plan, CopyF<B>()); });
    }
    switch (t->dtype) {
    case DN_F32: return unary_float<float>(op, plan);
    case DN_F64: return unary_float<double>(op, plan);
    case DN_I8: return unary_int<int8_t>(op, t->dtype, plan);
    case DN_U8: return unary_int<uint8_t>(op, t->dtype, plan);
    case DN_I16: return unary_int<int16_t>(op, t->dtype, plan);
    case DN_U16: return unary_int<uint16_t>(op, t->dtype, plan);
    case DN_I32: return unary_int<int32_t>(op, t->dtype, plan);
    case DN_U32: return unary_int<uint32_t>(op, t->dtype, plan);
    case DN_I64: return unary_int<int64_t>(op, t->dtype, plan);
    case DN_U64: return unary_int<uint64_t>(op, t->dtype, plan);
    case DN_BOOL: return run_unary<bool8, DN_NEGATE>(plan);
    default: return set_error(DN_ERR_INVALID_ARG, "bad dtype");
    }
}
