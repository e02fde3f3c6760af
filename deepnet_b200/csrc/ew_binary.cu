// ew_binary.cu — dn_binary: Add..MinElemwise, And/Or/Xor (TensorBackend.fs:97-104,114-116).
// Replaces CudaBackend.fs:327-334,347-349 and Kernels/Elemwise.cuh:164-216.
#include "ew_ops.cuh"

using namespace dn;

namespace {

template <class T>
dn_status binary_numeric(int op, int dtype, EwPlan &plan) {
    switch (op) {
    case DN_ADD: return ew_run(plan, BinaryF<T, DN_ADD>());
    case DN_SUBTRACT: return ew_run(plan, BinaryF<T, DN_SUBTRACT>());
    case DN_MULTIPLY: return ew_run(plan, BinaryF<T, DN_MULTIPLY>());
    case DN_DIVIDE: return ew_run(plan, BinaryF<T, DN_DIVIDE>());
    case DN_MODULO: return ew_run(plan, BinaryF<T, DN_MODULO>());
    case DN_MAX_ELEMWISE: return ew_run(plan, BinaryF<T, DN_MAX_ELEMWISE>());
    case DN_MIN_ELEMWISE: return ew_run(plan, BinaryF<T, DN_MIN_ELEMWISE>());
    case DN_POWER:
        if constexpr (kIsFloat<T>) return ew_run(plan, BinaryF<T, DN_POWER>());
        else return set_error(DN_ERR_UNSUPPORTED, "Power is not defined for type %s", dtype_name(dtype));
    default:
        return set_error(DN_ERR_UNSUPPORTED, "binary op %d is not defined for type %s", op, dtype_name(dtype));
    }
}

}  // namespace

extern "C" dn_status dn_binary(int32_t op, const dn_tensor *t, const dn_tensor *a, const dn_tensor *b) {
    if (!tensor_valid(t) || !tensor_valid(a) || !tensor_valid(b) || op < 0 || op >= DN_BINARY_OP_COUNT)
        return set_error(DN_ERR_INVALID_ARG, "binary: bad argument");
    if (t->dtype != a->dtype || t->dtype != b->dtype)
        return set_error(DN_ERR_INVALID_ARG, "binary: source and target types differ");
    EwPlan plan;
    const dn_tensor *srcs[2] = {a, b};
    dn_status st = ew_make_plan(plan, t, srcs, 2);
    if (st != DN_OK || plan.n == 0) return st;
    if (t->dtype == DN_BOOL) {
        switch (op) {
        case DN_AND: return ew_run(plan, BinaryF<bool8, DN_AND>());
        case DN_OR: return ew_run(plan, BinaryF<bool8, DN_OR>());
        case DN_XOR: return ew_run(plan, BinaryF<bool8, DN_XOR>());
        default: return set_error(DN_ERR_UNSUPPORTED, "binary op %d is not defined for type bool", op);
        }
    }
    if (op >= DN_AND) return set_error(DN_ERR_UNSUPPORTED, "logic ops are only defined for type bool");
    switch (t->dtype) {
    case DN_F32: return binary_numeric<float>(op, t->dtype, plan);
    case DN_F64: return binary_numeric<double>(op, t->dtype, plan);
    case DN_I8: return binary_numeric<int8_t>(op, t->dtype, plan);
    case DN_U8: return binary_numeric<uint8_t>(op, t->dtype, plan);
    case DN_I16: return binary_numeric<int16_t>(op, t->dtype, plan);
    case DN_U16: return binary_numeric<uint16_t>(op, t->dtype, plan);
    case DN_I32: return binary_numeric<int32_t>(op, t->dtype, plan);
    case DN_U32: return binary_numeric<uint32_t>(op, t->dtype, plan);
    case DN_I64: return binary_numeric<int64_t>(op, t->dtype, plan);
    case DN_U64: return binary_numeric<uint64_t>(op, t->dtype, plan);
    default: return set_error(DN_ERR_INVALID_ARG, "bad dtype");
    }
}
