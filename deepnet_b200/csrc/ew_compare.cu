// ew_compare.cu — dn_compare: Equal..GreaterOrEqual (TensorBackend.fs:106-111).
// Replaces CudaBackend.fs:336-341 and Kernels/Elemwise.cuh:223-272. The bool target is written 16 results per
// 128-bit store on the contiguous path (the reference writes one byte per thread).
#include "ew_ops.cuh"

using namespace dn;

namespace {
template <class T>
dn_status compare_typed(int op, EwPlan &plan) {
    switch (op) {
    case DN_EQUAL: return ew_run(plan, CompareF<T, DN_EQUAL>());
    case DN_NOT_EQUAL: return ew_run(plan, CompareF<T, DN_NOT_EQUAL>());
    case DN_LESS: return ew_run(plan, CompareF<T, DN_LESS>());
    case DN_LESS_OR_EQUAL: return ew_run(plan, CompareF<T, DN_LESS_OR_EQUAL>());
    case DN_GREATER: return ew_run(plan, CompareF<T, DN_GREATER>());
    default: return ew_run(plan, CompareF<T, DN_GREATER_OR_EQUAL>());
    }
}
}  // namespace

extern "C" dn_status dn_compare(int32_t op, const dn_tensor *t, const dn_tensor *a, const dn_tensor *b) {
    if (!tensor_valid(t) || !tensor_valid(a) || !tensor_valid(b) || op < 0 || op >= DN_COMPARE_OP_COUNT)
        return set_error(DN_ERR_INVALID_ARG, "compare: bad argument");
    if (t->dtype != DN_BOOL || a->dtype != b->dtype)
        return set_error(DN_ERR_INVALID_ARG, "compare: target must be bool and sources must have the same type");
    EwPlan plan;
    const dn_tensor *srcs[2] = {a, b};
    dn_status st = ew_make_plan(plan, t, srcs, 2);
    if (st != DN_OK || plan.n == 0) return st;
    DN_SWITCH_DTYPE(a->dtype, { return compare_typed<T>(op, plan); });
    return DN_OK;
}
