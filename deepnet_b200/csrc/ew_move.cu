// ew_move.cu — FillConst, FillIncrementing, Copy, IfThenElse, IsFinite entry points.
// Replaces CudaBackend.fs:272-298,342,344 and Kernels/Elemwise.cuh:13-43,134,223-308.
#include "ew_ops.cuh"

using namespace dn;

extern "C" {

dn_status dn_fill_const(const dn_tensor *t, const void *value) {
    if (!tensor_valid(t) || !value) return set_error(DN_ERR_INVALID_ARG, "FillConst: bad argument");
    EwPlan plan;
    dn_status st = ew_make_plan(plan, t, nullptr, 0);
    if (st != DN_OK || plan.n == 0) return st;
    DN_SWITCH_SIZE(dtype_size(t->dtype), {
        FillF<B> f;
        memcpy(&f.value, value, sizeof(B));
        if (t->dtype == DN_BOOL) f.value = f.value ? 1 : 0;
        return ew_run(plan, f);
    });
    return DN_OK;
}

// pos[0] of the host's SWAPPED layout (HostBackend.ElemwiseLayouts, HostBackend.fs:139-153, applied by
// FillIncrementing at HostBackend.fs:192-194): the largest dim with target stride 1 is swapped to the end, so
// "dim 0" is the original dim 0 unless that very dim was the one swapped away.
static int fill_incrementing_index_dim(const dn_tensor *t) {
    const int nd = t->ndims;
    if (nd == 0) return -1;
    int best = -1;
    int64_t best_size = -1;
    for (int d = 0; d < nd; ++d)
        if (t->stride[d] == 1 && t->shape[d] > best_size) {
            best = d;
            best_size = t->shape[d];
        }
    if (best == 0 && nd > 1) return nd - 1;  // dims 0 and nd-1 trade places
    return 0;
}

dn_status dn_fill_incrementing(const dn_tensor *t, const void *start, const void *incr) {
    if (!tensor_valid(t) || !start || !incr) return set_error(DN_ERR_INVALID_ARG, "FillIncrementing: bad argument");
    if (t->dtype == DN_BOOL) return set_error(DN_ERR_UNSUPPORTED, "FillIncrementing is not defined for bool");
    EwPlan plan;
    const dn_tensor *srcs[1] = {nullptr};
    dn_status st = ew_make_plan(plan, t, srcs, 1, /*index_operand=*/0, fill_incrementing_index_dim(t));
    if (st != DN_OK || plan.n == 0) return st;
    DN_SWITCH_DTYPE(t->dtype, {
        if constexpr (!kIsBool<T>) {
            FillIncrF<T> f;
            memcpy(&f.start, start, sizeof(T));
            memcpy(&f.incr, incr, sizeof(T));
            return ew_run(plan, f);
        }
    });
    return DN_OK;
}

dn_status dn_copy(const dn_tensor *t, const dn_tensor *a) {
    if (!tensor_valid(t) || !tensor_valid(a)) return set_error(DN_ERR_INVALID_ARG, "Copy: bad argument");
    if (t->dtype != a->dtype) return set_error(DN_ERR_INVALID_ARG, "Copy: source and target types differ");
    EwPlan plan;
    const dn_tensor *srcs[1] = {a};
    dn_status st = ew_make_plan(plan, t, srcs, 1);
    if (st != DN_OK || plan.n == 0) return st;
    DN_SWITCH_SIZE(dtype_size(t->dtype), { return ew_run(plan, CopyF<B>()); });
    return DN_OK;
}

dn_status dn_if_then_else(const dn_tensor *t, const dn_tensor *cond, const dn_tensor *if_true,
                          const dn_tensor *if_false) {
    if (!tensor_valid(t) || !tensor_valid(cond) || !tensor_valid(if_true) || !tensor_valid(if_false))
        return set_error(DN_ERR_INVALID_ARG, "IfThenElse: bad argument");
    if (cond->dtype != DN_BOOL || t->dtype != if_true->dtype || t->dtype != if_false->dtype)
        return set_error(DN_ERR_INVALID_ARG, "IfThenElse: cond must be bool and values must have the target's type");
    EwPlan plan;
    const dn_tensor *srcs[3] = {cond, if_true, if_false};
    dn_status st = ew_make_plan(plan, t, srcs, 3);
    if (st != DN_OK || plan.n == 0) return st;
    DN_SWITCH_SIZE(dtype_size(t->dtype), { return ew_run(plan, SelectF<B>()); });
    return DN_OK;
}

dn_status dn_is_finite(const dn_tensor *t, const dn_tensor *a) {
    if (!tensor_valid(t) || !tensor_valid(a)) return set_error(DN_ERR_INVALID_ARG, "IsFinite: bad argument");
    if (t->dtype != DN_BOOL) return set_error(DN_ERR_INVALID_ARG, "IsFinite: target must be bool");
    EwPlan plan;
    const dn_tensor *srcs[1] = {a};
    dn_status st = ew_make_plan(plan, t, srcs, 1);
    if (st != DN_OK || plan.n == 0) return st;
    if (a->dtype == DN_F32) return ew_run(plan, IsFiniteF<float>());
    if (a->dtype == DN_F64) return ew_run(plan, IsFiniteF<double>());
    // integers and bool are always finite (ScalarPrimitives.fs:177-184): fill with true
    EwPlan fill;
    st = ew_make_plan(fill, t, nullptr, 0);
    if (st != DN_OK) return st;
    FillF<uint8_t> f;
    f.value = 1;
    return ew_run(fill, f);
}

}  // extern "C"
