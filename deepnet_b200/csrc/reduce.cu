// reduce.cu — planning for last-axis reductions (host side) and the fold entry point dn_reduce_last_axis.
// Replaces CudaBackend.fs:351-356,489 and CudaKernels.fs:104-110,263-285 (initial values passed as scalar args).
#include "reduce.cuh"

#include <algorithm>

namespace dn {

dn_status red_make_plan(RedPlan &plan, const dn_tensor *t, const dn_tensor *a, const char *what) {
    if (!tensor_valid(t) || !tensor_valid(a)) return set_error(DN_ERR_INVALID_ARG, "%s: bad argument", what);
    if (a->ndims < 1 || t->ndims != a->ndims - 1)
        return set_error(DN_ERR_SHAPE_MISMATCH, "%s: source rank must be target rank + 1", what);
    for (int d = 0; d < t->ndims; ++d)
        if (t->shape[d] != a->shape[d])
            return set_error(DN_ERR_SHAPE_MISMATCH, "%s: target shape does not match the source's leading dims", what);
    const int nd = a->ndims;
    plan.src = data_ptr(a);
    plan.dst = data_ptr(t);
    plan.in_size = dtype_size(a->dtype);
    plan.out_size = dtype_size(t->dtype);
    plan.len = a->shape[nd - 1];
    plan.lstride_elems = a->stride[nd - 1];
    struct Dim { int64_t size, ss, ts; };
    Dim dims[DN_MAX_DIMS];
    int n = 0;
    int64_t rows = 1;
    for (int d = nd - 2; d >= 0; --d) rows *= a->shape[d];
    plan.nrows = rows;
    if (rows == 0) return DN_OK;  // (row-major strides in front of a zero-sized dim are 0: nothing to validate)
    for (int d = nd - 2; d >= 0; --d) {  // innermost-first
        if (a->shape[d] == 1) continue;
        if (t->stride[d] == 0) return set_error(DN_ERR_INVALID_ARG, "%s: the target must not be a broadcast view", what);
        dims[n++] = Dim{a->shape[d], a->stride[d], t->stride[d]};
    }
    std::stable_sort(dims, dims + n, [](const Dim &x, const Dim &y) {
        const int64_t ax = x.ss < 0 ? -x.ss : x.ss, ay = y.ss < 0 ? -y.ss : y.ss;
        return ax < ay;
    });
    int m = 0;
    for (int i = 0; i < n; ++i) {
        if (m > 0 && dims[i].ss == dims[m - 1].ss * dims[m - 1].size && dims[i].ts == dims[m - 1].ts * dims[m - 1].size) {
            dims[m - 1].size *= dims[i].size;
            continue;
        }
        dims[m++] = dims[i];
    }
    plan.nouter = m;
    for (int d = 0; d < m; ++d) {
        plan.oshape[d] = dims[d].size;
        plan.ostride_s[d] = dims[d].ss;
        plan.ostride_t[d] = dims[d].ts;
        if (dims[d].size >= ((int64_t)1 << 31))
            return set_error(DN_ERR_UNSUPPORTED, "%s: outer extent exceeds 2^31-1", what);
    }
    return DN_OK;
}

bool red_cols_can_vectorize(const RedPlan &plan, int vec) {
    if (vec <= 1 || plan.nouter < 1) return false;
    if (plan.ostride_s[0] != 1 || plan.oshape[0] % vec != 0) return false;
    const int64_t align = (int64_t)vec * plan.in_size;  // <= 16
    if (reinterpret_cast<uintptr_t>(plan.src) % align != 0) return false;
    if ((plan.lstride_elems * plan.in_size) % align != 0) return false;
    for (int d = 1; d < plan.nouter; ++d)
        if ((plan.ostride_s[d] * plan.in_size) % align != 0) return false;
    return true;
}

void red_fill_outer(RedOuter &o, const RedPlan &plan) {
    o.ndims = plan.nouter;
    for (int d = 0; d < DN_MAX_DIMS; ++d) {
        const bool on = d < plan.nouter;
        o.shape[d] = on ? (uint32_t)plan.oshape[d] : 1;
        o.div[d].init(o.shape[d]);
        o.sstride[d] = on ? plan.ostride_s[d] * plan.in_size : 0;
        o.tstride[d] = on ? plan.ostride_t[d] * plan.out_size : 0;
    }
}

}  // namespace dn

using namespace dn;

namespace {

template <class T>
dn_status fold_numeric(int op, const RedPlan &plan) {
    switch (op) {
    case DN_SUM: return red_run(plan, SumOp<T>());
    case DN_PRODUCT: return red_run(plan, ProductOp<T>());
    case DN_MIN:
        if constexpr (kIsFloat<T>) return red_run(plan, MinMaxFloatOp<T, false>());
        else return red_run(plan, MinMaxIntOp<T, false>());
    default:
        if constexpr (kIsFloat<T>) return red_run(plan, MinMaxFloatOp<T, true>());
        else return red_run(plan, MinMaxIntOp<T, true>());
    }
}

}  // namespace

extern "C" dn_status dn_reduce_last_axis(int32_t op, const dn_tensor *t, const dn_tensor *a) {
    if (op < 0 || op >= DN_REDUCE_OP_COUNT) return set_error(DN_ERR_INVALID_ARG, "reduce: bad op %d", op);
    RedPlan plan;
    dn_status st = red_make_plan(plan, t, a, "reduce");
    if (st != DN_OK) return st;
    if (op == DN_COUNT_TRUE) {
        if (a->dtype != DN_BOOL || t->dtype != DN_I64)
            return set_error(DN_ERR_INVALID_ARG, "CountTrueLastAxis: source must be bool and target int64");
        return red_run(plan, CountTrueOp());
    }
    if (op == DN_ALL || op == DN_ANY) {
        if (a->dtype != DN_BOOL || t->dtype != DN_BOOL)
            return set_error(DN_ERR_INVALID_ARG, "All/AnyLastAxis: source and target must be bool");
        return op == DN_ALL ? red_run(plan, AllAnyOp<true>()) : red_run(plan, AllAnyOp<false>());
    }
    if (t->dtype != a->dtype) return set_error(DN_ERR_INVALID_ARG, "reduce: source and target types differ");
    switch (a->dtype) {
    case DN_F32: return fold_numeric<float>(op, plan);
    case DN_F64: return fold_numeric<double>(op, plan);
    case DN_I8: return fold_numeric<int8_t>(op, plan);
    case DN_U8: return fold_numeric<uint8_t>(op, plan);
    case DN_I16: return fold_numeric<int16_t>(op, plan);
    case DN_U16: return fold_numeric<uint16_t>(op, plan);
    case DN_I32: return fold_numeric<int32_t>(op, plan);
    case DN_U32: return fold_numeric<uint32_t>(op, plan);
    case DN_I64: return fold_numeric<int64_t>(op, plan);
    case DN_U64: return fold_numeric<uint64_t>(op, plan);
    default: return set_error(DN_ERR_UNSUPPORTED, "numeric folds are not defined for type bool");
    }
}
