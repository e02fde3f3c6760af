// reduce_arg.cu — dn_arg_reduce_last_axis, dn_find_last_axis, dn_arg_reduce_combine.
// Replaces CudaBackend.fs:358-360, CudaKernels.fs:113-126,287-296 and Kernels/Reduction.cuh:79-159 (whose initial
// position is -1; the host's NotFound is followed here, SURVEY.md §8c rule 4).
#include "reduce.cuh"

using namespace dn;

namespace {

// Per-output fold of `nparts` (value, global index) partials gathered from the ranks of a leading-axis shard.
template <class T, bool IsMax>
__global__ void arg_combine_kernel(int64_t *out, const T *vals, const int64_t *idxs, int64_t n, int nparts,
                                   int64_t out_stride, int64_t part_stride_v, int64_t part_stride_i,
                                   int64_t elem_stride_v, int64_t elem_stride_i) {
    using Op = ArgOp<T, IsMax>;
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n) return;
    typename Op::State acc = Op::identity();
    for (int k = 0; k < nparts; ++k) {
        typename Op::State s;
        s.val = vals[k * part_stride_v + o * elem_stride_v];
        s.idx = idxs[k * part_stride_i + o * elem_stride_i];
        if (s.idx == (int64_t)DN_NOT_FOUND) continue;
        acc = Op::combine(acc, s);
    }
    out[o * out_stride] = acc.idx;
}

template <class T>
dn_status arg_combine_typed(int op, const dn_tensor *t, const dn_tensor *vals, const dn_tensor *idxs) {
    const int64_t n = vals->shape[1];
    const int nparts = (int)vals->shape[0];
    if (n == 0) return DN_OK;
    const unsigned grid = (unsigned)((n + 255) / 256);
    int64_t *out = reinterpret_cast<int64_t *>(data_ptr(t));
    const T *v = reinterpret_cast<const T *>(data_ptr(vals));
    const int64_t *ix = reinterpret_cast<const int64_t *>(data_ptr(idxs));
    const int64_t os = t->ndims ? t->stride[0] : 0;
    if (op == DN_ARG_MAX)
        DN_LAUNCH((arg_combine_kernel<T, true>), grid, 256, 0, out, v, ix, n, nparts, os, vals->stride[0],
                  idxs->stride[0], vals->stride[1], idxs->stride[1]);
    else
        DN_LAUNCH((arg_combine_kernel<T, false>), grid, 256, 0, out, v, ix, n, nparts, os, vals->stride[0],
                  idxs->stride[0], vals->stride[1], idxs->stride[1]);
    return launch_status("arg-reduce combine kernel");
}

}  // namespace

extern "C" {

dn_status dn_arg_reduce_last_axis(int32_t op, const dn_tensor *t, const dn_tensor *a) {
    if (op != DN_ARG_MIN && op != DN_ARG_MAX) return set_error(DN_ERR_INVALID_ARG, "arg-reduce: bad op %d", op);
    RedPlan plan;
    dn_status st = red_make_plan(plan, t, a, "arg-reduce");
    if (st != DN_OK) return st;
    if (t->dtype != DN_I64) return set_error(DN_ERR_INVALID_ARG, "arg-reduce: target must be int64");
    if (a->dtype == DN_BOOL) return set_error(DN_ERR_UNSUPPORTED, "ArgMin/ArgMax are not defined for type bool");
    DN_SWITCH_DTYPE(a->dtype, {
        if constexpr (!kIsBool<T>) {
            if (op == DN_ARG_MAX) return red_run(plan, ArgOp<T, true>());
            return red_run(plan, ArgOp<T, false>());
        }
    });
    return DN_OK;
}

dn_status dn_find_last_axis(const void *value, const dn_tensor *t, const dn_tensor *a) {
    if (!value) return set_error(DN_ERR_INVALID_ARG, "FindLastAxis: null value");
    RedPlan plan;
    dn_status st = red_make_plan(plan, t, a, "FindLastAxis");
    if (st != DN_OK) return st;
    if (t->dtype != DN_I64) return set_error(DN_ERR_INVALID_ARG, "FindLastAxis: target must be int64");
    DN_SWITCH_DTYPE(a->dtype, {
        FindOp<T> op;
        memcpy(&op.value, value, sizeof(T));
        return red_run(plan, op);
    });
    return DN_OK;
}

dn_status dn_arg_reduce_combine(int32_t op, const dn_tensor *t, const dn_tensor *vals, const dn_tensor *idxs) {
    if (!tensor_valid(t) || !tensor_valid(vals) || !tensor_valid(idxs))
        return set_error(DN_ERR_INVALID_ARG, "arg-reduce combine: bad argument");
    if (op != DN_ARG_MIN && op != DN_ARG_MAX) return set_error(DN_ERR_INVALID_ARG, "arg-reduce combine: bad op");
    if (t->dtype != DN_I64 || idxs->dtype != DN_I64 || vals->ndims != 2 || idxs->ndims != 2 || t->ndims > 1 ||
        vals->shape[0] != idxs->shape[0] || vals->shape[1] != idxs->shape[1] ||
        (t->ndims == 1 ? t->shape[0] : 1) != vals->shape[1])
        return set_error(DN_ERR_SHAPE_MISMATCH, "arg-reduce combine: expected vals/idxs [nparts, n] and target [n]");
    if (vals->dtype == DN_BOOL) return set_error(DN_ERR_UNSUPPORTED, "arg-reduce combine: bool");
    DN_SWITCH_DTYPE(vals->dtype, {
        if constexpr (!kIsBool<T>) return arg_combine_typed<T>(op, t, vals, idxs);
    });
    return DN_OK;
}

}  // extern "C"
