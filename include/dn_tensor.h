/*
 * dn_tensor.h — C ABI of libdeepnet_b200.so, the B200-native CUDA device backend for Deep.Net's Tensor<'T>.
 *
 * This is the drop-in boundary: every entry point below is what an F# `TensorCudaBackend<'T>` binds through
 * P/Invoke instead of the reference's NVRTC-compiled kernel tables. Citations are relative to the reference tree
 * (DeepMLNet/DeepNet), in the form file:line.
 *
 *   ITensorBackend<'T> (the 69-member operator interface) ....... Tensor/Tensor/TensorBackend.fs:64-146
 *   TensorCudaBackend<'T> (what this library replaces) .......... Tensor/Tensor/Cuda/CudaBackend.fs:120-492
 *   kernel tables / NVRTC module loader (replaced) .............. Tensor/Tensor/Cuda/CudaKernels.fs:22-406,
 *                                                                 Tensor/Tensor/Cuda/KernelCompiler.fs:94-272
 *   by-value tensor argument struct (replaced by dn_tensor) ..... Tensor/Tensor/Cuda/NativeTensor.fs:50-57,83-88,
 *                                                                 Tensor/Tensor/Cuda/Kernels/Tensor.cuh:7-15
 *   thread-local Cfg.Stream / Cfg.Stacktrace ..................... Tensor/Tensor/Cuda/CudaCfg.fs:14-60
 *
 * Conventions (identical to the reference's backend contract):
 *   - every operator is target-first, the target is pre-allocated, sources are already broadcast to the target
 *     shape (stride 0), all operands live on the current device, operands may alias;
 *   - offsets and strides are in ELEMENTS, strides may be 0 or negative, rank 0 is legal, zero-sized dims are legal;
 *   - bool is one byte (0 / 1), as marshalled by the reference (KernelCompiler.fs:190-193);
 *   - operators are asynchronous on the calling thread's stream (dn_set_stream); nothing synchronises unless stated;
 *   - plain C types only: no CUDA, torch or C++ types cross this boundary.
 */
#ifndef DN_TENSOR_H
#define DN_TENSOR_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DN_MAX_DIMS 8

/* Element types. Order is ABI. (NativeTensor.fs:22-34 maps the same set minus the 16-bit integers.) */
typedef enum dn_dtype {
    DN_F32 = 0, DN_F64 = 1, DN_I8 = 2, DN_U8 = 3, DN_I16 = 4, DN_U16 = 5,
    DN_I32 = 6, DN_U32 = 7, DN_I64 = 8, DN_U64 = 9, DN_BOOL = 10, DN_DTYPE_COUNT = 11
} dn_dtype;

/* Status codes and the .NET exception the F# binding raises for each (SURVEY.md §8b "Error conventions"). */
typedef enum dn_status {
    DN_OK = 0,
    DN_ERR_INVALID_ARG = 1,        /* ArgumentException / InvalidOperationException                         */
    DN_ERR_UNSUPPORTED = 2,        /* NotSupportedException (CudaBackend.fs:126-129, CudaKernels.fs:38)     */
    DN_ERR_OUT_OF_MEMORY = 3,      /* OutOfCudaMemoryException (CudaUtils.fs:183-210)                       */
    DN_ERR_INDEX_OUT_OF_RANGE = 4, /* IndexOutOfRangeException "invalid index during gather or scatter"     */
    DN_ERR_CUDA = 5,               /* CudaException                                                         */
    DN_ERR_NO_DEVICE = 6,          /* CudaException "Cannot create CUDA context" (CudaBackend.fs:28-38)     */
    DN_ERR_SHAPE_MISMATCH = 7,     /* InvalidOperationException                                             */
    DN_ERR_SINGULAR_MATRIX = 8     /* SingularMatrixException "cannot invert singular matrix"               */
} dn_status;

/* Tensor view descriptor: replaces NativeTensor {Ptr; Offset; Shape; Stride} (NativeTensor.fs:50-57). */
typedef struct dn_tensor {
    void   *base;                  /* device pointer to element 0 of the storage                            */
    int64_t offset;                /* in elements                                                           */
    int32_t ndims;                 /* 0..DN_MAX_DIMS                                                        */
    int32_t dtype;                 /* dn_dtype                                                              */
    int64_t shape[DN_MAX_DIMS];
    int64_t stride[DN_MAX_DIMS];   /* in elements; 0 = broadcast, negative = reversed                       */
} dn_tensor;

/* ITensorBackend unary members, TensorBackend.fs:74-94,113. */
typedef enum dn_unary_op {
    DN_UNARY_PLUS = 0, DN_UNARY_MINUS, DN_ABS, DN_SGN, DN_LOG, DN_LOG10, DN_EXP, DN_SIN, DN_COS, DN_TAN,
    DN_ASIN, DN_ACOS, DN_ATAN, DN_SINH, DN_COSH, DN_TANH, DN_SQRT, DN_CEILING, DN_FLOOR, DN_ROUND,
    DN_TRUNCATE, DN_NEGATE /* bool only */, DN_UNARY_OP_COUNT
} dn_unary_op;

/* ITensorBackend binary members, TensorBackend.fs:97-104,114-116. */
typedef enum dn_binary_op {
    DN_ADD = 0, DN_SUBTRACT, DN_MULTIPLY, DN_DIVIDE, DN_MODULO, DN_POWER, DN_MAX_ELEMWISE, DN_MIN_ELEMWISE,
    DN_AND, DN_OR, DN_XOR, DN_BINARY_OP_COUNT
} dn_binary_op;

/* ITensorBackend comparison members, TensorBackend.fs:106-111. */
typedef enum dn_compare_op {
    DN_EQUAL = 0, DN_NOT_EQUAL, DN_LESS, DN_LESS_OR_EQUAL, DN_GREATER, DN_GREATER_OR_EQUAL, DN_COMPARE_OP_COUNT
} dn_compare_op;

/* ITensorBackend *LastAxis folds, TensorBackend.fs:125-131. */
typedef enum dn_reduce_op {
    DN_SUM = 0, DN_PRODUCT, DN_MIN, DN_MAX, DN_ALL, DN_ANY, DN_COUNT_TRUE, DN_REDUCE_OP_COUNT
} dn_reduce_op;

/* ITensorBackend Arg*LastAxis, TensorBackend.fs:133-134. */
typedef enum dn_arg_reduce_op { DN_ARG_MIN = 0, DN_ARG_MAX = 1 } dn_arg_reduce_op;

/* SpecialIdx.NotFound, Tensor/Tensor/TensorRng.fs:24 — result of Arg*LastAxis / FindLastAxis when nothing matches. */
#define DN_NOT_FOUND (INT64_MIN + 4)

/* ---------------------------------------------------------------------------------------------------------------
 * Device, stream and error state.  Replaces module Cuda (Tensor/Tensor/Cuda/CudaUtils.fs:19-210) and Cfg.
 * ------------------------------------------------------------------------------------------------------------- */

/* CudaInit.check (CudaBackend.fs:28-38): binds the calling thread to `device`, creates the primary context, the
 * memory pool and the per-device scratch. DN_ERR_NO_DEVICE if there is no usable GPU. Idempotent. */
dn_status dn_init(int32_t device);
dn_status dn_device_count(int32_t *count);
dn_status dn_set_device(int32_t device);
dn_status dn_get_device(int32_t *device);
/* Cfg.Stream (CudaCfg.fs:25-27): thread-local; `stream` is a cudaStream_t / CUstream passed as an opaque pointer,
 * NULL = the default stream. */
dn_status dn_set_stream(void *stream);
dn_status dn_get_stream(void **stream);
/* cuCtxSynchronize on the calling thread's stream (Benchmark.fs:211-213 calls this after every op). */
dn_status dn_sync(void);
/* Cfg.Stacktrace (CudaCfg.fs:33-35): when non-zero, gather/scatter synchronise and return
 * DN_ERR_INDEX_OUT_OF_RANGE on a bad index (CudaKernels.fs:335-342). When zero, a bad index sets a sticky
 * per-device flag that dn_poll_index_error reads (the reference executes `trap` and loses the context). */
dn_status dn_set_check_errors(int32_t enabled);
dn_status dn_poll_index_error(int32_t *had_error);
/* Message of the last non-OK status returned on this thread (never NULL). */
const char *dn_last_error(void);
/* Number of device kernels this library has launched in this process (bench.py reports it as gpu_launches). */
int64_t dn_launch_count(void);
const char *dn_version(void);

/* ---------------------------------------------------------------------------------------------------------------
 * Storage.  Replaces TensorCudaStorage<'T> (CudaBackend.fs:51-108), Cuda.newDevVar (CudaUtils.fs:183-210) and the
 * event-per-operand keep-alive (CudaUtils.fs:122-177): allocation and free are stream-ordered on the calling
 * thread's stream, so a storage may be freed right after the last op that uses it was enqueued.
 * ------------------------------------------------------------------------------------------------------------- */
dn_status dn_alloc(int64_t nbytes, void **ptr);          /* nbytes <= 0 allocates 1 byte (CudaBackend.fs:56-58) */
dn_status dn_free(void *ptr);
dn_status dn_alloc_host(int64_t nbytes, void **ptr);     /* pinned host memory (CudaRegMem.fs:123-146)          */
dn_status dn_free_host(void *ptr);
dn_status dn_memset_zero(void *ptr, int64_t nbytes);
/* Transfer (CudaBackend.fs:206-270) for C-contiguous blocks; async on the stream when the host side is pinned. */
dn_status dn_memcpy_h2d(void *dst_dev, const void *src_host, int64_t nbytes);
dn_status dn_memcpy_d2h(void *dst_host, const void *src_dev, int64_t nbytes);
dn_status dn_memcpy_d2d(void *dst_dev, const void *src_dev, int64_t nbytes);
/* Same as dn_memcpy_d2h but does not wait: the host buffer (pinned) is valid after dn_sync() on the same stream.
 * This is the `Cfg.Stream <> NullStream` branch of Transfer (CudaBackend.fs:243-247, AsyncCopyFromDevice). */
dn_status dn_memcpy_d2h_async(void *dst_host, const void *src_dev, int64_t nbytes);
/* Stream ordering helpers for callers that overlap transfers with compute on several streams: record an event on
 * the calling thread's stream / make the calling thread's stream wait for it. Events are created by dn_event_create
 * and are opaque. */
dn_status dn_event_create(void **event);
dn_status dn_event_destroy(void *event);
dn_status dn_event_record(void *event);
dn_status dn_stream_wait_event(void *event);
/* ITensorBackend.Item get/set (CudaBackend.fs:77-93,199-201): synchronous single-element access; `pos` has
 * t->ndims entries; `value` points to one element of t->dtype. */
dn_status dn_get_item(const dn_tensor *t, const int64_t *pos, void *value);
dn_status dn_set_item(const dn_tensor *t, const int64_t *pos, const void *value);

/* ---------------------------------------------------------------------------------------------------------------
 * Element-wise operators.  Target and sources have identical shapes.
 * ------------------------------------------------------------------------------------------------------------- */
/* FillConst (TensorBackend.fs:71; CudaBackend.fs:272-275; Elemwise.cuh:13-27). `value`: one host element of t->dtype. */
dn_status dn_fill_const(const dn_tensor *t, const void *value);
/* FillIncrementing (TensorBackend.fs:72; Elemwise.cuh:29-43; host ScalarOps.fs:367-370): t[p] = start + incr*p[0]. */
dn_status dn_fill_incrementing(const dn_tensor *t, const void *start, const void *incr);
/* Copy (TensorBackend.fs:67; CudaBackend.fs:282-298). Same dtype. */
dn_status dn_copy(const dn_tensor *t, const dn_tensor *a);
/* Convert (TensorBackend.fs:69; CudaBackend.fs:300-302; Elemwise.cuh:50-63): static_cast between any two dtypes. */
dn_status dn_convert(const dn_tensor *t, const dn_tensor *a);
/* UnaryPlus..Truncate, Negate (TensorBackend.fs:74-94,113; CudaBackend.fs:305-325,346). */
dn_status dn_unary(int32_t op, const dn_tensor *t, const dn_tensor *a);
/* Add..MinElemwise, And/Or/Xor (TensorBackend.fs:97-104,114-116; CudaBackend.fs:327-334,347-349). */
dn_status dn_binary(int32_t op, const dn_tensor *t, const dn_tensor *a, const dn_tensor *b);
/* Equal..GreaterOrEqual (TensorBackend.fs:106-111; CudaBackend.fs:336-341): t is DN_BOOL. */
dn_status dn_compare(int32_t op, const dn_tensor *t, const dn_tensor *a, const dn_tensor *b);
/* IsFinite (TensorBackend.fs:95; CudaBackend.fs:342): t is DN_BOOL. */
dn_status dn_is_finite(const dn_tensor *t, const dn_tensor *a);
/* IfThenElse (TensorBackend.fs:118; CudaBackend.fs:344): cond is DN_BOOL. */
dn_status dn_if_then_else(const dn_tensor *t, const dn_tensor *cond, const dn_tensor *if_true,
                          const dn_tensor *if_false);

/* Fused element-wise expression (new; SURVEY.md §8f-3 — the reference reaches fusion only through the Symbolic
 * layer's generated "elements" kernels, Examples/LearnMnist/Program.fs:14). One pass over the operands evaluates a
 * short straight-line program over six virtual registers: the sources are preloaded into registers 0..nsrc-1, every
 * instruction computes r[dst] = op(r[a] [, r[b]]) or r[dst] = imm, the target receives r[dst of the last
 * instruction]. Each instruction rounds to the element type exactly like the corresponding single operator
 * (no fused multiply-add), so the result is bit-identical to the sequence of dn_unary / dn_binary calls it replaces
 * while moving (nsrc + 1) instead of ~3 tensors per operator through HBM. f32 / f64; all operands of one type.
 * Example, c = a*b + sin(a):  {BINARY MULTIPLY 2 <- 0,1} {UNARY SIN 3 <- 0} {BINARY ADD 2 <- 2,3}. */
typedef enum dn_fused_kind { DN_FUSED_UNARY = 0, DN_FUSED_BINARY = 1, DN_FUSED_CONST = 2 } dn_fused_kind;
typedef struct dn_fused_instr {
    int32_t kind;   /* dn_fused_kind                                                            */
    int32_t op;     /* dn_unary_op (UNARY) / dn_binary_op Add..MinElemwise (BINARY) / unused    */
    int32_t dst;    /* 0..DN_FUSED_REGS-1                                                       */
    int32_t a, b;   /* operand registers; b is ignored by UNARY, both by CONST                  */
    double  imm;    /* CONST: the value, converted to the element type                          */
} dn_fused_instr;
#define DN_FUSED_REGS 6
#define DN_FUSED_MAX_INSTRS 12
#define DN_FUSED_MAX_SRCS 3
dn_status dn_fused_elemwise(const dn_tensor *t, const dn_tensor *const *srcs, int32_t nsrc,
                            const dn_fused_instr *prog, int32_t ninstr);

/* ---------------------------------------------------------------------------------------------------------------
 * Last-axis reductions.  a has shape [..., L], t has shape [...].
 * ------------------------------------------------------------------------------------------------------------- */
/* Sum/Product/Min/Max/All/Any/CountTrue LastAxis (TensorBackend.fs:125-131; CudaBackend.fs:351-356,489).
 * CountTrue: a is DN_BOOL and t is DN_I64; All/Any: both DN_BOOL; others: same dtype. */
dn_status dn_reduce_last_axis(int32_t op, const dn_tensor *t, const dn_tensor *a);
/* ArgMin/ArgMaxLastAxis (TensorBackend.fs:133-134; CudaBackend.fs:358-359): t is DN_I64; host semantics
 * (ScalarOps.fs:638-654): first strict extremum, DN_NOT_FOUND if nothing beats the initial value. */
dn_status dn_arg_reduce_last_axis(int32_t op, const dn_tensor *t, const dn_tensor *a);
/* FindLastAxis (TensorBackend.fs:135; CudaBackend.fs:360): first index with a == *value, else DN_NOT_FOUND. */
dn_status dn_find_last_axis(const void *value, const dn_tensor *t, const dn_tensor *a);

/* ---------------------------------------------------------------------------------------------------------------
 * Indexing.
 * ------------------------------------------------------------------------------------------------------------- */
/* Gather (TensorBackend.fs:119; CudaBackend.fs:362-370; GatherScatter.cuh:26-68): idxs has a->ndims entries,
 * NULL entry = `None` (identity on that dimension); non-NULL entries are DN_I64 with the shape of t. */
dn_status dn_gather(const dn_tensor *t, const dn_tensor *const *idxs, int32_t nidxs, const dn_tensor *a);
/* Scatter (TensorBackend.fs:120; CudaBackend.fs:372-381; GatherScatter.cuh:72-114): zero-fills t, then
 * t[idx(p)] += a[p]; idxs has t->ndims entries with the shape of a. */
dn_status dn_scatter(const dn_tensor *t, const dn_tensor *const *idxs, int32_t nidxs, const dn_tensor *a);
/* countTrue ∘ flatten with the synchronous read-back the frontend performs before it allocates the target of
 * MaskedGet / TrueIndices (Tensor.fs:2259-2262,3025). Blocking. */
dn_status dn_count_true(const dn_tensor *a, int64_t *count);
/* MaskedGet (TensorBackend.fs:121; host ScalarOps.fs:667-681): masks has a->ndims entries, NULL = NoMask; each
 * non-NULL mask is a 1-D DN_BOOL tensor of length a->shape[d]. t has a->ndims dims, t->shape[d] = countTrue(mask d). */
dn_status dn_masked_get(const dn_tensor *t, const dn_tensor *a, const dn_tensor *const *masks, int32_t nmasks);
/* MaskedSet (TensorBackend.fs:122; host ScalarOps.fs:683-697): mirror image; a may be broadcast (stride 0). */
dn_status dn_masked_set(const dn_tensor *t, const dn_tensor *const *masks, int32_t nmasks, const dn_tensor *a);
/* TrueIndices (TensorBackend.fs:123; host ScalarOps.fs:699-707): t is DN_I64 [nTrue, a->ndims]. */
dn_status dn_true_indices(const dn_tensor *t, const dn_tensor *a);

/* ---------------------------------------------------------------------------------------------------------------
 * Dense contractions.  f32 / f64.
 * ------------------------------------------------------------------------------------------------------------- */
/* VecVecDot / MatVecDot (TensorBackend.fs:137-138; CudaBackend.fs:383-408). */
dn_status dn_vec_vec_dot(const dn_tensor *t, const dn_tensor *a, const dn_tensor *b);
dn_status dn_mat_vec_dot(const dn_tensor *t, const dn_tensor *a, const dn_tensor *b);
/* MatMatDot / BatchedMatMatDot (TensorBackend.fs:139-140; CudaBackend.fs:410-449): t[..,M,N] = a[..,M,K]·b[..,K,N]. */
dn_status dn_mat_mat_dot(const dn_tensor *t, const dn_tensor *a, const dn_tensor *b);
dn_status dn_batched_mat_mat_dot(const dn_tensor *t, const dn_tensor *a, const dn_tensor *b);

/* BatchedInvert (TensorBackend.fs:142; CudaBackend.fs:451-484; host HostBackend.fs:548-577 = LAPACK getrf + getri):
 * t[..., n, n] = inverse of a[..., n, n], f32 / f64, any strides; t and a may be the same view. Partial pivoting.
 * Blocking (the reference synchronises on cuBLAS' info array, CudaBackend.fs:467-476): returns
 * DN_ERR_SINGULAR_MATRIX when a pivot is exactly zero. */
dn_status dn_batched_invert(const dn_tensor *t, const dn_tensor *a);

/* ---------------------------------------------------------------------------------------------------------------
 * Multi-GPU combine steps for leading-axis sharding (new; the reference has no multi-GPU path — SURVEY.md §8e).
 * The bulk tensors never move: every rank reduces its slab with the operators above; these entry points only
 * fold the per-rank partial results that the host has gathered (torch.distributed / NCCL all_gather).
 * ------------------------------------------------------------------------------------------------------------- */
/* Fold `nparts` partial (value, index) pairs per output into t (DN_I64): best value wins, lowest global index on
 * ties, DN_NOT_FOUND partials never win. vals: [nparts, n] of a's dtype, idxs: [nparts, n] DN_I64 (global indices). */
dn_status dn_arg_reduce_combine(int32_t op, const dn_tensor *t, const dn_tensor *vals, const dn_tensor *idxs);

#ifdef __cplusplus
}
#endif
#endif /* DN_TENSOR_H */
