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
/* Tells the library that `stream` is about to be destroyed: it is dropped from the device's list of streams that
 * storage releases are ordered after (dn_free / dn_free_deferred). */
dn_status dn_release_stream(void *stream);
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
/* Stream-ordered on the calling thread's stream; if the library knows other streams on the device (dn_set_stream),
 * the release is additionally ordered after everything they hold. */
dn_status dn_free(void *ptr);
/* Release from a thread that owns no stream on the storage's device — a .NET finalizer thread (the reference frees
 * from the finalizer, CudaBackend.fs:73-74), a destructor run by a collector thread. Any device, any thread, never
 * touches a stream: the pointer is queued on its device and released by the next dn_alloc / dn_sync issued by a
 * thread working on that device, behind a fence over all streams the library knows there. */
dn_status dn_free_deferred(void *ptr);
dn_status dn_alloc_host(int64_t nbytes, void **ptr);     /* pinned host memory                                   */
dn_status dn_free_host(void *ptr);
/* CudaRegMem.register / unregister (CudaRegMem.fs:114-146): page-locks EXISTING host memory (a pinned managed
 * array) so that transfers from / to it are asynchronous DMA. DN_ERR_INVALID_ARG when the range cannot be registered
 * (the reference's CannotCudaRegisterMemoryException: the caller falls back to plain pinning). */
dn_status dn_host_register(void *ptr, int64_t nbytes);
dn_status dn_host_unregister(void *ptr);
dn_status dn_memset_zero(void *ptr, int64_t nbytes);
/* Transfer (CudaBackend.fs:206-270) for C-contiguous blocks; async on the stream when the host side is pinned. */
dn_status dn_memcpy_h2d(void *dst_dev, const void *src_host, int64_t nbytes);
dn_status dn_memcpy_d2h(void *dst_host, const void *src_dev, int64_t nbytes);
dn_status dn_memcpy_d2d(void *dst_dev, const void *src_dev, int64_t nbytes);
/* Same as dn_memcpy_d2h but does not wait: the host buffer (pinned) is valid after dn_sync() on the same stream.
 * This is the `Cfg.Stream <> NullStream` branch of Transfer (CudaBackend.fs:243-247, AsyncCopyFromDevice). */
dn_status dn_memcpy_d2h_async(void *dst_host, const void *src_dev, int64_t nbytes);
/* ITensorBackend.Transfer (TensorBackend.fs:68; CudaBackend.fs:206-270) for ARBITRARY views of equal shape and type:
 * host_t->base is a host pointer, dev_t->base a device pointer. C-contiguous pairs are one DMA; any other layout is
 * handled natively (strided pack / unpack of the host view through pinned staging, layout change on the device with
 * the strided copy kernel), where the reference copies on the host and through a temporary device tensor.
 * h2d is ordered on the calling thread's stream; d2h blocks until the host view holds the data. */
dn_status dn_transfer_h2d(const dn_tensor *dev_t, const dn_tensor *host_t);
dn_status dn_transfer_d2h(const dn_tensor *host_t, const dn_tensor *dev_t);
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
/* Precision of float32 MatMatDot / BatchedMatMatDot (process-wide). The reference calls cuBLAS SGEMM (full fp32;
 * its test "Single matrix dot", Tensor.Test/CudaTests.fs:52-62, compares with the host at rel 1e-5).
 *   DN_MATH_FP32 (default): problems below 2^27 multiply-adds (every product the reference's tests form) run on an
 *                           exact fp32 kernel; larger ones as 3xTF32 on the tensor cores (operands split into two
 *                           tf32 halves, three tcgen05 MMAs per k-step): input error 2^-21 instead of tf32's 2^-11;
 *                           the tensor core's accumulator truncates after every MMA, which leaves a norm-wise
 *                           relative error of ~7e-6 at K = 1024, growing linearly with K;
 *   DN_MATH_TF32          : one tcgen05 pass with tf32 inputs (10-bit mantissa), fp32 accumulation; rel 1e-2 of
 *                           the fp64 result (BASELINE.json north_star), three times the throughput. Opt-in;
 *   DN_MATH_FP32_STRICT   : the exact fp32 kernel (CUDA cores, fused multiply-add, round to nearest) for every size. */
typedef enum dn_math_mode { DN_MATH_FP32 = 0, DN_MATH_TF32 = 1, DN_MATH_FP32_STRICT = 2 } dn_math_mode;
dn_status dn_set_math_mode(int32_t mode);
dn_status dn_get_math_mode(int32_t *mode);
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
 * Multi-GPU: leading-axis sharding over the GPUs of one box (new; the reference is single device —
 * Tensor/Tensor/Cuda/CudaUtils.fs:42-46 creates one context; SURVEY.md §8e defines the partitioning).
 *
 * Every rank owns a contiguous slab of dim 0 of every operand; element-wise operators need no communication and
 * are the ordinary entry points above on each device. A *shard group* connects the ranks for the operators whose
 * result every rank needs in full (reductions, arg-reductions, Find, TrueIndices, MaskedGet):
 *   - every rank owns a WINDOW of device memory that all other ranks can address (NVLink / NVSwitch peer access
 *     inside one process, CUDA IPC mappings between processes); full results live in its symmetric heap
 *     (dn_shard_heap_alloc: the same sequence of allocations on every rank yields the same window offsets);
 *   - a sharded operator is ONE kernel launch per rank: the reduction kernel stores each output row into the local
 *     result AND the same offset of every peer's result, and its last CTA publishes the rank's epoch into every
 *     rank's flag array; the launch is followed by stream memory operations (cuStreamWaitValue32) that hold the
 *     rank's stream until every rank has published, so later work on the stream sees the complete replicated
 *     result. No NCCL launch is involved, no kernel spins, the bulk tensors never move.
 * Process models: (a) one process driving all devices (the F# host: nlocal == world), (b) one process per GPU
 * (torchrun: nlocal == 1; the 64-byte window handles are exchanged by the host through any side channel).
 * Ranks may share a device (tests on a single GPU). Calls on one rank must come from one thread at a time and are
 * COLLECTIVE: every rank issues the same sequence of dn_shard_* operator calls. A result buffer may be reused as
 * the target of a later collective only after at least one other collective (or dn_shard_barrier) in between; the
 * library inserts that barrier itself when the same target comes twice in a row (outside group brackets).
 * All dn_shard_* calls run on the rank's stream (dn_shard_set_stream; default: a stream owned by the group) and
 * leave the calling thread's current device and stream unchanged.
 * ------------------------------------------------------------------------------------------------------------- */
#define DN_SHARD_MAX_RANKS 8
#define DN_SHARD_HANDLE_BYTES 64
/* Creates the group object and the windows of the `nlocal` ranks this process drives (local_ranks[i] runs on CUDA
 * device local_devices[i]); heap_bytes = size of each rank's symmetric heap. */
dn_status dn_shard_group_create(int32_t world, int32_t nlocal, const int32_t *local_ranks,
                                const int32_t *local_devices, int64_t heap_bytes, void **group);
/* The window handle of a LOCAL rank (DN_SHARD_HANDLE_BYTES bytes), to be sent to the other processes. */
dn_status dn_shard_group_handle(void *group, int32_t rank, void *handle);
/* Maps the windows of the non-local ranks (handles: world x DN_SHARD_HANDLE_BYTES, entries of local ranks are
 * ignored; may be NULL when every rank is local) and enables peer access between the devices. */
dn_status dn_shard_group_connect(void *group, const void *handles);
dn_status dn_shard_group_destroy(void *group);
dn_status dn_shard_set_stream(void *group, int32_t rank, void *stream);
dn_status dn_shard_sync(void *group, int32_t rank);          /* waits for the rank's stream; reports a barrier that timed out */
/* Contiguous slab [begin, begin+count) of `nrows` rows owned by `rank`; the remainder goes to the first ranks. */
dn_status dn_shard_slab(int64_t nrows, int32_t rank, int32_t world, int64_t *begin, int64_t *count);
/* Symmetric heap: a bump allocator (256-byte granules); reset frees everything. Collective by convention. */
dn_status dn_shard_heap_alloc(void *group, int32_t rank, int64_t nbytes, void **ptr);
dn_status dn_shard_heap_reset(void *group, int32_t rank);
/* Flag barrier over peer memory on the ranks' streams (one tiny kernel per rank). */
dn_status dn_shard_barrier(void *group, int32_t rank);
/* Brackets (as ncclGroupStart / ncclGroupEnd). Inside a bracket a collective launches its kernel (stores + signal)
 * at once and DEFERS the wait half of the barrier — and whatever needs every rank's contribution — to
 * dn_shard_group_end, which enqueues ONE wait per rank for the rank's last collective (flags only grow).
 *   - ONE thread driving several ranks MUST bracket:  start;  for each local rank r: dn_shard_<op>(g, r, ...);  end.
 *     A stream already waiting for a peer whose kernel the same thread has yet to launch would turn any blocking call
 *     on the way there into a deadlock.
 *   - Any caller MAY bracket several collectives per rank (e.g. ArgMax and Max of the same logits): their waits
 *     coalesce into one, and the kernels run back to back. Results are ordered on the rank's stream after
 *     dn_shard_group_end; targets inside one bracket must be distinct; a reduction over the sharded axis or a count
 *     exchange (they need their wait first) must be the rank's last collective of the bracket.
 * Outside brackets every call completes its own wait. */
dn_status dn_shard_group_start(void *group);
dn_status dn_shard_group_end(void *group);

/* Reductions over an axis OTHER than the sharded one (every output row lives on one rank). a_local: this rank's
 * slab [rows_local, ..., L]; t_full: the FULL result [rows_total, ...] in the symmetric heap; the rank computes rows
 * [row_begin, row_begin + rows_local) and stores them into every rank's t_full. Semantics of dn_reduce_last_axis /
 * dn_arg_reduce_last_axis / dn_find_last_axis. */
dn_status dn_shard_reduce_last_axis(void *group, int32_t rank, int32_t op, const dn_tensor *t_full,
                                    int64_t row_begin, const dn_tensor *a_local);
dn_status dn_shard_arg_reduce_last_axis(void *group, int32_t rank, int32_t op, const dn_tensor *t_full,
                                        int64_t row_begin, const dn_tensor *a_local);
dn_status dn_shard_find_last_axis(void *group, int32_t rank, const void *value, const dn_tensor *t_full,
                                  int64_t row_begin, const dn_tensor *a_local);
/* Min/Max AND ArgMin/ArgMax of the same source in ONE pass over it (op: dn_arg_reduce_op; float32 / float64):
 * t_val_full receives exactly what MinLastAxis / MaxLastAxis would, t_idx_full what ArgMin / ArgMaxLastAxis would.
 * group == NULL: plain single-device call (row_begin must be 0). */
dn_status dn_shard_minmax_arg_last_axis(void *group, int32_t rank, int32_t op, const dn_tensor *t_val_full,
                                        const dn_tensor *t_idx_full, int64_t row_begin, const dn_tensor *a_local);
/* Reductions over the SHARDED axis (includes whole-tensor folds of a flattened slab). a_local: [..., n_local] with
 * the sharded axis last (the frontend's axis->last permutation), holding global positions [axis_begin,
 * axis_begin + n_local); t: [...] anywhere in device memory, receives the full result on every rank. Per-rank
 * partials are stored into every rank's window and folded locally IN RANK ORDER, so the result is identical on
 * every rank: float Min/Max keep the host's order-dependent NaN rule, ArgMin/ArgMax first-occurrence semantics
 * through (value, global index) pairs, Find the lowest global index. kind: 0 = dn_reduce_op, 1 = dn_arg_reduce_op,
 * 2 = Find (`value` points to one element of a's dtype). */
dn_status dn_shard_reduce_sharded_axis(void *group, int32_t rank, int32_t kind, int32_t op, const void *value,
                                       const dn_tensor *t, int64_t axis_begin, const dn_tensor *a_local);
/* All-gather of row blocks already present in the local t_full (symmetric heap): rows [row_begin, row_begin+nrows)
 * of the C-contiguous t_full are stored into every rank's t_full. nrows may differ between ranks (ragged). */
dn_status dn_shard_all_gather_rows(void *group, int32_t rank, const dn_tensor *t_full, int64_t row_begin,
                                   int64_t nrows);
/* countTrue of the local mask slab, exchanged: counts[r] = number of true elements on rank r. Blocking (the
 * frontend needs the total before it can allocate the result, Tensor.fs:2259-2262,3025). A single thread that drives
 * several ranks issues _begin on every rank first, then _end on every rank. */
dn_status dn_shard_count_true(void *group, int32_t rank, const dn_tensor *mask_local, int64_t *counts);
dn_status dn_shard_count_true_begin(void *group, int32_t rank, const dn_tensor *mask_local);
dn_status dn_shard_count_true_end(void *group, int32_t rank, int64_t *counts);
/* TrueIndices of a bool tensor sharded along dim 0: local compaction into rows [row_offset, row_offset + nrows_local)
 * of t_full [nTrue_total, ndims] with dim-0 coordinates shifted by dim0_begin, blocks replicated in rank order
 * (= the logical row-major order of the full tensor, ScalarOps.fs:699-707). */
dn_status dn_shard_true_indices(void *group, int32_t rank, const dn_tensor *t_full, int64_t row_offset,
                                int64_t nrows_local, const dn_tensor *mask_local, int64_t dim0_begin);
/* MaskedGet with a full-shape mask, both sharded along dim 0 (ScalarOps.fs:667-681): t_full is 1-D [nTrue_total]. */
dn_status dn_shard_masked_get(void *group, int32_t rank, const dn_tensor *t_full, int64_t elem_offset,
                              int64_t nelems_local, const dn_tensor *a_local, const dn_tensor *mask_local);

/* Fold `nparts` partial (value, index) pairs per output into t (DN_I64): best value wins, lowest global index on
 * ties, DN_NOT_FOUND partials never win. vals: [nparts, n] of a's dtype, idxs: [nparts, n] DN_I64 (global indices). */
dn_status dn_arg_reduce_combine(int32_t op, const dn_tensor *t, const dn_tensor *vals, const dn_tensor *idxs);

#ifdef __cplusplus
}
#endif
#endif /* DN_TENSOR_H */
