// linalg.cu — BatchedInvert (Tensor/Tensor/TensorBackend.fs:142; frontend Tensor.fs:2809-2838).
//
// Replaces CudaBackend.fs:451-484: a row-major -> column-major copy of the source, cuBLAS getrfBatched +
// getriBatched through device pointer arrays, two synchronous `info` checks and a copy-back of the result.
// Here one CTA inverts one matrix IN PLACE in the target (after a strided copy of the source into it, exactly
// like the host backend: HostBackend.fs:552-554) by Gauss-Jordan elimination with partial (row) pivoting — the
// same pivot rule as LAPACK's getrf (largest magnitude in the column, first occurrence), so well-conditioned
// results agree with the host's getrf + getri to rounding. Matrices up to 200 KiB live in shared memory for the
// whole elimination (one launch for the whole batch); larger ones are eliminated in global memory with one launch
// pair per pivot step (pivot + row exchange per matrix, then a rank-1 update spread over the whole GPU). Singular
// matrices (an exactly zero pivot, LAPACK info > 0) raise DN_ERR_SINGULAR_MATRIX = SingularMatrixException.
#include "ew_ops.cuh"

using namespace dn;

namespace {

constexpr int kInvThreads = 256;

struct InvParams {
    char *t;             // target, holds a copy of the source on entry
    int64_t row_stride;  // elements
    int64_t col_stride;
    int32_t n;
    int32_t nbatch_dims;
    uint32_t bshape[DN_MAX_DIMS];   // batch dims, innermost-first
    FastDiv bdiv[DN_MAX_DIMS];
    int64_t bstride[DN_MAX_DIMS];   // elements
    int *singular;       // device flag, set to 1 + batch index of a singular matrix
    int use_smem;
};

template <class T>
__device__ __forceinline__ T abs_of(T v) { return v < T(0) ? -v : v; }

// W is an n x n row-major working copy with leading dimension ld.
template <class T>
__device__ void gauss_jordan(T *W, int n, int64_t ld, int *piv, int *singular, int batch) {
    __shared__ int s_p;
    __shared__ T s_pivinv;
    __shared__ T s_best[kInvThreads / 32];
    __shared__ int s_widx[kInvThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = 0; k < n; ++k) {
        // pivot: largest |W[i][k]|, i >= k, lowest i on ties (idamax)
        T best = T(-1);
        int bi = n;
        for (int i = k + tid; i < n; i += kInvThreads) {
            const T v = abs_of(W[i * ld + k]);
            if (v > best) { best = v; bi = i; }  // ascending i per thread: strict > keeps the first
        }
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) {
            const T ob = __shfl_xor_sync(0xffffffffu, best, s);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, s);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (lane == 0) { s_best[warp] = best; s_widx[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
            T b = s_best[0];
            int p = s_widx[0];
            for (int w = 1; w < kInvThreads / 32; ++w)
                if (s_best[w] > b || (s_best[w] == b && s_widx[w] < p)) { b = s_best[w]; p = s_widx[w]; }
            s_p = p;
            piv[k] = p;
            if (!(b > T(0))) {  // zero (or NaN) pivot column: singular
                atomicCAS(singular, 0, batch + 1);
                s_p = -1;
            } else {
                s_pivinv = T(1) / W[p * ld + k];
            }
        }
        __syncthreads();
        const int p = s_p;
        if (p < 0) return;
        if (p != k)
            for (int j = tid; j < n; j += kInvThreads) {
                const T a = W[k * ld + j];
                W[k * ld + j] = W[p * ld + j];
                W[p * ld + j] = a;
            }
        __syncthreads();
        const T pivinv = s_pivinv;
        // scale the pivot row; its k-th entry becomes 1 / pivot
        for (int j = tid; j < n; j += kInvThreads) {
            const T v = W[k * ld + j];
            W[k * ld + j] = j == k ? pivinv : v * pivinv;
        }
        __syncthreads();
        // eliminate column k from every other row: a warp per row, lanes along the row
        for (int i = warp; i < n; i += kInvThreads / 32) {
            if (i == k) continue;
            const T f = W[i * ld + k];
            __syncwarp();
            for (int j = lane; j < n; j += 32) {
                const T r = W[k * ld + j];
                W[i * ld + j] = j == k ? -f * r : W[i * ld + j] - f * r;
            }
        }
        __syncthreads();
    }
    // undo the row exchanges as column exchanges, in reverse order
    for (int k = n - 1; k >= 0; --k) {
        const int p = piv[k];
        if (p != k)
            for (int i = tid; i < n; i += kInvThreads) {
                const T a = W[i * ld + k];
                W[i * ld + k] = W[i * ld + p];
                W[i * ld + p] = a;
            }
        __syncthreads();
    }
}

template <class T>
__global__ void __launch_bounds__(kInvThreads) batched_invert_kernel(const __grid_constant__ InvParams p, int *piv_global) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = p.n;
    // batch offset
    uint32_t rem = blockIdx.x;
    int64_t off = 0;
#pragma unroll
    for (int d = 0; d < DN_MAX_DIMS; ++d) {
        if (d >= p.nbatch_dims) break;
        const uint32_t q = p.bdiv[d].div(rem);
        off += (int64_t)(rem - q * p.bshape[d]) * p.bstride[d];
        rem = q;
    }
    T *M = reinterpret_cast<T *>(p.t) + off;
    int *piv = piv_global + (int64_t)blockIdx.x * n;
    {
        T *W = reinterpret_cast<T *>(smem_raw);
        const int ld = n | 1;  // odd leading dimension: column walks are bank-conflict free
        for (int e = threadIdx.x; e < n * n; e += kInvThreads) {
            const int i = e / n, j = e - i * n;
            W[i * ld + j] = M[(int64_t)i * p.row_stride + (int64_t)j * p.col_stride];
        }
        __syncthreads();
        gauss_jordan<T>(W, n, ld, piv, p.singular, (int)blockIdx.x);
        __syncthreads();
        if (*reinterpret_cast<volatile int *>(p.singular)) return;  // leave the target untouched beyond the copy
        for (int e = threadIdx.x; e < n * n; e += kInvThreads) {
            const int i = e / n, j = e - i * n;
            M[(int64_t)i * p.row_stride + (int64_t)j * p.col_stride] = W[i * ld + j];
        }
    }
}

// ---- matrices too large for shared memory: one launch pair per pivot step, the whole GPU per step -------------
__device__ __forceinline__ int64_t inv_batch_offset(const InvParams &p, uint32_t b) {
    uint32_t rem = b;
    int64_t off = 0;
#pragma unroll
    for (int d = 0; d < DN_MAX_DIMS; ++d) {
        if (d >= p.nbatch_dims) break;
        const uint32_t q = p.bdiv[d].div(rem);
        off += (int64_t)(rem - q * p.bshape[d]) * p.bstride[d];
        rem = q;
    }
    return off;
}

// Step k, part 1 (one CTA per matrix): pivot search in column k, row exchange, scaling of the pivot row, and a copy
// of column k (the elimination factors) into colk so that part 2 has no read-after-write hazard on that column.
template <class T>
__global__ void __launch_bounds__(kInvThreads) invert_pivot_kernel(const __grid_constant__ InvParams p, int *piv_global, T *colk_global, int k) {
    if (*reinterpret_cast<volatile int *>(p.singular)) return;
    __shared__ int s_p;
    __shared__ T s_pivinv;
    __shared__ T s_best[kInvThreads / 32];
    __shared__ int s_widx[kInvThreads / 32];
    const int n = p.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    T *W = reinterpret_cast<T *>(p.t) + inv_batch_offset(p, blockIdx.x);
    const int64_t ld = p.row_stride;
    int *piv = piv_global + (int64_t)blockIdx.x * n;
    T *colk = colk_global + (int64_t)blockIdx.x * n;
    T best = T(-1);
    int bi = n;
    for (int i = k + tid; i < n; i += kInvThreads) {
        const T v = abs_of(W[i * ld + k]);
        if (v > best) { best = v; bi = i; }
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const T ob = __shfl_xor_sync(0xffffffffu, best, s);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, s);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) { s_best[warp] = best; s_widx[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
        T b = s_best[0];
        int q = s_widx[0];
        for (int w = 1; w < kInvThreads / 32; ++w)
            if (s_best[w] > b || (s_best[w] == b && s_widx[w] < q)) { b = s_best[w]; q = s_widx[w]; }
        s_p = q;
        piv[k] = q;
        if (!(b > T(0))) {
            atomicCAS(p.singular, 0, (int)blockIdx.x + 1);
            s_p = -1;
        } else {
            s_pivinv = T(1) / W[q * ld + k];
        }
    }
    __syncthreads();
    const int q = s_p;
    if (q < 0) return;
    const T pivinv = s_pivinv;
    for (int j = tid; j < n; j += kInvThreads) {  // exchange rows k and q, scale the new row k
        const T top = W[q * ld + j];
        if (q != k) W[q * ld + j] = W[k * ld + j];
        W[k * ld + j] = j == k ? pivinv : top * pivinv;
    }
    __syncthreads();
    for (int i = tid; i < n; i += kInvThreads) colk[i] = W[i * ld + k];  // after the exchange; colk[k] is unused
}

// Step k, part 2 (grid: column blocks x row blocks x matrices): W[i][j] -= colk[i] * W[k][j]; W[i][k] = -colk[i] / pivot.
template <class T>
__global__ void __launch_bounds__(kInvThreads) invert_eliminate_kernel(const __grid_constant__ InvParams p, const T *colk_global, int k) {
    if (*reinterpret_cast<volatile int *>(p.singular)) return;
    const int n = p.n;
    T *W = reinterpret_cast<T *>(p.t) + inv_batch_offset(p, blockIdx.z);
    const int64_t ld = p.row_stride;
    const T *colk = colk_global + (int64_t)blockIdx.z * n;
    const int j = blockIdx.x * 64 + (threadIdx.x & 63);
    const int i0 = blockIdx.y * 32 + (threadIdx.x >> 6) * 8;
    if (j >= n) return;
    const T r = W[k * ld + j];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int i = i0 + u;
        if (i < n && i != k) {
            const T f = colk[i];
            W[i * ld + j] = j == k ? -f * r : W[i * ld + j] - f * r;
        }
    }
}

// Undo the row exchanges as column exchanges in reverse order (one CTA per matrix).
template <class T>
__global__ void __launch_bounds__(kInvThreads) invert_unpivot_kernel(const __grid_constant__ InvParams p, const int *piv_global) {
    if (*reinterpret_cast<volatile int *>(p.singular)) return;
    const int n = p.n;
    T *W = reinterpret_cast<T *>(p.t) + inv_batch_offset(p, blockIdx.x);
    const int64_t ld = p.row_stride;
    const int *piv = piv_global + (int64_t)blockIdx.x * n;
    // a thread owns whole rows, so the exchanges of one row need no synchronisation between steps
    for (int i = threadIdx.x; i < n; i += kInvThreads)
        for (int k = n - 1; k >= 0; --k) {
            const int q = piv[k];
            if (q != k) {
                const T a = W[i * ld + k];
                W[i * ld + k] = W[i * ld + q];
                W[i * ld + q] = a;
            }
        }
}

template <class T>
void invert_stepwise(const InvParams &p, int64_t batch, int *piv, T *colk) {
    const int n = p.n;
    const dim3 egrid((unsigned)((n + 63) / 64), (unsigned)((n + 31) / 32), (unsigned)batch);
    for (int k = 0; k < n; ++k) {
        DN_LAUNCH((invert_pivot_kernel<T>), (unsigned)batch, kInvThreads, 0, p, piv, colk, k);
        DN_LAUNCH((invert_eliminate_kernel<T>), egrid, kInvThreads, 0, p, colk, k);
    }
    DN_LAUNCH((invert_unpivot_kernel<T>), (unsigned)batch, kInvThreads, 0, p, piv);
}

}  // namespace

extern "C" dn_status dn_batched_invert(const dn_tensor *t, const dn_tensor *a) {
    if (!tensor_valid(t) || !tensor_valid(a)) return set_error(DN_ERR_INVALID_ARG, "BatchedInvert: bad argument");
    if (t->dtype != a->dtype || (t->dtype != DN_F32 && t->dtype != DN_F64))
        return set_error(DN_ERR_UNSUPPORTED, "BatchedInvert: this operation is only supported for floating point numbers");
    if (t->ndims < 2 || !same_shape(t, a) || t->shape[t->ndims - 1] != t->shape[t->ndims - 2])
        return set_error(DN_ERR_SHAPE_MISMATCH, "BatchedInvert: need tensors of equal shape [..., n, n]");
    const int nd = t->ndims;
    const int64_t n = t->shape[nd - 1];
    int64_t batch = 1;
    for (int d = 0; d < nd - 2; ++d) batch *= t->shape[d];
    if (n == 0 || batch == 0) return DN_OK;
    if (n > 32768 || batch >= ((int64_t)1 << 31)) return set_error(DN_ERR_UNSUPPORTED, "BatchedInvert: matrix or batch too large");
    for (int d = 0; d < nd; ++d)
        if (t->stride[d] == 0 && t->shape[d] > 1) return set_error(DN_ERR_INVALID_ARG, "BatchedInvert: the target must not be a broadcast view");
    const int esize = dtype_size(t->dtype);
    const size_t smem_bytes = (size_t)n * (size_t)(n | 1) * esize;
    const bool use_smem = smem_bytes <= 200 * 1024;
    // inversion is done in place in the target: copy first (HostBackend.fs:552-554). The global-memory path
    // needs a dense row-major matrix; a target that is not gets a dense scratch tensor and a copy back.
    const bool row_major = t->stride[nd - 1] == 1 && t->stride[nd - 2] >= n;
    dn_tensor work = *t;
    void *scratch = nullptr;
    dn_status st;
    if (!use_smem && !row_major) {
        st = scratch_alloc((size_t)batch * n * n * esize, &scratch);
        if (st != DN_OK) return st;
        work.base = scratch;
        work.offset = 0;
        int64_t stride = 1;
        for (int d = nd - 1; d >= 0; --d) {
            work.stride[d] = stride;
            stride *= t->shape[d];
        }
    }
    if (!(work.base == a->base && work.offset == a->offset && memcmp(work.stride, a->stride, sizeof(int64_t) * nd) == 0)) {
        st = dn_copy(&work, a);
        if (st != DN_OK) {
            scratch_free(scratch);
            return st;
        }
    }
    InvParams p;
    p.t = data_ptr(&work);
    p.row_stride = work.stride[nd - 2];
    p.col_stride = work.stride[nd - 1];
    p.n = (int32_t)n;
    p.use_smem = use_smem ? 1 : 0;
    p.nbatch_dims = 0;
    for (int k = 0; k < DN_MAX_DIMS; ++k) {
        p.bshape[k] = 1;
        p.bdiv[k].init(1);
        p.bstride[k] = 0;
    }
    for (int d = nd - 3; d >= 0; --d) {  // innermost batch dim first
        p.bshape[p.nbatch_dims] = (uint32_t)work.shape[d];
        p.bdiv[p.nbatch_dims].init((uint32_t)work.shape[d]);
        p.bstride[p.nbatch_dims] = work.stride[d];
        ++p.nbatch_dims;
    }
    void *aux = nullptr;  // [singular flag | pivots | column copies of the stepwise path]
    const size_t piv_bytes = (sizeof(int) * (size_t)(batch * n + 4) + 15) / 16 * 16;
    st = scratch_alloc(piv_bytes + (use_smem ? 0 : (size_t)batch * n * esize), &aux);
    if (st != DN_OK) {
        scratch_free(scratch);
        return st;
    }
    cudaMemsetAsync(aux, 0, sizeof(int) * 4, current_stream());
    p.singular = reinterpret_cast<int *>(aux);
    int *piv = reinterpret_cast<int *>(aux) + 4;
    const size_t dyn = use_smem ? smem_bytes : 0;
    cudaError_t e = cudaSuccess;
    if (!use_smem) {
        if (batch > 65535) {
            scratch_free(aux);
            scratch_free(scratch);
            return set_error(DN_ERR_UNSUPPORTED, "BatchedInvert: more than 65535 matrices of this size");
        }
        char *colk = reinterpret_cast<char *>(aux) + piv_bytes;
        if (t->dtype == DN_F32) invert_stepwise<float>(p, batch, piv, reinterpret_cast<float *>(colk));
        else invert_stepwise<double>(p, batch, piv, reinterpret_cast<double *>(colk));
    } else if (t->dtype == DN_F32) {
        if (dyn > 48 * 1024) e = cudaFuncSetAttribute(batched_invert_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        if (e == cudaSuccess) DN_LAUNCH((batched_invert_kernel<float>), (unsigned)batch, kInvThreads, dyn, p, piv);
    } else {
        if (dyn > 48 * 1024) e = cudaFuncSetAttribute(batched_invert_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        if (e == cudaSuccess) DN_LAUNCH((batched_invert_kernel<double>), (unsigned)batch, kInvThreads, dyn, p, piv);
    }
    st = e == cudaSuccess ? launch_status("BatchedInvert kernel") : cuda_error(e, "BatchedInvert");
    // CheckBlasInfo (CudaBackend.fs:467-476): the reference synchronises to learn whether a matrix was singular
    int flag = 0;
    if (st == DN_OK) {
        e = cudaMemcpyAsync(&flag, p.singular, sizeof(int), cudaMemcpyDeviceToHost, current_stream());
        if (e == cudaSuccess) e = cudaStreamSynchronize(current_stream());
        if (e != cudaSuccess) st = cuda_error(e, "BatchedInvert");
    }
    if (st == DN_OK && flag == 0 && scratch) st = dn_copy(t, &work);
    scratch_free(aux);
    scratch_free(scratch);
    if (st != DN_OK) return st;
    if (flag != 0) return set_error(DN_ERR_SINGULAR_MATRIX, "cannot invert singular matrix (batch element %d)", flag - 1);
    return DN_OK;
}
