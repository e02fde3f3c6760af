// common.cuh — shared host/device plumbing of libdeepnet_b200.so.
//
// Replaces the launch glue of the reference's CUDA backend: module Cuda (Tensor/Tensor/Cuda/CudaUtils.fs:19-210),
// the kernel argument marshalling (Tensor/Tensor/Cuda/NativeTensor.fs:50-208) and the per-(dtype, rank) NVRTC
// module cache (Tensor/Tensor/Cuda/KernelCompiler.fs:94-272). Here every kernel is precompiled for sm_100a and
// takes a rank-agnostic parameter block, so there is nothing to compile or cache at run time.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <type_traits>

#include "../../include/dn_tensor.h"

namespace dn {

// ------------------------------------------------------------------------------------------------------------
// Error state (thread-local message, like an exception message) and launch accounting.
// ------------------------------------------------------------------------------------------------------------
dn_status set_error(dn_status st, const char *fmt, ...);
dn_status cuda_error(cudaError_t err, const char *what);
cudaStream_t current_stream();
void set_thread_stream(cudaStream_t s);
bool check_errors_enabled();
extern std::atomic<int64_t> g_launch_count;
int sm_count();
// Per-device sticky index-error flag (device memory, one int).
int *index_error_flag();

#define DN_CUDA_TRY(expr)                                          \
    do {                                                           \
        cudaError_t _e = (expr);                                   \
        if (_e != cudaSuccess) return ::dn::cuda_error(_e, #expr); \
    } while (0)

// Every kernel launch goes through this so that dn_launch_count() is exact.
#define DN_LAUNCH(kernel, grid, block, smem, ...)                                          \
    do {                                                                                   \
        kernel<<<(grid), (block), (smem), ::dn::current_stream()>>>(__VA_ARGS__);          \
        ::dn::g_launch_count.fetch_add(1, std::memory_order_relaxed);                      \
    } while (0)

inline dn_status launch_status(const char *what) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();
        return cuda_error(e, what);
    }
    return DN_OK;
}

// Sharded launches (shard.cu, peer.cuh): the peer-store / exit-barrier block that the NEXT output-storing kernel
// launched by this thread has to carry. peer_take moves it into `out` and clears it (false: plain launch).
struct PeerSync;
bool peer_take(PeerSync &out);

// countTrue of a bool view into a device counter on the current stream, without a read-back (index.cu).
dn_status count_true_async(const dn_tensor *a, unsigned long long *dev_total);

// Stream-ordered scratch memory for multi-pass kernels (partials, block counts).
dn_status scratch_alloc(size_t nbytes, void **ptr);
void scratch_free(void *ptr);

// ------------------------------------------------------------------------------------------------------------
// dtype helpers
// ------------------------------------------------------------------------------------------------------------
inline int dtype_size(int dt) {
    switch (dt) {
    case DN_F32: case DN_I32: case DN_U32: return 4;
    case DN_F64: case DN_I64: case DN_U64: return 8;
    case DN_I16: case DN_U16: return 2;
    case DN_I8: case DN_U8: case DN_BOOL: return 1;
    default: return 0;
    }
}
inline const char *dtype_name(int dt) {
    static const char *names[] = {"single", "double", "sbyte", "byte", "int16", "uint16",
                                  "int32", "uint32", "int64", "uint64", "bool"};
    return (dt >= 0 && dt < DN_DTYPE_COUNT) ? names[dt] : "?";
}

// bool tensors are bytes (0/1) on the device; kernels use this tag type so that loads normalise (!= 0) and
// stores write exactly 0 or 1 (KernelCompiler.fs:190-193 marshals bool as one byte).
struct bool8 {
    uint8_t v;
    bool8() = default;
    __host__ __device__ bool8(bool b) : v(b ? 1 : 0) {}
    __host__ __device__ operator bool() const { return v != 0; }
};
static_assert(sizeof(bool8) == 1, "bool8 must be one byte");

template <int DT> struct CType;
template <> struct CType<DN_F32> { using type = float; };
template <> struct CType<DN_F64> { using type = double; };
template <> struct CType<DN_I8> { using type = int8_t; };
template <> struct CType<DN_U8> { using type = uint8_t; };
template <> struct CType<DN_I16> { using type = int16_t; };
template <> struct CType<DN_U16> { using type = uint16_t; };
template <> struct CType<DN_I32> { using type = int32_t; };
template <> struct CType<DN_U32> { using type = uint32_t; };
template <> struct CType<DN_I64> { using type = int64_t; };
template <> struct CType<DN_U64> { using type = uint64_t; };
template <> struct CType<DN_BOOL> { using type = bool8; };

bool tensor_valid(const dn_tensor *t);
bool same_shape(const dn_tensor *a, const dn_tensor *b);
int64_t num_elements(const dn_tensor *t);
inline char *data_ptr(const dn_tensor *t) {
    return static_cast<char *>(t->base) + t->offset * (int64_t)dtype_size(t->dtype);
}

// ------------------------------------------------------------------------------------------------------------
// Fast unsigned division by a runtime constant (n < 2^31, 1 <= d < 2^31): q = (umulhi(n, m) + n) >> s.
// ------------------------------------------------------------------------------------------------------------
struct FastDiv {
    uint32_t d, m, s;
    __host__ void init(uint32_t divisor) {
        d = divisor ? divisor : 1;
        for (s = 0; s < 32; ++s)
            if ((1u << s) >= d) break;
        uint64_t one = 1;
        m = (uint32_t)(((one << 32) * ((one << s) - d)) / d + 1);
    }
    __device__ __forceinline__ uint32_t div(uint32_t n) const { return (__umulhi(n, m) + n) >> s; }
};

}  // namespace dn
