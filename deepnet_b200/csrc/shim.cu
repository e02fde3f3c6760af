// shim.cu — device / stream / storage entry points of the C ABI (include/dn_tensor.h).
//
// Replaces module Cuda (Tensor/Tensor/Cuda/CudaUtils.fs:19-210: context, stream callbacks, GC-pressure allocator),
// Cfg (Tensor/Tensor/Cuda/CudaCfg.fs:14-60), CudaRegMem (Tensor/Tensor/Cuda/CudaRegMem.fs:123-146) and the
// storage plumbing of TensorCudaStorage (Tensor/Tensor/Cuda/CudaBackend.fs:51-108).
//
// Design: one primary context per device; storage comes from the device's stream-ordered memory pool
// (cudaMallocAsync / cudaFreeAsync on the calling thread's stream) with the release threshold lifted, so a free
// right after the last consumer was enqueued is safe and allocation in a steady-state loop costs no driver call.
// This is what makes the reference's event-per-operand keep-alive machinery (CudaUtils.fs:122-177) unnecessary.
#include "common.cuh"

#include <mutex>

namespace dn {

namespace {
thread_local char t_err[1024] = "";
thread_local cudaStream_t t_stream = nullptr;
thread_local int t_check_errors = 0;

constexpr int kMaxDevices = 64;
struct DeviceState {
    std::once_flag once;
    bool ok = false;
    int sms = 0;
    int *index_error = nullptr;
    cudaError_t init_err = cudaSuccess;
};
DeviceState g_dev[kMaxDevices];
}  // namespace

std::atomic<int64_t> g_launch_count{0};

dn_status set_error(dn_status st, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof t_err, fmt, ap);
    va_end(ap);
    return st;
}

dn_status cuda_error(cudaError_t err, const char *what) {
    if (err == cudaErrorMemoryAllocation)
        return set_error(DN_ERR_OUT_OF_MEMORY, "CUDA memory allocation failed in %s", what);
    if (err == cudaErrorNoDevice || err == cudaErrorInsufficientDriver || err == cudaErrorInvalidDevice)
        return set_error(DN_ERR_NO_DEVICE, "Cannot create CUDA context: %s (%s)", cudaGetErrorString(err), what);
    return set_error(DN_ERR_CUDA, "CUDA error %s: %s (%s)", cudaGetErrorName(err), cudaGetErrorString(err), what);
}

cudaStream_t current_stream() { return t_stream; }
bool check_errors_enabled() { return t_check_errors != 0; }

static DeviceState *device_state() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
    DeviceState &s = g_dev[dev];
    std::call_once(s.once, [&] {
        cudaDeviceProp prop;
        cudaError_t e = cudaGetDeviceProperties(&prop, dev);
        if (e == cudaSuccess) {
            s.sms = prop.multiProcessorCount;
            cudaMemPool_t pool;
            e = cudaDeviceGetDefaultMemPool(&pool, dev);
            if (e == cudaSuccess) {
                uint64_t threshold = UINT64_MAX;  // keep freed blocks cached in the pool
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
            }
        }
        if (e == cudaSuccess) e = cudaMalloc((void **)&s.index_error, sizeof(int));
        if (e == cudaSuccess) e = cudaMemset(s.index_error, 0, sizeof(int));
        s.init_err = e;
        s.ok = e == cudaSuccess;
    });
    return &s;
}

int sm_count() {
    DeviceState *s = device_state();
    return (s && s->ok && s->sms > 0) ? s->sms : 148;
}

int *index_error_flag() {
    DeviceState *s = device_state();
    return (s && s->ok) ? s->index_error : nullptr;
}

dn_status scratch_alloc(size_t nbytes, void **ptr) {
    DN_CUDA_TRY(cudaMallocAsync(ptr, nbytes ? nbytes : 1, current_stream()));
    return DN_OK;
}
void scratch_free(void *ptr) {
    if (ptr) cudaFreeAsync(ptr, current_stream());
}

bool tensor_valid(const dn_tensor *t) {
    if (!t || t->ndims < 0 || t->ndims > DN_MAX_DIMS) return false;
    if (t->dtype < 0 || t->dtype >= DN_DTYPE_COUNT) return false;
    for (int d = 0; d < t->ndims; ++d)
        if (t->shape[d] < 0) return false;
    return true;
}

bool same_shape(const dn_tensor *a, const dn_tensor *b) {
    if (a->ndims != b->ndims) return false;
    for (int d = 0; d < a->ndims; ++d)
        if (a->shape[d] != b->shape[d]) return false;
    return true;
}

int64_t num_elements(const dn_tensor *t) {
    int64_t n = 1;
    for (int d = 0; d < t->ndims; ++d) n *= t->shape[d];
    return n;
}

}  // namespace dn

using namespace dn;

extern "C" {

const char *dn_last_error(void) { return t_err; }
int64_t dn_launch_count(void) { return g_launch_count.load(std::memory_order_relaxed); }
const char *dn_version(void) { return "deepnet_b200 0.1 (sm_100a)"; }

dn_status dn_device_count(int32_t *count) {
    if (!count) return set_error(DN_ERR_INVALID_ARG, "dn_device_count: null argument");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        return cuda_error(e, "cudaGetDeviceCount");
    }
    *count = n;
    return DN_OK;
}

dn_status dn_set_device(int32_t device) {
    DN_CUDA_TRY(cudaSetDevice(device));
    return DN_OK;
}

dn_status dn_get_device(int32_t *device) {
    if (!device) return set_error(DN_ERR_INVALID_ARG, "dn_get_device: null argument");
    int d = 0;
    DN_CUDA_TRY(cudaGetDevice(&d));
    *device = d;
    return DN_OK;
}

dn_status dn_init(int32_t device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return cuda_error(e, "cudaGetDeviceCount");
    if (n == 0 || device < 0 || device >= n)
        return set_error(DN_ERR_NO_DEVICE, "Cannot create CUDA context: device %d of %d not available", device, n);
    DN_CUDA_TRY(cudaSetDevice(device));
    DN_CUDA_TRY(cudaFree(nullptr));  // force primary context creation
    DeviceState *s = device_state();
    if (!s) return set_error(DN_ERR_NO_DEVICE, "Cannot create CUDA context: no device state");
    if (!s->ok) return cuda_error(s->init_err, "device initialisation");
    return DN_OK;
}

dn_status dn_set_stream(void *stream) {
    t_stream = static_cast<cudaStream_t>(stream);
    return DN_OK;
}

dn_status dn_get_stream(void **stream) {
    if (!stream) return set_error(DN_ERR_INVALID_ARG, "dn_get_stream: null argument");
    *stream = t_stream;
    return DN_OK;
}

dn_status dn_sync(void) {
    DN_CUDA_TRY(cudaStreamSynchronize(t_stream));
    return DN_OK;
}

dn_status dn_set_check_errors(int32_t enabled) {
    t_check_errors = enabled ? 1 : 0;
    return DN_OK;
}

dn_status dn_poll_index_error(int32_t *had_error) {
    if (!had_error) return set_error(DN_ERR_INVALID_ARG, "dn_poll_index_error: null argument");
    int *flag = index_error_flag();
    if (!flag) return set_error(DN_ERR_NO_DEVICE, "device not initialised");
    int h = 0;
    DN_CUDA_TRY(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, t_stream));
    DN_CUDA_TRY(cudaStreamSynchronize(t_stream));
    if (h) DN_CUDA_TRY(cudaMemsetAsync(flag, 0, sizeof(int), t_stream));
    *had_error = h;
    return DN_OK;
}

dn_status dn_alloc(int64_t nbytes, void **ptr) {
    if (!ptr) return set_error(DN_ERR_INVALID_ARG, "dn_alloc: null argument");
    if (!device_state()) return set_error(DN_ERR_NO_DEVICE, "Cannot create CUDA context");
    size_t n = nbytes > 0 ? (size_t)nbytes : 1;  // CUDA cannot allocate size zero (CudaBackend.fs:56-58)
    cudaError_t e = cudaMallocAsync(ptr, n, t_stream);
    if (e == cudaErrorMemoryAllocation) {
        // The reference forces a GC and retries (CudaUtils.fs:183-210); the analogue here is to return the
        // pool's cached blocks to the driver and retry once.
        cudaGetLastError();
        cudaStreamSynchronize(t_stream);
        int dev = 0;
        cudaGetDevice(&dev);
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
        e = cudaMallocAsync(ptr, n, t_stream);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        *ptr = nullptr;
        return cuda_error(e, "dn_alloc");
    }
    return DN_OK;
}

dn_status dn_free(void *ptr) {
    if (!ptr) return DN_OK;
    DN_CUDA_TRY(cudaFreeAsync(ptr, t_stream));
    return DN_OK;
}

dn_status dn_alloc_host(int64_t nbytes, void **ptr) {
    if (!ptr) return set_error(DN_ERR_INVALID_ARG, "dn_alloc_host: null argument");
    DN_CUDA_TRY(cudaMallocHost(ptr, nbytes > 0 ? (size_t)nbytes : 1));
    return DN_OK;
}

dn_status dn_free_host(void *ptr) {
    if (!ptr) return DN_OK;
    DN_CUDA_TRY(cudaFreeHost(ptr));
    return DN_OK;
}

dn_status dn_memset_zero(void *ptr, int64_t nbytes) {
    if (nbytes <= 0) return DN_OK;
    DN_CUDA_TRY(cudaMemsetAsync(ptr, 0, (size_t)nbytes, t_stream));
    return DN_OK;
}

// Transfer (CudaBackend.fs:206-270). With pinned host memory the copy is asynchronous on the stream; with
// pageable memory the runtime stages it and returns once the host buffer is reusable — in both cases the
// result is ordered with the kernels on the calling thread's stream. D2H blocks until the data has arrived,
// which is what a caller that is about to read the host buffer needs.
dn_status dn_memcpy_h2d(void *dst_dev, const void *src_host, int64_t nbytes) {
    if (nbytes <= 0) return DN_OK;
    DN_CUDA_TRY(cudaMemcpyAsync(dst_dev, src_host, (size_t)nbytes, cudaMemcpyHostToDevice, t_stream));
    return DN_OK;
}

dn_status dn_memcpy_d2h(void *dst_host, const void *src_dev, int64_t nbytes) {
    if (nbytes <= 0) return DN_OK;
    DN_CUDA_TRY(cudaMemcpyAsync(dst_host, src_dev, (size_t)nbytes, cudaMemcpyDeviceToHost, t_stream));
    DN_CUDA_TRY(cudaStreamSynchronize(t_stream));
    return DN_OK;
}

dn_status dn_memcpy_d2h_async(void *dst_host, const void *src_dev, int64_t nbytes) {
    if (nbytes <= 0) return DN_OK;
    DN_CUDA_TRY(cudaMemcpyAsync(dst_host, src_dev, (size_t)nbytes, cudaMemcpyDeviceToHost, t_stream));
    return DN_OK;
}

dn_status dn_event_create(void **event) {
    if (!event) return set_error(DN_ERR_INVALID_ARG, "dn_event_create: null argument");
    cudaEvent_t e;
    DN_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    *event = e;
    return DN_OK;
}

dn_status dn_event_destroy(void *event) {
    if (event) DN_CUDA_TRY(cudaEventDestroy(static_cast<cudaEvent_t>(event)));
    return DN_OK;
}

dn_status dn_event_record(void *event) {
    DN_CUDA_TRY(cudaEventRecord(static_cast<cudaEvent_t>(event), t_stream));
    return DN_OK;
}

dn_status dn_stream_wait_event(void *event) {
    DN_CUDA_TRY(cudaStreamWaitEvent(t_stream, static_cast<cudaEvent_t>(event), 0));
    return DN_OK;
}

dn_status dn_memcpy_d2d(void *dst_dev, const void *src_dev, int64_t nbytes) {
    if (nbytes <= 0) return DN_OK;
    DN_CUDA_TRY(cudaMemcpyAsync(dst_dev, src_dev, (size_t)nbytes, cudaMemcpyDeviceToDevice, t_stream));
    return DN_OK;
}

static dn_status item_address(const dn_tensor *t, const int64_t *pos, char **addr) {
    if (!tensor_valid(t) || (t->ndims > 0 && !pos)) return set_error(DN_ERR_INVALID_ARG, "item: bad argument");
    int64_t off = t->offset;
    for (int d = 0; d < t->ndims; ++d) {
        if (pos[d] < 0 || pos[d] >= t->shape[d])
            return set_error(DN_ERR_INDEX_OUT_OF_RANGE, "index out of range for tensor dimension %d", d);
        off += pos[d] * t->stride[d];
    }
    *addr = static_cast<char *>(t->base) + off * dtype_size(t->dtype);
    return DN_OK;
}

dn_status dn_get_item(const dn_tensor *t, const int64_t *pos, void *value) {
    char *addr = nullptr;
    dn_status st = item_address(t, pos, &addr);
    if (st != DN_OK) return st;
    if (!value) return set_error(DN_ERR_INVALID_ARG, "dn_get_item: null value");
    DN_CUDA_TRY(cudaMemcpyAsync(value, addr, dtype_size(t->dtype), cudaMemcpyDeviceToHost, t_stream));
    DN_CUDA_TRY(cudaStreamSynchronize(t_stream));
    if (t->dtype == DN_BOOL) *static_cast<uint8_t *>(value) = *static_cast<uint8_t *>(value) ? 1 : 0;
    return DN_OK;
}

dn_status dn_set_item(const dn_tensor *t, const int64_t *pos, const void *value) {
    char *addr = nullptr;
    dn_status st = item_address(t, pos, &addr);
    if (st != DN_OK) return st;
    if (!value) return set_error(DN_ERR_INVALID_ARG, "dn_set_item: null value");
    uint64_t tmp = 0;
    memcpy(&tmp, value, dtype_size(t->dtype));
    if (t->dtype == DN_BOOL) tmp = (tmp & 0xff) ? 1 : 0;
    DN_CUDA_TRY(cudaMemcpyAsync(addr, &tmp, dtype_size(t->dtype), cudaMemcpyHostToDevice, t_stream));
    DN_CUDA_TRY(cudaStreamSynchronize(t_stream));
    return DN_OK;
}

}  // extern "C"
