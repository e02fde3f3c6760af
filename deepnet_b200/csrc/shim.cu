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
#include <vector>

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
    // storage lifetime across threads and streams (dn_free / dn_free_deferred)
    std::mutex mu;
    std::vector<cudaStream_t> streams;   // streams the library has been told to use on this device (+ the null stream)
    std::vector<void *> deferred;        // frees queued by threads that do not own a stream here (finalizers)
    std::atomic<int> ndeferred{0};
    cudaEvent_t fence = nullptr;
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
// Library-internal switch of the calling thread's stream (shard.cu binds a rank's stream for the duration of a call):
// unlike dn_set_stream it does not enter the stream into the device's list of streams that storage releases fence.
void set_thread_stream(cudaStream_t s) { t_stream = s; }
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
                // A block freed on stream A may be handed to stream B only once the free has completed (or an
                // event dependency between the streams already exists) — never by making B wait on A behind the
                // caller's back. Hidden cross-stream dependencies would serialise the per-thread streams the
                // reference's threading model relies on (CudaCfg.fs:25-27) and deadlock ranks that share a device
                // (dn_shard_*: rank A's stream waits for a signal that rank B's kernel has yet to send).
                int off = 0;
                cudaMemPoolSetAttribute(pool, cudaMemPoolReuseAllowInternalDependencies, &off);
            }
        }
        if (e == cudaSuccess) e = cudaMalloc((void **)&s.index_error, sizeof(int));
        if (e == cudaSuccess) e = cudaMemset(s.index_error, 0, sizeof(int));
        s.streams.push_back(nullptr);  // the default stream is always a candidate holder of work
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

// Orders the calling thread's stream after everything enqueued so far on every OTHER stream the library knows on
// this device. Used before a free whose last consumer may have been enqueued by another thread / on another stream
// (the reference keeps operands alive with an event per operand instead, CudaUtils.fs:122-177).
static void fence_other_streams(DeviceState &s) {
    for (cudaStream_t k : s.streams) {
        if (k == t_stream) continue;
        if (!s.fence && cudaEventCreateWithFlags(&s.fence, cudaEventDisableTiming) != cudaSuccess) break;
        if (cudaEventRecord(s.fence, k) != cudaSuccess || cudaStreamWaitEvent(t_stream, s.fence, 0) != cudaSuccess)
            cudaGetLastError();  // a stream that no longer exists has nothing in flight
    }
}

// Frees queued by dn_free_deferred for the current device, released by a thread that owns a stream on it.
static void drain_deferred(DeviceState &s) {
    if (s.ndeferred.load(std::memory_order_acquire) == 0) return;
    std::lock_guard<std::mutex> lock(s.mu);
    if (s.deferred.empty()) return;
    fence_other_streams(s);
    for (void *p : s.deferred) cudaFreeAsync(p, t_stream);
    s.deferred.clear();
    s.ndeferred.store(0, std::memory_order_release);
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
    if (DeviceState *s = device_state()) {
        std::lock_guard<std::mutex> lock(s->mu);
        bool known = false;
        for (cudaStream_t k : s->streams) known = known || k == t_stream;
        if (!known) s->streams.push_back(t_stream);
    }
    return DN_OK;
}

dn_status dn_release_stream(void *stream) {
    if (DeviceState *s = device_state()) {
        std::lock_guard<std::mutex> lock(s->mu);
        for (size_t i = 0; i < s->streams.size(); ++i)
            if (s->streams[i] == static_cast<cudaStream_t>(stream)) {
                s->streams.erase(s->streams.begin() + i);
                break;
            }
    }
    if (t_stream == static_cast<cudaStream_t>(stream)) t_stream = nullptr;
    return DN_OK;
}

dn_status dn_get_stream(void **stream) {
    if (!stream) return set_error(DN_ERR_INVALID_ARG, "dn_get_stream: null argument");
    *stream = t_stream;
    return DN_OK;
}

dn_status dn_sync(void) {
    if (DeviceState *s = device_state()) drain_deferred(*s);
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
    DeviceState *ds = device_state();
    if (!ds) return set_error(DN_ERR_NO_DEVICE, "Cannot create CUDA context");
    drain_deferred(*ds);
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

// Stream-ordered on the calling thread's stream. When the library has been given more than one stream on this
// device (dn_set_stream from several threads, or a thread that switched streams), the free is first ordered after
// everything those streams hold, so a block is never recycled under a kernel enqueued elsewhere.
dn_status dn_free(void *ptr) {
    if (!ptr) return DN_OK;
    if (DeviceState *s = device_state()) {
        std::lock_guard<std::mutex> lock(s->mu);
        bool other = false;
        for (cudaStream_t k : s->streams) other = other || k != t_stream;
        if (other) fence_other_streams(*s);
    }
    DN_CUDA_TRY(cudaFreeAsync(ptr, t_stream));
    return DN_OK;
}

// For threads that own no stream on the storage's device — .NET finalizers (CudaBackend.fs:73-74 frees from the
// finalizer thread), destructors run by a GC thread: looks up the owning device, queues the pointer there and
// returns without touching any stream. The next dn_alloc / dn_sync issued by a thread working on that device
// releases the queue behind a fence over all of the device's known streams.
dn_status dn_free_deferred(void *ptr) {
    if (!ptr) return DN_OK;
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, ptr);
    if (e != cudaSuccess || attr.type != cudaMemoryTypeDevice || attr.device < 0 || attr.device >= kMaxDevices) {
        cudaGetLastError();
        return set_error(DN_ERR_INVALID_ARG, "dn_free_deferred: not a device allocation");
    }
    DeviceState &s = g_dev[attr.device];
    std::lock_guard<std::mutex> lock(s.mu);
    s.deferred.push_back(ptr);
    s.ndeferred.fetch_add(1, std::memory_order_release);
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

// CudaRegMem.register (CudaRegMem.fs:123-146): page-locks existing host memory (a pinned managed array) so that
// transfers are asynchronous DMA. DN_ERR_INVALID_ARG when the range cannot be registered (misaligned, already
// registered, ...) — the caller falls back to plain pinning, like CannotCudaRegisterMemoryException does.
dn_status dn_host_register(void *ptr, int64_t nbytes) {
    if (!ptr || nbytes <= 0) return set_error(DN_ERR_INVALID_ARG, "dn_host_register: bad argument");
    cudaError_t e = cudaHostRegister(ptr, (size_t)nbytes, cudaHostRegisterDefault);
    if (e == cudaErrorHostMemoryAlreadyRegistered) {
        cudaGetLastError();
        return DN_OK;
    }
    if (e == cudaErrorInvalidValue || e == cudaErrorNotSupported) {
        cudaGetLastError();
        return set_error(DN_ERR_INVALID_ARG, "cannot register host memory with CUDA: %s", cudaGetErrorString(e));
    }
    DN_CUDA_TRY(e);
    return DN_OK;
}

dn_status dn_host_unregister(void *ptr) {
    if (!ptr) return DN_OK;
    cudaError_t e = cudaHostUnregister(ptr);
    if (e == cudaErrorHostMemoryNotRegistered) {
        cudaGetLastError();
        return DN_OK;
    }
    DN_CUDA_TRY(e);
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

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// Transfer of arbitrary views (ITensorBackend.Transfer, CudaBackend.fs:206-270). The reference makes the source
// row-major with a HOST-side copy and the target through a temporary device tensor + Copy; here both sides are
// handled natively so every host language gets it: the host view is moved as raw bytes (directly when it is dense
// in memory, else packed / unpacked by a plain strided memcpy loop into pinned staging — data movement, no
// arithmetic) and the layout change happens on the device with the strided copy kernel (dn_copy).
// ---------------------------------------------------------------------------------------------------------------
namespace {

bool view_is_c_contiguous(const dn_tensor *t) {
    int64_t expect = 1;
    for (int d = t->ndims - 1; d >= 0; --d) {
        if (t->shape[d] != 1 && t->stride[d] != expect) return false;
        expect *= t->shape[d];
    }
    return true;
}

dn_tensor contiguous_like(void *base, const dn_tensor *t) {
    dn_tensor c = *t;
    c.base = base;
    c.offset = 0;
    int64_t st = 1;
    for (int d = t->ndims - 1; d >= 0; --d) {
        c.stride[d] = st;
        st *= t->shape[d];
    }
    return c;
}

// Walks a host view in logical row-major order, moving runs of its innermost dimension to / from a dense buffer.
void host_pack(char *dense, const dn_tensor *v, bool to_dense) {
    const int isz = dtype_size(v->dtype);
    const int nd = v->ndims;
    char *base = static_cast<char *>(v->base) + v->offset * isz;
    if (nd == 0) {
        if (to_dense) memcpy(dense, base, isz); else memcpy(base, dense, isz);
        return;
    }
    const int64_t inner = v->shape[nd - 1], istride = v->stride[nd - 1];
    int64_t idx[DN_MAX_DIMS] = {};
    int64_t outer = 1;
    for (int d = 0; d < nd - 1; ++d) outer *= v->shape[d];
    for (int64_t o = 0; o < outer; ++o) {
        int64_t off = 0;
        for (int d = 0; d < nd - 1; ++d) off += idx[d] * v->stride[d];
        char *row = base + off * isz;
        char *out = dense + o * inner * isz;
        if (istride == 1) {
            if (to_dense) memcpy(out, row, (size_t)inner * isz); else memcpy(row, out, (size_t)inner * isz);
        } else {
            for (int64_t i = 0; i < inner; ++i) {
                if (to_dense) memcpy(out + i * isz, row + i * istride * isz, isz);
                else memcpy(row + i * istride * isz, out + i * isz, isz);
            }
        }
        for (int d = nd - 2; d >= 0; --d) {
            if (++idx[d] < v->shape[d]) break;
            idx[d] = 0;
        }
    }
}

dn_status check_transfer(const dn_tensor *a, const dn_tensor *b, const char *what) {
    if (!tensor_valid(a) || !tensor_valid(b)) return set_error(DN_ERR_INVALID_ARG, "%s: bad argument", what);
    if (a->dtype != b->dtype || !same_shape(a, b))
        return set_error(DN_ERR_SHAPE_MISMATCH, "%s: source and target must have the same shape and type", what);
    return DN_OK;
}

}  // namespace

extern "C" {

dn_status dn_transfer_h2d(const dn_tensor *dev_t, const dn_tensor *host_t) {
    dn_status st = check_transfer(dev_t, host_t, "Transfer (host -> device)");
    if (st != DN_OK) return st;
    const int64_t n = num_elements(dev_t), nbytes = n * dtype_size(dev_t->dtype);
    if (n == 0) return DN_OK;
    const bool host_c = view_is_c_contiguous(host_t), dev_c = view_is_c_contiguous(dev_t);
    void *pinned = nullptr, *staging = nullptr;
    const char *src = data_ptr(host_t);
    if (!host_c) {  // pack the host view into pinned staging (kept until the copy has left the host)
        DN_CUDA_TRY(cudaMallocHost(&pinned, (size_t)nbytes));
        host_pack(static_cast<char *>(pinned), host_t, true);
        src = static_cast<const char *>(pinned);
    }
    if (dev_c) {
        st = dn_memcpy_h2d(data_ptr(dev_t), src, nbytes);
    } else {
        st = scratch_alloc((size_t)nbytes, &staging);
        if (st == DN_OK) st = dn_memcpy_h2d(staging, src, nbytes);
        if (st == DN_OK) {
            const dn_tensor c = contiguous_like(staging, dev_t);
            st = dn_copy(dev_t, &c);
        }
        scratch_free(staging);
    }
    if (pinned) {
        cudaStreamSynchronize(t_stream);
        cudaFreeHost(pinned);
    }
    return st;
}

dn_status dn_transfer_d2h(const dn_tensor *host_t, const dn_tensor *dev_t) {
    dn_status st = check_transfer(host_t, dev_t, "Transfer (device -> host)");
    if (st != DN_OK) return st;
    const int64_t n = num_elements(dev_t), nbytes = n * dtype_size(dev_t->dtype);
    if (n == 0) return DN_OK;
    const bool host_c = view_is_c_contiguous(host_t), dev_c = view_is_c_contiguous(dev_t);
    void *pinned = nullptr, *staging = nullptr;
    const char *src = data_ptr(dev_t);
    if (!dev_c) {  // make the device side row-major first
        st = scratch_alloc((size_t)nbytes, &staging);
        if (st != DN_OK) return st;
        const dn_tensor c = contiguous_like(staging, dev_t);
        st = dn_copy(&c, dev_t);
        src = static_cast<const char *>(staging);
    }
    if (st == DN_OK) {
        if (host_c) {
            st = dn_memcpy_d2h(data_ptr(host_t), src, nbytes);
        } else {
            cudaError_t e = cudaMallocHost(&pinned, (size_t)nbytes);
            if (e != cudaSuccess) st = cuda_error(e, "cudaMallocHost");
            if (st == DN_OK) st = dn_memcpy_d2h(pinned, src, nbytes);  // blocking
            if (st == DN_OK) host_pack(static_cast<char *>(pinned), host_t, false);
            if (pinned) cudaFreeHost(pinned);
        }
    }
    scratch_free(staging);
    return st;
}

}  // extern "C"

extern "C" {

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
