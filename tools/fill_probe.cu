// fill_probe.cu — write-only bandwidth probe for B200: which store flavour / grid shape reaches the highest HBM
// write rate? (development tool; results recorded in DESIGN.md §4.1)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fill_probe tools/fill_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

enum { ST_DEFAULT, ST_CS, ST_WT, ST_NOALLOC, ST_V8, ST_CG, ST_EVICT_FIRST, ST_EVICT_LAST_NEVER };

template <int MODE>
__device__ __forceinline__ void store16(uint4 *p, uint4 v, uint64_t pol) {
    if constexpr (MODE == ST_DEFAULT) *p = v;
    else if constexpr (MODE == ST_CS) __stcs(p, v);
    else if constexpr (MODE == ST_WT) __stwt(p, v);
    else if constexpr (MODE == ST_CG) __stcg(p, v);
    else if constexpr (MODE == ST_NOALLOC)
        asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else if constexpr (MODE == ST_EVICT_FIRST)
        asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
    else *p = v;
}

// grid-stride, U stores of 16 B per thread per iteration, consecutive threads -> consecutive 16 B
template <int MODE, int U>
__global__ void __launch_bounds__(256) fill_gs(uint4 *dst, size_t n16, uint4 v) {
    uint64_t pol = 0;
    if constexpr (MODE == ST_EVICT_FIRST) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    const size_t stride = (size_t)gridDim.x * 256 * U;
    for (size_t base = (size_t)blockIdx.x * 256 * U + threadIdx.x; base < n16; base += stride) {
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (base + (size_t)u * 256 < n16) store16<MODE>(dst + base + (size_t)u * 256, v, pol);
    }
}

// 32 B per thread contiguous (two adjacent 16-B stores, like the product kernel's work item)
template <int MODE, int U>
__global__ void __launch_bounds__(256) fill_32(uint4 *dst, size_t n16, uint4 v) {
    const size_t n32 = n16 / 2;
    const size_t stride = (size_t)gridDim.x * 256 * U;
    for (size_t base = (size_t)blockIdx.x * 256 * U + threadIdx.x; base < n32; base += stride) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t i = base + (size_t)u * 256;
            if (i < n32) {
                if constexpr (MODE == ST_V8) {
                    asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst + 2 * i), "r"(v.x), "r"(v.y), "r"(v.z),
                                 "r"(v.w), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                                 : "memory");
                } else {
                    store16<MODE>(dst + 2 * i, v, 0);
                    store16<MODE>(dst + 2 * i + 1, v, 0);
                }
            }
        }
    }
}

// each CTA owns one contiguous chunk (blocked instead of grid-strided)
template <int MODE>
__global__ void __launch_bounds__(256) fill_blocked(uint4 *dst, size_t n16, uint4 v, size_t per_cta) {
    const size_t b = (size_t)blockIdx.x * per_cta;
    size_t e = b + per_cta;
    if (e > n16) e = n16;
    for (size_t i = b + threadIdx.x; i < e; i += 256 * 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (i + u * 256 < e) store16<MODE>(dst + i + u * 256, v, 0);
    }
}

// TMA bulk store: smem -> global, 16 KiB per bulk op
__global__ void __launch_bounds__(128) fill_bulk(char *dst, size_t nbytes, uint4 v, int chunk) {
    extern __shared__ __align__(128) unsigned char sm[];
    for (int i = threadIdx.x; i < chunk / 16; i += blockDim.x) reinterpret_cast<uint4 *>(sm)[i] = v;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(sm);
        int pending = 0;
        for (size_t off = (size_t)blockIdx.x * chunk; off + chunk <= nbytes; off += (size_t)gridDim.x * chunk) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + off), "r"(saddr), "r"(chunk) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (++pending >= 8) {
                asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
                pending = 4;
            }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

template <class L>
float time_it(L launch, int reps = 7) {
    cudaEvent_t s, e;
    cudaEventCreate(&s);
    cudaEventCreate(&e);
    launch();
    cudaDeviceSynchronize();
    float best = 1e9f, tot = 0;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(s);
        launch();
        cudaEventRecord(e);
        cudaEventSynchronize(e);
        float ms;
        cudaEventElapsedTime(&ms, s, e);
        if (ms < best) best = ms;
        tot += ms;
    }
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) printf("  CUDA error: %s\n", cudaGetErrorString(err));
    return best;
}

int main() {
    const size_t nbytes = (size_t)1 << 30;
    const size_t n16 = nbytes / 16;
    char *buf;
    cudaMalloc(&buf, nbytes);
    uint4 *dst = reinterpret_cast<uint4 *>(buf);
    const uint4 v = make_uint4(1, 2, 3, 4);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    auto report = [&](const char *name, float ms) { printf("%-44s %8.3f ms %8.1f GB/s\n", name, ms, nbytes / ms / 1e6); };

    report("cudaMemsetAsync", time_it([&] { cudaMemsetAsync(buf, 1, nbytes); }));
    for (int per_sm : {2, 4, 8, 16, 32}) {
        const int grid = sms * per_sm;
        char name[96];
        snprintf(name, sizeof name, "gs st.default U=4 grid=%dxSM", per_sm);
        report(name, time_it([&] { fill_gs<ST_DEFAULT, 4><<<grid, 256>>>(dst, n16, v); }));
        snprintf(name, sizeof name, "gs st.cs U=4 grid=%dxSM", per_sm);
        report(name, time_it([&] { fill_gs<ST_CS, 4><<<grid, 256>>>(dst, n16, v); }));
    }
    const int grid = sms * 8;
    report("gs st.cs U=1", time_it([&] { fill_gs<ST_CS, 1><<<grid, 256>>>(dst, n16, v); }));
    report("gs st.cs U=2", time_it([&] { fill_gs<ST_CS, 2><<<grid, 256>>>(dst, n16, v); }));
    report("gs st.cs U=8", time_it([&] { fill_gs<ST_CS, 8><<<grid, 256>>>(dst, n16, v); }));
    report("gs st.wt U=4", time_it([&] { fill_gs<ST_WT, 4><<<grid, 256>>>(dst, n16, v); }));
    report("gs st.cg U=4", time_it([&] { fill_gs<ST_CG, 4><<<grid, 256>>>(dst, n16, v); }));
    report("gs st.L1::no_allocate U=4", time_it([&] { fill_gs<ST_NOALLOC, 4><<<grid, 256>>>(dst, n16, v); }));
    report("gs st L2::evict_first U=4", time_it([&] { fill_gs<ST_EVICT_FIRST, 4><<<grid, 256>>>(dst, n16, v); }));
    report("32B/thread st.cs x2 U=2 (product)", time_it([&] { fill_32<ST_CS, 2><<<sms * 32, 256>>>(dst, n16, v); }));
    report("32B/thread st.default x2 U=2", time_it([&] { fill_32<ST_DEFAULT, 2><<<sms * 32, 256>>>(dst, n16, v); }));
    report("32B/thread st.v8 U=2", time_it([&] { fill_32<ST_V8, 2><<<sms * 32, 256>>>(dst, n16, v); }));
    report("32B/thread st.v8 U=4 grid 8xSM", time_it([&] { fill_32<ST_V8, 4><<<sms * 8, 256>>>(dst, n16, v); }));
    for (int per_sm : {1, 2, 4, 8}) {
        const int g = sms * per_sm;
        const size_t per_cta = (n16 + g - 1) / g;
        char name[96];
        snprintf(name, sizeof name, "blocked st.cs grid=%dxSM", per_sm);
        report(name, time_it([&] { fill_blocked<ST_CS><<<g, 256>>>(dst, n16, v, per_cta); }));
    }
    for (int chunk : {4096, 16384, 65536}) {
        cudaFuncSetAttribute(fill_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, chunk);
        for (int per_sm : {1, 2, 4}) {
            char name[96];
            snprintf(name, sizeof name, "TMA bulk store chunk=%d grid=%dxSM", chunk, per_sm);
            report(name, time_it([&] { fill_bulk<<<sms * per_sm, 128, chunk>>>(buf, nbytes, v, chunk); }));
        }
    }
    // reference: copy (read+write) through cudaMemcpy
    char *src;
    cudaMalloc(&src, nbytes);
    float ms = time_it([&] { cudaMemcpyAsync(buf, src, nbytes, cudaMemcpyDeviceToDevice); });
    printf("%-44s %8.3f ms %8.1f GB/s (read+write)\n", "cudaMemcpy D2D", ms, 2.0 * nbytes / ms / 1e6);
    return 0;
}
