// gemm.cu — MatMatDot / BatchedMatMatDot / MatVecDot / VecVecDot (Tensor/Tensor/TensorBackend.fs:137-140).
//
// Replaces the cuBLAS call sites of the reference (CudaBackend.fs:383-449, CudaBLAS.fs:17-64) and the operand
// staging of BlasSupport.fs:148-190 (column-major temporaries + strided copy-back of every row-major result).
//
// float32 MatMatDot runs on the 5th-generation tensor cores: a persistent, warp-specialised tcgen05 kernel
//   warp 0      TMA producer: cp.async.bulk.tensor 2-D loads of A[128 x 32] and B[BN x 32] fp32 tiles (128-byte
//               swizzle) into a multi-stage shared-memory ring, completion on mbarriers;
//   warp 1      MMA issuer: one elected thread issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=BN, K=8) on
//               shared-memory descriptors, fp32 accumulators live in TMEM (two stages, so the epilogue of tile i
//               overlaps the main loop of tile i+1); tcgen05.commit releases smem slots / publishes accumulators;
//   warps 2-5   epilogue: tcgen05.ld 32 lanes x 32 columns -> registers -> row-major C written DIRECTLY with its
//               own strides (the reference needs a column-major temp and a copy-back: BlasSupport.fs:181).
// C = A·B with A[M,K] and B[K,N] given as arbitrary strided views: an operand whose K axis is contiguous and
// TMA-aligned is consumed in place ("K-major"); anything else (an N-contiguous B, a transposed A, odd pitches) is
// first repacked K-major by the element-wise transpose kernel (dn_copy) into stream-ordered scratch.
// Precision (dn_set_math_mode): DN_MATH_FP32, the default — the exact SIMT kernel for small problems, 3xTF32 (hi/lo
// split of both operands, three MMAs per k-step; ~1e-5 relative, limited by the tensor core's truncating
// accumulator) for large ones; DN_MATH_FP32_STRICT — SIMT for every size; DN_MATH_TF32 reads the inputs once as
// TF32 (10-bit mantissa), fp32 accumulation: rel 1e-2 of the fp64 oracle per BASELINE.json north_star, at three
// times the throughput of the default.
// float64 has no tcgen05 path: a shared-memory tiled SIMT kernel.
#include <cuda.h>

#include <cstdlib>

#include "ew_ops.cuh"

using namespace dn;

namespace {

// ---------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded spin: a protocol bug traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (uint32_t spin = 0;; ++spin) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (ok) return;
        if (spin > (1u << 28)) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_c), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// K-major, 128-byte swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// start>>4 [0,14) | LBO>>4 [16,30) = 1 (unused for swizzled K-major) | SBO>>4 [32,46) = 1024 B (8 rows x 128 B)
// | version = 1 [46,48) | layout_type = SWIZZLE_128B (2) [61,64)
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

// MN-major tf32 operands: the only shared-memory layout UMMA accepts is the 128-byte swizzle with a 32-byte atom
// (layout_type 1, cute Swizzle<2,5,2>; TMA: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B). The tile is stored as
// (MN extent / 32) consecutive blocks of [32 k-rows x 128 B], each row holding 32 consecutive MN elements — exactly
// what a TMA box {32 MN, 32 K} writes. Canonical form ((8,n),(4,k)):((1,LBO),(8,SBO)) in 16-byte units:
// LBO = one block = 4096 B, SBO = one 4-row k-group = 512 B. One tf32 MMA (K = 8) consumes two k-groups, so the
// start address advances by 1024 B per k-step.
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(4096 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) |
           (1ull << 46) | (1ull << 61);
}

constexpr int kBM = 128;          // UMMA M
constexpr int kBK = 32;           // fp32 elements per k-block = one 128-byte swizzle row
constexpr int kGemmThreads = 192; // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue

struct GemmParams {
    float *c;
    int64_t ldc_m, ldc_n;  // element strides of C
    int32_t M, N, K;
    int32_t tiles_m, tiles_n;
    // batched launch of the CTA-pair kernel: tile index = batch * tiles_m * tiles_n + tile in matrix
    int32_t nbatch = 1;
    int32_t a_batched = 0, b_batched = 0;  // 0: the operand is shared by all batch elements (broadcast)
    int64_t c_batch = 0;                   // element stride of C between batch elements
};

// SPLIT (3xTF32, fp32-accurate): a stage also holds the low-order tf32 halves of both operands.
template <int BN, bool SPLIT = false>
struct GemmCfg {
    static constexpr int kHalfBytes = (kBM + BN) * kBK * 4;
    static constexpr int kStageBytes = kHalfBytes * (SPLIT ? 2 : 1);
    static constexpr int kStages = (200 * 1024) / kStageBytes > 8 ? 8 : (200 * 1024) / kStageBytes;
    static constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;  // two accumulator stages, power of two >= 32
    static constexpr int kEpiBytes = 4 * 32 * 33 * 4;  // per-epilogue-warp transpose staging
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/ + kEpiBytes;
};

// A_MN / B_MN: the operand is MN-major (its M resp. N axis is the contiguous one) and is consumed in place.
// SPLIT: every operand arrives as two tf32-exact arrays, hi = tf32(x) and lo = tf32(x - hi) (split_tf32_kernel);
// each k-step issues A_lo·B_hi + A_hi·B_lo + A_hi·B_hi into the same fp32 accumulator — the 3xTF32 scheme, whose
// result carries ~2^-21 relative input error instead of tf32's 2^-11 (the dropped A_lo·B_lo term is ~2^-22).
template <int BN, bool A_MN, bool B_MN, bool SPLIT = false>
__global__ void __launch_bounds__(kGemmThreads, 1) gemm_tf32_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                  const __grid_constant__ CUtensorMap map_b,
                                                                  const __grid_constant__ CUtensorMap map_a_lo,
                                                                  const __grid_constant__ CUtensorMap map_b_lo,
                                                                  const GemmParams p) {
    static_assert(!SPLIT || (!A_MN && !B_MN), "split operands are always K-major scratch");
    using Cfg = GemmCfg<BN, SPLIT>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + Cfg::kStages * Cfg::kStageBytes);
    uint64_t *full = bars, *empty = bars + Cfg::kStages;
    uint64_t *tmem_full = bars + 2 * Cfg::kStages, *tmem_empty = tmem_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);
    float *epi = reinterpret_cast<float *>(smem + Cfg::kStages * Cfg::kStageBytes + 256);  // 4 warps x 32 x 33 floats

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = p.tiles_m * p.tiles_n;
    const int num_kb = (p.K + kBK - 1) / kBK;

    if (warp == 1 && elect_one()) {
        for (int s = 0; s < Cfg::kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)Cfg::kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m_blk = tile % p.tiles_m, n_blk = tile / p.tiles_m;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t *sa = smem + stage * Cfg::kStageBytes;
                    uint8_t *sb = sa + kBM * kBK * 4;
                    mbar_arrive_expect_tx(&full[stage], Cfg::kStageBytes);
                    if constexpr (A_MN) {
#pragma unroll
                        for (int blk = 0; blk < kBM / 32; ++blk)  // box {32 M, 32 K}: coordinates (m, k)
                            tma_load_2d(sa + blk * 4096, &map_a, &full[stage], m_blk * kBM + blk * 32, kb * kBK);
                    } else {
                        tma_load_2d(sa, &map_a, &full[stage], kb * kBK, m_blk * kBM);
                    }
                    if constexpr (B_MN) {
#pragma unroll
                        for (int blk = 0; blk < BN / 32; ++blk)
                            tma_load_2d(sb + blk * 4096, &map_b, &full[stage], n_blk * BN + blk * 32, kb * kBK);
                    } else {
                        tma_load_2d(sb, &map_b, &full[stage], kb * kBK, n_blk * BN);
                    }
                    if constexpr (SPLIT) {
                        tma_load_2d(sa + Cfg::kHalfBytes, &map_a_lo, &full[stage], kb * kBK, m_blk * kBM);
                        tma_load_2d(sb + Cfg::kHalfBytes, &map_b_lo, &full[stage], kb * kBK, n_blk * BN);
                    }
                    if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) @4, a/b_format TF32 (2) @7/@10,
            // a/b K-major (0) @15/@16, N>>3 @17, M>>4 @24
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                                   ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tcgen05_fence_after();
                const uint32_t tmem_c = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
                    const uint32_t sb = sa + kBM * kBK * 4;
                    const uint64_t da = A_MN ? make_mnmajor_sw128_desc(sa) : make_kmajor_sw128_desc(sa);
                    const uint64_t db = B_MN ? make_mnmajor_sw128_desc(sb) : make_kmajor_sw128_desc(sb);
                    // UMMA K = 8 tf32: K-major advances 32 bytes inside the swizzle row, MN-major one 8-row group
                    constexpr uint64_t stepA = A_MN ? (1024 >> 4) : (32 >> 4), stepB = B_MN ? (1024 >> 4) : (32 >> 4);
                    if constexpr (SPLIT) {
                        const uint64_t da_lo = make_kmajor_sw128_desc(sa + Cfg::kHalfBytes);
                        const uint64_t db_lo = make_kmajor_sw128_desc(sb + Cfg::kHalfBytes);
#pragma unroll
                        for (int k = 0; k < kBK / 8; ++k) {  // small terms first
                            umma_tf32(tmem_c, da_lo + (uint64_t)k * stepA, db + (uint64_t)k * stepB, idesc, (kb | k) ? 1u : 0u);
                            umma_tf32(tmem_c, da + (uint64_t)k * stepA, db_lo + (uint64_t)k * stepB, idesc, 1u);
                            umma_tf32(tmem_c, da + (uint64_t)k * stepA, db + (uint64_t)k * stepB, idesc, 1u);
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < kBK / 8; ++k)
                            umma_tf32(tmem_c, da + (uint64_t)k * stepA, db + (uint64_t)k * stepB, idesc, (kb | k) ? 1u : 0u);
                    }
                    umma_commit(&empty[stage]);
                    if (kb == num_kb - 1) umma_commit(&tmem_full[acc]);
                    if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // epilogue warps: TMEM lane quarter = warp % 4
        const int quarter = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        const bool vec_ok = p.ldc_n == 1 && (p.ldc_m % 4) == 0 && (reinterpret_cast<uintptr_t>(p.c) & 15) == 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m_blk = tile % p.tiles_m, n_blk = tile / p.tiles_m;
            mbar_wait(&tmem_full[acc], acc_phase);
            tcgen05_fence_after();
            const int row = m_blk * kBM + quarter * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                const int col = n_blk * BN + c0;
                if (col >= p.N) break;  // warp-uniform
                uint32_t r[32];
                tmem_ld_32x32(taddr + (uint32_t)c0, r);
                if (vec_ok && col + 32 <= p.N) {
                    // Row-major C: transpose the warp's 32x32 block through shared memory so that every store
                    // instruction writes 4 rows x 128 contiguous bytes (lane -> row lane/8, columns 4*(lane%8)..+3)
                    // instead of 32 rows x 16 bytes.
                    float *blk = epi + (warp - 2) * (32 * 33);
#pragma unroll
                    for (int j = 0; j < 32; ++j) blk[lane * 33 + j] = __uint_as_float(r[j]);
                    __syncwarp();
                    const int cg = (lane & 7) * 4;
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int rr = it * 4 + (lane >> 3);
                        const int grow = m_blk * kBM + quarter * 32 + rr;
                        if (grow < p.M) {
                            const float *sp = blk + rr * 33 + cg;
                            *reinterpret_cast<float4 *>(p.c + (int64_t)grow * p.ldc_m + col + cg) = make_float4(sp[0], sp[1], sp[2], sp[3]);
                        }
                    }
                    __syncwarp();
                } else if (row < p.M) {
                    float *crow = p.c + (int64_t)row * p.ldc_m + (int64_t)col * p.ldc_n;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (col + j < p.N) crow[(int64_t)j * p.ldc_n] = __uint_as_float(r[j]);
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::kTmemCols));
    }
}

// ---------------------------------------------------------------------------------------------------------------
// cta_group::2: a CTA PAIR (cluster of 2, two SMs of one TPC) computes one 256 x 256 tile with M = 256 MMAs.
// Each CTA stages its own 128 rows of A and HALF of B (128 of the 256 N rows) — 32 KiB per stage instead of the
// 48 KiB a single CTA needs for a 128 x 256 tile — and the tensor core reads both halves of B out of both SMs'
// shared memory: per SM the operand traffic through shared memory is 2/3 of the one-CTA kernel's per flop, which is
// what caps that kernel at ~84 % of cuBLAS. Accumulators: each CTA's TMEM holds its 128 rows x 256
// columns (two stages = all 512 columns).
// Protocol (PTX forms as in CUTLASS cute/arch/copy_sm100_tma.hpp, mma_sm100_umma.hpp, cutlass/arch/barrier.h):
//   - both CTAs issue their TMA loads with .cta_group::2, completing transaction bytes on the LEADER's (rank 0) full
//     barrier (address with the peer bit cleared); the leader alone arms it with expect_tx for both CTAs' bytes;
//   - the leader's MMA thread issues tcgen05.mma.cta_group::2; tcgen05.commit.cta_group::2 ... multicast::cluster
//     arrives on the empty / tmem_full barriers of BOTH CTAs;
//   - the epilogue warps of both CTAs arrive on the LEADER's tmem_empty barrier (8 arrivals).
// ---------------------------------------------------------------------------------------------------------------
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // cute::Sm100MmaPeerBitMask

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void *dst, const CUtensorMap *map, uint64_t *leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(leader_bar) & kPeerBitMask), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(void *dst, const CUtensorMap *map, uint64_t *leader_bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(leader_bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_c), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t *bar) {  // arrives on `bar` in both CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t *bar) {  // from either CTA, on the leader's barrier
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}

constexpr int kBN2 = 256;  // tile N of the pair; each CTA stages kBN2 / 2 rows of B

// Tile order: groups of kGroupM tile rows are walked column by column, so that the ~74 tiles in flight cover a
// compact (8 x ~9) patch of C and share their A and B panels in L2 (row-major order makes every wave read ALL of A).
constexpr int kGroupM = 8;
__device__ __forceinline__ void tile_coords(int tile, int tiles_m, int tiles_n, int &m_blk, int &n_blk) {
    const int per_group = kGroupM * tiles_n;
    const int group = tile / per_group, in_group = tile - group * per_group;
    const int m_first = group * kGroupM;
    const int rows = tiles_m - m_first < kGroupM ? tiles_m - m_first : kGroupM;
    n_blk = in_group / rows;
    m_blk = m_first + (in_group - n_blk * rows);
}

template <bool SPLIT = false>
struct GemmCfg2 {
    static constexpr int kHalfBytes = (kBM + kBN2 / 2) * kBK * 4;   // 32 KiB per CTA: its rows of A + its half of B
    static constexpr int kStageBytes = kHalfBytes * (SPLIT ? 2 : 1);  // 3xTF32: plus the low-order halves
    static constexpr int kStages = SPLIT ? 3 : 6;
    static constexpr int kTmemCols = 2 * kBN2;
    static constexpr int kEpiBytes = 4 * 32 * 33 * 4;
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256 + kEpiBytes;
};

template <bool A_MN, bool B_MN, bool SPLIT = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
    gemm_tf32_2cta_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                          const __grid_constant__ CUtensorMap map_a_lo, const __grid_constant__ CUtensorMap map_b_lo,
                          const GemmParams p) {
    static_assert(!SPLIT || (!A_MN && !B_MN), "split operands are always K-major scratch");
    using Cfg = GemmCfg2<SPLIT>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + Cfg::kStages * Cfg::kStageBytes);
    uint64_t *full = bars, *empty = bars + Cfg::kStages;
    uint64_t *tmem_full = bars + 2 * Cfg::kStages, *tmem_empty = tmem_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);
    float *epi = reinterpret_cast<float *>(smem + Cfg::kStages * Cfg::kStageBytes + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int tiles_per_mat = p.tiles_m * p.tiles_n;  // tiles of 256 x 256
    const int num_tiles = tiles_per_mat * p.nbatch;   // the maps are 3-D: (k, row, batch element)
    const int num_kb = (p.K + kBK - 1) / kBK;

    if (warp == 1 && elect_one()) {
        for (int s = 0; s < Cfg::kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 8);  // 4 epilogue warps of each CTA
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)Cfg::kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tcgen05_fence_before();
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = pair; tile < num_tiles; tile += npairs) {
                int m_blk, n_blk;
                const int bi = tile / tiles_per_mat;
                tile_coords(tile - bi * tiles_per_mat, p.tiles_m, p.tiles_n, m_blk, n_blk);
                const int ba = p.a_batched ? bi : 0, bb = p.b_batched ? bi : 0;
                const int m0 = m_blk * 2 * kBM + (int)rank * kBM, n0 = n_blk * kBN2 + (int)rank * (kBN2 / 2);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t *sa = smem + stage * Cfg::kStageBytes;
                    uint8_t *sb = sa + kBM * kBK * 4;
                    if (leader) mbar_arrive_expect_tx(&full[stage], 2 * Cfg::kStageBytes);
                    if constexpr (A_MN) {
#pragma unroll
                        for (int blk = 0; blk < kBM / 32; ++blk)
                            tma_load_3d_2sm(sa + blk * 4096, &map_a, &full[stage], m0 + blk * 32, kb * kBK, ba);
                    } else {
                        tma_load_3d_2sm(sa, &map_a, &full[stage], kb * kBK, m0, ba);
                    }
                    if constexpr (B_MN) {
#pragma unroll
                        for (int blk = 0; blk < kBN2 / 2 / 32; ++blk)
                            tma_load_3d_2sm(sb + blk * 4096, &map_b, &full[stage], n0 + blk * 32, kb * kBK, bb);
                    } else {
                        tma_load_3d_2sm(sb, &map_b, &full[stage], kb * kBK, n0, bb);
                    }
                    if constexpr (SPLIT) {
                        tma_load_3d_2sm(sa + Cfg::kHalfBytes, &map_a_lo, &full[stage], kb * kBK, m0, ba);
                        tma_load_3d_2sm(sb + Cfg::kHalfBytes, &map_b_lo, &full[stage], kb * kBK, n0, bb);
                    }
                    if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (leader && elect_one()) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                                   ((uint32_t)(kBN2 >> 3) << 17) | ((uint32_t)((2 * kBM) >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = pair; tile < num_tiles; tile += npairs) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tcgen05_fence_after();
                const uint32_t tmem_c = tmem_base + (uint32_t)(acc * kBN2);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
                    const uint32_t sb = sa + kBM * kBK * 4;
                    const uint64_t da = A_MN ? make_mnmajor_sw128_desc(sa) : make_kmajor_sw128_desc(sa);
                    const uint64_t db = B_MN ? make_mnmajor_sw128_desc(sb) : make_kmajor_sw128_desc(sb);
                    constexpr uint64_t stepA = A_MN ? (1024 >> 4) : (32 >> 4), stepB = B_MN ? (1024 >> 4) : (32 >> 4);
                    if constexpr (SPLIT) {  // 3xTF32: small terms first
                        const uint64_t da_lo = make_kmajor_sw128_desc(sa + Cfg::kHalfBytes);
                        const uint64_t db_lo = make_kmajor_sw128_desc(sb + Cfg::kHalfBytes);
#pragma unroll
                        for (int k = 0; k < kBK / 8; ++k) {
                            umma_tf32_2sm(tmem_c, da_lo + (uint64_t)k * stepA, db + (uint64_t)k * stepB, idesc, (kb | k) ? 1u : 0u);
                            umma_tf32_2sm(tmem_c, da + (uint64_t)k * stepA, db_lo + (uint64_t)k * stepB, idesc, 1u);
                            umma_tf32_2sm(tmem_c, da + (uint64_t)k * stepA, db + (uint64_t)k * stepB, idesc, 1u);
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < kBK / 8; ++k)
                            umma_tf32_2sm(tmem_c, da + (uint64_t)k * stepA, db + (uint64_t)k * stepB, idesc, (kb | k) ? 1u : 0u);
                    }
                    umma_commit_2sm(&empty[stage]);
                    if (kb == num_kb - 1) umma_commit_2sm(&tmem_full[acc]);
                    if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        const int quarter = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        const bool vec_ok = p.ldc_n == 1 && (p.ldc_m % 4) == 0 && (p.c_batch % 4) == 0 && (reinterpret_cast<uintptr_t>(p.c) & 15) == 0;
        for (int tile = pair; tile < num_tiles; tile += npairs) {
            int m_blk, n_blk;
            const int bi = tile / tiles_per_mat;
            tile_coords(tile - bi * tiles_per_mat, p.tiles_m, p.tiles_n, m_blk, n_blk);
            float *cmat = p.c + (int64_t)bi * p.c_batch;
            const int mbase = m_blk * 2 * kBM + (int)rank * kBM;
            mbar_wait(&tmem_full[acc], acc_phase);
            tcgen05_fence_after();
            const int row = mbase + quarter * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * kBN2);
#pragma unroll 1
            for (int c0 = 0; c0 < kBN2; c0 += 32) {
                const int col = n_blk * kBN2 + c0;
                if (col >= p.N) break;  // warp-uniform
                uint32_t r[32];
                tmem_ld_32x32(taddr + (uint32_t)c0, r);
                if (vec_ok && col + 32 <= p.N) {
                    float *blk = epi + (warp - 2) * (32 * 33);
#pragma unroll
                    for (int j = 0; j < 32; ++j) blk[lane * 33 + j] = __uint_as_float(r[j]);
                    __syncwarp();
                    const int cg = (lane & 7) * 4;
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int rr = it * 4 + (lane >> 3);
                        const int grow = mbase + quarter * 32 + rr;
                        if (grow < p.M) {
                            const float *sp = blk + rr * 33 + cg;
                            *reinterpret_cast<float4 *>(cmat + (int64_t)grow * p.ldc_m + col + cg) = make_float4(sp[0], sp[1], sp[2], sp[3]);
                        }
                    }
                    __syncwarp();
                } else if (row < p.M) {
                    float *crow = cmat + (int64_t)row * p.ldc_m + (int64_t)col * p.ldc_n;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (col + j < p.N) crow[(int64_t)j * p.ldc_n] = __uint_as_float(r[j]);
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tcgen05_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::kTmemCols));
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Host side of the tf32 path
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// 2-D fp32 tensor map: `inner` contiguous elements per line, `outer` lines `pitch` elements apart;
// box = box_outer lines x 32 elements (128 bytes, one swizzle row).
// K-major operand: inner = K, outer = rows (M or N), box_outer = tile rows.
// MN-major operand: inner = rows (M or N), outer = K, box_outer = 32 k-rows.
dn_status make_map(CUtensorMap *map, const float *base, int64_t outer, int64_t inner, int64_t pitch, int box_outer,
                   bool mn_major = false) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return set_error(DN_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)pitch * 4};
    cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(DN_ERR_CUDA, "cuTensorMapEncodeTiled failed with code %d", (int)r);
    return DN_OK;
}

// 3-D fp32 tensor map for the batched CTA-pair kernel: (inner, outer, batch); box = {32, box_outer, 1}. A shared
// (broadcast) operand has nbatch = 1 and is always addressed with batch coordinate 0.
dn_status make_map3(CUtensorMap *map, const float *base, int64_t outer, int64_t inner, int64_t pitch, int box_outer,
                    bool mn_major, int64_t nbatch, int64_t batch_stride) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return set_error(DN_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)outer, (cuuint64_t)(nbatch > 0 ? nbatch : 1)};
    const int64_t bs = nbatch > 1 ? batch_stride : pitch * outer;  // unused when there is one batch element
    cuuint64_t strides[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)bs * 4};
    cuuint32_t box[3] = {(cuuint32_t)kBK, (cuuint32_t)box_outer, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(DN_ERR_CUDA, "cuTensorMapEncodeTiled (3-D) failed with code %d", (int)r);
    return DN_OK;
}

// A 2-D operand view [rows, K] (element strides rs, ks).
struct Operand2D {
    const float *ptr;
    int64_t rows, K, rs, ks;
};

bool tma_ready(const Operand2D &o) {  // K-major in place
    return o.ks == 1 && o.rs >= o.K && (o.rs % 4) == 0 && (reinterpret_cast<uintptr_t>(o.ptr) & 15) == 0;
}
bool tma_ready_mn(const Operand2D &o) {  // MN-major in place: the row index is the contiguous one
    return o.rs == 1 && o.ks >= o.rows && (o.ks % 4) == 0 && (reinterpret_cast<uintptr_t>(o.ptr) & 15) == 0;
}

// Repack into a fresh K-major buffer with a 16-byte aligned pitch (stream-ordered scratch).
dn_status repack_kmajor(Operand2D &o, void **scratch) {
    const int64_t pitch = (o.K + 3) / 4 * 4;
    dn_status st = scratch_alloc((size_t)o.rows * pitch * 4, scratch);
    if (st != DN_OK) return st;
    dn_tensor src{}, dst{};
    src.base = const_cast<float *>(o.ptr);
    src.offset = 0;
    src.ndims = 2;
    src.dtype = DN_F32;
    src.shape[0] = o.rows; src.shape[1] = o.K;
    src.stride[0] = o.rs; src.stride[1] = o.ks;
    dst = src;
    dst.base = *scratch;
    dst.stride[0] = pitch; dst.stride[1] = 1;
    st = dn_copy(&dst, &src);
    if (st != DN_OK) return st;
    o.ptr = static_cast<const float *>(*scratch);
    o.rs = pitch;
    o.ks = 1;
    return DN_OK;
}

template <int BN, bool A_MN, bool B_MN, bool SPLIT = false>
dn_status launch_tf32(const CUtensorMap &ma, const CUtensorMap &mb, const CUtensorMap &ma_lo, const CUtensorMap &mb_lo,
                      const GemmParams &p) {
    using Cfg = GemmCfg<BN, SPLIT>;
    // the opt-in to > 48 KB of dynamic shared memory is a per-DEVICE function attribute: one flag per device
    // (a process may drive several devices, dn_set_device / dn_shard_*), set at most once each
    static std::atomic<bool> configured[64];
    int dev = 0;
    DN_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
        DN_CUDA_TRY(cudaFuncSetAttribute(gemm_tf32_kernel<BN, A_MN, B_MN, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes));
        if (dev >= 0 && dev < 64) configured[dev].store(true, std::memory_order_release);
    }
    const int tiles = p.tiles_m * p.tiles_n;
    const int grid = tiles < sm_count() ? tiles : sm_count();
    DN_LAUNCH((gemm_tf32_kernel<BN, A_MN, B_MN, SPLIT>), grid, kGemmThreads, Cfg::kSmemBytes, ma, mb, ma_lo, mb_lo, p);
    return launch_status("tcgen05 GEMM kernel");
}

template <bool A_MN, bool B_MN, bool SPLIT = false>
dn_status launch_tf32_2cta(const CUtensorMap &ma, const CUtensorMap &mb, const CUtensorMap &ma_lo, const CUtensorMap &mb_lo,
                           const GemmParams &p) {
    using Cfg = GemmCfg2<SPLIT>;
    static std::atomic<bool> configured[64];
    int dev = 0;
    DN_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
        DN_CUDA_TRY(cudaFuncSetAttribute(gemm_tf32_2cta_kernel<A_MN, B_MN, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes));
        if (dev >= 0 && dev < 64) configured[dev].store(true, std::memory_order_release);
    }
    const int64_t tiles = (int64_t)p.tiles_m * p.tiles_n * p.nbatch;
    const int pairs = tiles < sm_count() / 2 ? (int)tiles : sm_count() / 2;
    DN_LAUNCH((gemm_tf32_2cta_kernel<A_MN, B_MN, SPLIT>), 2 * pairs, kGemmThreads, Cfg::kSmemBytes, ma, mb, ma_lo, mb_lo, p);
    return launch_status("tcgen05 2-CTA GEMM kernel");
}

template <int BN>
dn_status launch_tf32_major(bool a_mn, bool b_mn, const CUtensorMap &ma, const CUtensorMap &mb, const GemmParams &p) {
    if (a_mn) return b_mn ? launch_tf32<BN, true, true>(ma, mb, ma, mb, p) : launch_tf32<BN, true, false>(ma, mb, ma, mb, p);
    return b_mn ? launch_tf32<BN, false, true>(ma, mb, ma, mb, p) : launch_tf32<BN, false, false>(ma, mb, ma, mb, p);
}

// C[M,N] (strides cm, cn) = A[M,K] (am, ak) · B[K,N] (bk, bn), fp32 operands read as tf32 (DN_MATH_TF32).
// large problems: CTA pairs on 256 x 256 tiles (DN_GEMM_2CTA=0, a test hook, keeps the one-CTA kernel)
bool tf32_pair_eligible(int64_t M, int64_t N) {
    static const bool allow = [] { const char *e = getenv("DN_GEMM_2CTA"); return !(e && e[0] == '0'); }();
    return allow && M >= 512 && N >= 256;
}

// nbatch > 1 (CTA-pair kernel only): ONE launch for all batch elements; *_bs are the element strides between batch
// elements, 0 = the operand is shared (broadcast batch dims).
dn_status gemm_f32_tf32(float *c, int64_t cm, int64_t cn, const float *a, int64_t am, int64_t ak, const float *b, int64_t bk,
                   int64_t bn, int64_t M, int64_t N, int64_t K, int64_t nbatch = 1, int64_t c_bs = 0, int64_t a_bs = 0,
                   int64_t b_bs = 0) {
    if (M == 0 || N == 0) return DN_OK;
    if (M >= (1ll << 31) || N >= (1ll << 31) || K >= (1ll << 31)) return set_error(DN_ERR_UNSUPPORTED, "MatMatDot: extent exceeds 2^31-1");
    if (K == 0) {  // empty sum: C = 0
        dn_tensor t{};
        t.base = c; t.ndims = 2; t.dtype = DN_F32;
        t.shape[0] = M; t.shape[1] = N; t.stride[0] = cm; t.stride[1] = cn;
        const float zero = 0.f;
        return dn_fill_const(&t, &zero);
    }
    Operand2D A{a, M, K, am, ak}, B{b, N, K, bn, bk};  // B viewed as [N, K]
    void *sa = nullptr, *sb = nullptr;
    dn_status st = DN_OK;
    // operand classes: K-major in place, MN-major in place, or repacked K-major
    const bool a_mn = !tma_ready(A) && tma_ready_mn(A), b_mn = !tma_ready(B) && tma_ready_mn(B);
    if (!a_mn && !tma_ready(A)) st = repack_kmajor(A, &sa);
    if (st == DN_OK && !b_mn && !tma_ready(B)) st = repack_kmajor(B, &sb);
    CUtensorMap ma, mb;
    GemmParams p;
    p.c = c; p.ldc_m = cm; p.ldc_n = cn;
    p.M = (int32_t)M; p.N = (int32_t)N; p.K = (int32_t)K;
    const int BN = N <= 32 ? 32 : (N <= 128 ? 128 : 256);
    const bool two_cta = tf32_pair_eligible(M, N);
    if (nbatch > 1 && (!two_cta || sa || sb))
        return set_error(DN_ERR_INVALID_ARG, "internal: batched tf32 launch needs the CTA-pair kernel and in-place operands");
    if (st == DN_OK && two_cta) {
        p.tiles_m = (int32_t)((M + 2 * kBM - 1) / (2 * kBM));
        p.tiles_n = (int32_t)((N + kBN2 - 1) / kBN2);
        p.nbatch = (int32_t)nbatch;
        p.a_batched = a_bs != 0;
        p.b_batched = b_bs != 0;
        p.c_batch = c_bs;
        const int64_t na = a_bs != 0 ? nbatch : 1, nb = b_bs != 0 ? nbatch : 1;
        st = a_mn ? make_map3(&ma, A.ptr, K, A.rows, A.ks, kBK, true, na, a_bs) : make_map3(&ma, A.ptr, A.rows, K, A.rs, kBM, false, na, a_bs);
        if (st == DN_OK)
            st = b_mn ? make_map3(&mb, B.ptr, K, B.rows, B.ks, kBK, true, nb, b_bs)
                      : make_map3(&mb, B.ptr, B.rows, K, B.rs, kBN2 / 2, false, nb, b_bs);
        if (st == DN_OK) {
            if (a_mn) st = b_mn ? launch_tf32_2cta<true, true>(ma, mb, ma, mb, p) : launch_tf32_2cta<true, false>(ma, mb, ma, mb, p);
            else st = b_mn ? launch_tf32_2cta<false, true>(ma, mb, ma, mb, p) : launch_tf32_2cta<false, false>(ma, mb, ma, mb, p);
        }
        scratch_free(sa);
        scratch_free(sb);
        return st;
    }
    p.tiles_m = (int32_t)((M + kBM - 1) / kBM);
    p.tiles_n = (int32_t)((N + BN - 1) / BN);
    if (st == DN_OK) st = a_mn ? make_map(&ma, A.ptr, K, A.rows, A.ks, kBK, true) : make_map(&ma, A.ptr, A.rows, K, A.rs, kBM);
    if (st == DN_OK) st = b_mn ? make_map(&mb, B.ptr, K, B.rows, B.ks, kBK, true) : make_map(&mb, B.ptr, B.rows, K, B.rs, BN);
    if (st == DN_OK) {
        if (BN == 32) st = launch_tf32_major<32>(a_mn, b_mn, ma, mb, p);
        else if (BN == 128) st = launch_tf32_major<128>(a_mn, b_mn, ma, mb, p);
        else st = launch_tf32_major<256>(a_mn, b_mn, ma, mb, p);
    }
    scratch_free(sa);
    scratch_free(sb);
    return st;
}

// ---------------------------------------------------------------------------------------------------------------
// float64: shared-memory tiled SIMT kernel (64 x 64 tile, 4 x 4 per thread)
// ---------------------------------------------------------------------------------------------------------------
constexpr int kDT = 64, kDK = 16;

// Batch decomposition for the SIMT kernel: blockIdx.z (+ z0) -> element offsets of the three operands.
struct SimtBatch {
    int32_t ndims;
    uint32_t shape[DN_MAX_DIMS];   // innermost batch dim first
    FastDiv div[DN_MAX_DIMS];
    int64_t cs[DN_MAX_DIMS], as[DN_MAX_DIMS], bs[DN_MAX_DIMS];  // elements; 0 = broadcast
    uint32_t z0;
};

// Shared-memory tiled SIMT GEMM (64 x 64 tile, 4 x 4 per thread), one launch for a whole batch. float64 always
// runs here (B200 tensor cores have no fp64 MMA worth the name); float32 runs here for batches of SMALL matrices,
// where a tcgen05 launch + two tensor maps per batch element (the reference: cuBLAS gemmBatched over pointer
// arrays, CudaBackend.fs:429-449) would be launch-bound — and the result is exact fp32 rather than tf32.
template <class T>
__global__ void __launch_bounds__(256) gemm_simt_kernel(T *c, int64_t cm, int64_t cn, const T *a, int64_t am, int64_t ak,
                                                       const T *b, int64_t bk, int64_t bn, int M, int N, int K,
                                                       const __grid_constant__ SimtBatch batch, const int *guard) {
    if (guard && *guard == 0) return;  // conditional re-run (3xTF32 with non-finite inputs): nothing to do
    __shared__ T sa[kDK][kDT + 1], sb[kDK][kDT + 1];
    {
        uint32_t rem = blockIdx.z + batch.z0;
#pragma unroll
        for (int d = 0; d < DN_MAX_DIMS; ++d) {
            if (d >= batch.ndims) break;
            const uint32_t q = batch.div[d].div(rem);
            const int64_t x = rem - q * batch.shape[d];
            c += x * batch.cs[d];
            a += x * batch.as[d];
            b += x * batch.bs[d];
            rem = q;
        }
    }
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.y * kDT, n0 = blockIdx.x * kDT;
    T acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += kDK) {
        for (int i = threadIdx.x; i < kDT * kDK; i += 256) {
            // choose the faster-varying index to follow the operand's contiguous axis
            int mm, kk;
            if (ak == 1) { kk = i % kDK; mm = i / kDK; } else { mm = i % kDT; kk = i / kDT; }
            const int gm = m0 + mm, gk = k0 + kk;
            sa[kk][mm] = (gm < M && gk < K) ? a[(int64_t)gm * am + (int64_t)gk * ak] : T(0);
            int nn, kb;
            if (bk == 1) { kb = i % kDK; nn = i / kDK; } else { nn = i % kDT; kb = i / kDT; }
            const int gn = n0 + nn, gk2 = k0 + kb;
            sb[kb][nn] = (gn < N && gk2 < K) ? b[(int64_t)gk2 * bk + (int64_t)gn * bn] : T(0);
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kDK; ++kk) {
            T av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { av[i] = sa[kk][ty * 4 + i]; bv[i] = sb[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gm = m0 + ty * 4 + i, gn = n0 + tx * 4 + j;
            if (gm < M && gn < N) c[(int64_t)gm * cm + (int64_t)gn * cn] = acc[i][j];
        }
}

// One launch (per 65535 batch elements) for the whole batch. `bt`/`ba`/`bb`: batch strides, innermost batch dim first.
template <class T>
dn_status gemm_simt(T *c, int64_t cm, int64_t cn, const T *a, int64_t am, int64_t ak, const T *b, int64_t bk, int64_t bn,
                    int64_t M, int64_t N, int64_t K, int nbd, const int64_t *bshape, const int64_t *bt, const int64_t *ba,
                    const int64_t *bb, const int *guard = nullptr) {
    if (M == 0 || N == 0) return DN_OK;
    if (M >= (1ll << 31) || N >= (1ll << 31) || K >= (1ll << 31)) return set_error(DN_ERR_UNSUPPORTED, "MatMatDot: extent exceeds 2^31-1");
    SimtBatch batch;
    batch.ndims = nbd;
    batch.z0 = 0;
    int64_t nbatch = 1;
    for (int d = 0; d < DN_MAX_DIMS; ++d) {
        const bool on = d < nbd;
        batch.shape[d] = on ? (uint32_t)bshape[d] : 1;
        batch.div[d].init(batch.shape[d]);
        batch.cs[d] = on ? bt[d] : 0;
        batch.as[d] = on ? ba[d] : 0;
        batch.bs[d] = on ? bb[d] : 0;
        if (on) nbatch *= bshape[d];
    }
    if (nbatch == 0) return DN_OK;
    if (nbatch >= (1ll << 31)) return set_error(DN_ERR_UNSUPPORTED, "BatchedMatMatDot: more than 2^31-1 matrices");
    dim3 grid((unsigned)((N + kDT - 1) / kDT), (unsigned)((M + kDT - 1) / kDT), 1);
    if (grid.y > 65535) return set_error(DN_ERR_UNSUPPORTED, "MatMatDot (SIMT path): M too large");
    for (int64_t z0 = 0; z0 < nbatch; z0 += 65535) {
        grid.z = (unsigned)(nbatch - z0 < 65535 ? nbatch - z0 : 65535);
        batch.z0 = (uint32_t)z0;
        DN_LAUNCH((gemm_simt_kernel<T>), grid, 256, 0, c, cm, cn, a, am, ak, b, bk, bn, (int)M, (int)N, (int)K, batch, guard);
    }
    return launch_status("SIMT GEMM kernel");
}

// ---------------------------------------------------------------------------------------------------------------
// float32 at fp32 accuracy (DN_MATH_FP32, the default): 3xTF32 on the tensor cores
// ---------------------------------------------------------------------------------------------------------------
std::atomic<int> g_math_mode{DN_MATH_FP32};

__device__ __forceinline__ float to_tf32(float x) {  // round to nearest, ties away: the result is a tf32-exact fp32 word
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

// hi = tf32(x), lo = tf32(x - hi) for a K-major [rows, K] operand (row pitch `rs` elements) into dense [rows, pitch]
// arrays. Both outputs are exactly representable in tf32, so whatever the tensor core does with the low 13 bits of
// its inputs (it ignores them) does not matter. Non-finite x: hi = x, lo = 0, and *nonfinite is raised: an infinity
// times the OTHER operand's low half can come out with the wrong sign (inf - inf = NaN where the true product is
// +-inf), so such a product is recomputed by the exact kernel (gemm_f32_split).
__global__ void __launch_bounds__(256) split_tf32_kernel(const float *src, int64_t rs, int64_t rows, int64_t K, float *hi,
                                                        float *lo, int64_t pitch, int *nonfinite) {
    const int64_t n4 = pitch / 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows * n4; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / n4, k = (i - r * n4) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const float *row = src + r * rs;
        if (k + 4 <= K) v = *reinterpret_cast<const float4 *>(row + k);
        else {
            if (k < K) v.x = row[k];
            if (k + 1 < K) v.y = row[k + 1];
            if (k + 2 < K) v.z = row[k + 2];
        }
        float4 h, l;
        h.x = to_tf32(v.x); h.y = to_tf32(v.y); h.z = to_tf32(v.z); h.w = to_tf32(v.w);
        l.x = isfinite(v.x) ? to_tf32(v.x - h.x) : 0.f;
        l.y = isfinite(v.y) ? to_tf32(v.y - h.y) : 0.f;
        l.z = isfinite(v.z) ? to_tf32(v.z - h.z) : 0.f;
        l.w = isfinite(v.w) ? to_tf32(v.w - h.w) : 0.f;
        if (!(isfinite(v.x) && isfinite(v.y) && isfinite(v.z) && isfinite(v.w))) *nonfinite = 1;
        *reinterpret_cast<float4 *>(hi + r * pitch + k) = h;
        *reinterpret_cast<float4 *>(lo + r * pitch + k) = l;
    }
}

// Makes `o` K-major (in place when it already is, else through the transposing copy) and splits it; on return
// hi / lo are dense [rows, pitch] scratch arrays.
dn_status split_operand(Operand2D o, float **hi, float **lo, int64_t *pitch, void **scratch_hi, void **scratch_lo, int *nonfinite) {
    void *packed = nullptr;
    dn_status st = DN_OK;
    if (!tma_ready(o)) st = repack_kmajor(o, &packed);
    *pitch = (o.K + 3) / 4 * 4;
    if (st == DN_OK) st = scratch_alloc((size_t)o.rows * *pitch * 4, scratch_hi);
    if (st == DN_OK) st = scratch_alloc((size_t)o.rows * *pitch * 4, scratch_lo);
    if (st == DN_OK) {
        *hi = static_cast<float *>(*scratch_hi);
        *lo = static_cast<float *>(*scratch_lo);
        int64_t ctas = (o.rows * (*pitch / 4) + 255) / 256;
        const int64_t cap = (int64_t)sm_count() * 8;
        if (ctas > cap) ctas = cap;
        DN_LAUNCH(split_tf32_kernel, (unsigned)ctas, 256, 0, o.ptr, o.rs, o.rows, o.K, *hi, *lo, *pitch, nonfinite);
        st = launch_status("tf32 split kernel");
    }
    scratch_free(packed);
    return st;
}

dn_status gemm_f32_split(float *c, int64_t cm, int64_t cn, const float *a, int64_t am, int64_t ak, const float *b, int64_t bk,
                         int64_t bn, int64_t M, int64_t N, int64_t K) {
    Operand2D A{a, M, K, am, ak}, B{b, N, K, bn, bk};  // B viewed as [N, K]
    float *ahi = nullptr, *alo = nullptr, *bhi = nullptr, *blo = nullptr;
    int64_t pa = 0, pb = 0;
    void *s0 = nullptr, *s1 = nullptr, *s2 = nullptr, *s3 = nullptr, *s_flag = nullptr;
    dn_status st = scratch_alloc(sizeof(int), &s_flag);
    if (st != DN_OK) return st;
    int *nonfinite = static_cast<int *>(s_flag);
    cudaMemsetAsync(nonfinite, 0, sizeof(int), current_stream());
    st = split_operand(A, &ahi, &alo, &pa, &s0, &s1, nonfinite);
    if (st == DN_OK) st = split_operand(B, &bhi, &blo, &pb, &s2, &s3, nonfinite);
    CUtensorMap ma, mb, mal, mbl;
    GemmParams p;
    p.c = c; p.ldc_m = cm; p.ldc_n = cn;
    p.M = (int32_t)M; p.N = (int32_t)N; p.K = (int32_t)K;
    if (st == DN_OK && tf32_pair_eligible(M, N)) {  // CTA pairs, 256 x 256 tiles, three 64 KiB stages
        p.tiles_m = (int32_t)((M + 2 * kBM - 1) / (2 * kBM));
        p.tiles_n = (int32_t)((N + kBN2 - 1) / kBN2);
        st = make_map3(&ma, ahi, M, K, pa, kBM, false, 1, 0);
        if (st == DN_OK) st = make_map3(&mal, alo, M, K, pa, kBM, false, 1, 0);
        if (st == DN_OK) st = make_map3(&mb, bhi, N, K, pb, kBN2 / 2, false, 1, 0);
        if (st == DN_OK) st = make_map3(&mbl, blo, N, K, pb, kBN2 / 2, false, 1, 0);
        if (st == DN_OK) st = launch_tf32_2cta<false, false, true>(ma, mb, mal, mbl, p);
    } else {
    const int BN = N <= 32 ? 32 : 128;
    p.tiles_m = (int32_t)((M + kBM - 1) / kBM);
    p.tiles_n = (int32_t)((N + BN - 1) / BN);
    if (st == DN_OK) st = make_map(&ma, ahi, M, K, pa, kBM);
    if (st == DN_OK) st = make_map(&mal, alo, M, K, pa, kBM);
    if (st == DN_OK) st = make_map(&mb, bhi, N, K, pb, BN);
    if (st == DN_OK) st = make_map(&mbl, blo, N, K, pb, BN);
    if (st == DN_OK)
        st = BN == 32 ? launch_tf32<32, false, false, true>(ma, mb, mal, mbl, p)
                      : launch_tf32<128, false, false, true>(ma, mb, mal, mbl, p);
    }
    // an operand held an infinity or a NaN: the exact kernel recomputes the product (its launch returns at once otherwise)
    if (st == DN_OK)
        st = gemm_simt<float>(c, cm, cn, a, am, ak, b, bk, bn, M, N, K, 0, nullptr, nullptr, nullptr, nullptr, nonfinite);
    scratch_free(s0);
    scratch_free(s1);
    scratch_free(s2);
    scratch_free(s3);
    scratch_free(s_flag);
    return st;
}

// MatMatDot on float32 (dn_set_math_mode).
//   DN_MATH_FP32 (default): small problems (everything the reference's own tests multiply, e.g. "Single matrix dot",
//     Tensor.Test/CudaTests.fs:52-62, rel 1e-5 against the host) run on the exact fp32 SIMT kernel; large ones as 3xTF32
//     on the tensor cores: the INPUT error drops from tf32's 2^-11 to ~2^-21, what remains is the tensor core's
//     accumulator, which truncates (round toward zero) after every MMA — measured norm-wise 7e-6 at K = 1024
//     (a single tf32 pass: 5e-4; the SIMT kernel: 1e-7), growing linearly with K.
//   DN_MATH_FP32_STRICT: the SIMT kernel for every size (IEEE fp32 fused multiply-adds, round to nearest).
//   DN_MATH_TF32: one tf32 pass (rel 1e-2 per north_star), three times the throughput of the default.
dn_status gemm_f32(float *c, int64_t cm, int64_t cn, const float *a, int64_t am, int64_t ak, const float *b, int64_t bk,
                   int64_t bn, int64_t M, int64_t N, int64_t K) {
    const int mode = g_math_mode.load(std::memory_order_relaxed);
    if (mode == DN_MATH_TF32 || M == 0 || N == 0 || K == 0) return gemm_f32_tf32(c, cm, cn, a, am, ak, b, bk, bn, M, N, K);
    if (M >= (1ll << 31) || N >= (1ll << 31) || K >= (1ll << 31)) return set_error(DN_ERR_UNSUPPORTED, "MatMatDot: extent exceeds 2^31-1");
    if (mode == DN_MATH_FP32_STRICT || M * N * K < ((int64_t)1 << 27))
        return gemm_simt<float>(c, cm, cn, a, am, ak, b, bk, bn, M, N, K, 0, nullptr, nullptr, nullptr, nullptr);
    return gemm_f32_split(c, cm, cn, a, am, ak, b, bk, bn, M, N, K);
}

dn_status gemm_f64(double *c, int64_t cm, int64_t cn, const double *a, int64_t am, int64_t ak, const double *b, int64_t bk,
                   int64_t bn, int64_t M, int64_t N, int64_t K) {
    return gemm_simt<double>(c, cm, cn, a, am, ak, b, bk, bn, M, N, K, 0, nullptr, nullptr, nullptr, nullptr);
}

dn_status check_mm(const dn_tensor *t, const dn_tensor *a, const dn_tensor *b, int nd, const char *what) {
    if (!tensor_valid(t) || !tensor_valid(a) || !tensor_valid(b)) return set_error(DN_ERR_INVALID_ARG, "%s: bad argument", what);
    if (t->dtype != a->dtype || t->dtype != b->dtype) return set_error(DN_ERR_INVALID_ARG, "%s: operand types differ", what);
    if (t->dtype != DN_F32 && t->dtype != DN_F64)
        return set_error(DN_ERR_UNSUPPORTED, "%s is only supported for single and double (the reference falls back to "
                                             "broadcast-multiply + sumAxis for other types on the host only)", what);
    if (t->ndims != nd || a->ndims != nd || b->ndims != nd) return set_error(DN_ERR_SHAPE_MISMATCH, "%s: wrong rank", what);
    if (a->shape[nd - 1] != b->shape[nd - 2] || t->shape[nd - 2] != a->shape[nd - 2] || t->shape[nd - 1] != b->shape[nd - 1])
        return set_error(DN_ERR_SHAPE_MISMATCH, "%s: incompatible shapes", what);
    for (int d = 0; d < nd - 2; ++d)
        if (a->shape[d] != t->shape[d] || b->shape[d] != t->shape[d])
            return set_error(DN_ERR_SHAPE_MISMATCH, "%s: batch dimensions differ", what);
    return DN_OK;
}

dn_status mm_one(const dn_tensor *t, const dn_tensor *a, const dn_tensor *b, int64_t to, int64_t ao, int64_t bo) {
    const int nd = t->ndims;
    const int64_t M = a->shape[nd - 2], K = a->shape[nd - 1], N = b->shape[nd - 1];
    if (t->dtype == DN_F32)
        return gemm_f32(reinterpret_cast<float *>(data_ptr(t)) + to, t->stride[nd - 2], t->stride[nd - 1],
                        reinterpret_cast<const float *>(data_ptr(a)) + ao, a->stride[nd - 2], a->stride[nd - 1],
                        reinterpret_cast<const float *>(data_ptr(b)) + bo, b->stride[nd - 2], b->stride[nd - 1], M, N, K);
    return gemm_f64(reinterpret_cast<double *>(data_ptr(t)) + to, t->stride[nd - 2], t->stride[nd - 1],
                    reinterpret_cast<const double *>(data_ptr(a)) + ao, a->stride[nd - 2], a->stride[nd - 1],
                    reinterpret_cast<const double *>(data_ptr(b)) + bo, b->stride[nd - 2], b->stride[nd - 1], M, N, K);
}

// ---------------------------------------------------------------------------------------------------------------
// VecVecDot / MatVecDot: HBM-bound, one pass over the operands
// ---------------------------------------------------------------------------------------------------------------
// VecVecDot partial sums. VEC > 1: both vectors contiguous and 16-byte aligned -> 128-bit loads, 4 per operand in flight.
template <class T, int VEC>
__global__ void __launch_bounds__(256) dot_partial_kernel(const T *a, int64_t as, const T *b, int64_t bs, int64_t n, T *partials) {
    T acc = 0;
    if constexpr (VEC > 1) {
        const int64_t nv = n / VEC;
        const int64_t stride = (int64_t)gridDim.x * blockDim.x;
        int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        for (; i + 3 * stride < nv; i += 4 * stride) {
            Pack<T, VEC> x[4], y[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                x[u] = load_pack<Pack<T, VEC>>(reinterpret_cast<const char *>(a + (i + u * stride) * VEC));
                y[u] = load_pack<Pack<T, VEC>>(reinterpret_cast<const char *>(b + (i + u * stride) * VEC));
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int j = 0; j < VEC; ++j) acc += x[u].v[j] * y[u].v[j];
        }
        for (; i < nv; i += stride) {
            const Pack<T, VEC> x = load_pack<Pack<T, VEC>>(reinterpret_cast<const char *>(a + i * VEC));
            const Pack<T, VEC> y = load_pack<Pack<T, VEC>>(reinterpret_cast<const char *>(b + i * VEC));
#pragma unroll
            for (int j = 0; j < VEC; ++j) acc += x.v[j] * y.v[j];
        }
        if (blockIdx.x == 0 && threadIdx.x == 0)
            for (int64_t k = nv * VEC; k < n; ++k) acc += a[k] * b[k];
    } else {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
            acc += a[i * as] * b[i * bs];
    }
    __shared__ T sm[8];
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        T s = 0;
        for (int w = 0; w < 8; ++w) s += sm[w];
        partials[blockIdx.x] = s;
    }
}
template <class T>
__global__ void dot_final_kernel(const T *partials, int n, T *out) {
    T acc = 0;
    for (int i = threadIdx.x; i < n; i += 32) acc += partials[i];
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
    if (threadIdx.x == 0) *out = acc;
}

// y[m] = sum_k A[m,k] x[k].
// mode 2: k contiguous in A and x, 16-byte aligned rows -> a warp per row with 128-bit loads (x stays in L1/L2);
// mode 1: k is A's faster axis but not vectorisable -> warp per row, scalar.
// (m the faster axis, i.e. a transposed view: matvec_t_kernel below.)
template <class T>
__global__ void __launch_bounds__(256) matvec_kernel(T *y, int64_t ys, const T *a, int64_t am, int64_t ak, const T *x, int64_t xs,
                                                    int64_t M, int64_t K, int mode) {
    constexpr int VEC = 16 / (int)sizeof(T);
    {
        const int lane = threadIdx.x & 31;
        for (int64_t m = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); m < M; m += (int64_t)gridDim.x * 8) {
            T acc = 0;
            const T *row = a + m * am;
            if (mode == 2) {
                const int64_t kv = K / VEC;
                int64_t k = lane;
                for (; k + 96 < kv; k += 128) {
                    Pack<T, VEC> av[4], xv[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        av[u] = load_pack<Pack<T, VEC>>(reinterpret_cast<const char *>(row + (k + 32 * u) * VEC));
                        xv[u] = *reinterpret_cast<const Pack<T, VEC> *>(x + (k + 32 * u) * VEC);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int j = 0; j < VEC; ++j) acc += av[u].v[j] * xv[u].v[j];
                }
                for (; k < kv; k += 32) {
                    const Pack<T, VEC> av = load_pack<Pack<T, VEC>>(reinterpret_cast<const char *>(row + k * VEC));
                    const Pack<T, VEC> xv = *reinterpret_cast<const Pack<T, VEC> *>(x + k * VEC);
#pragma unroll
                    for (int j = 0; j < VEC; ++j) acc += av.v[j] * xv.v[j];
                }
                for (int64_t kk = kv * VEC + lane; kk < K; kk += 32) acc += row[kk] * x[kk];
            } else {
                for (int64_t k = lane; k < K; k += 32) acc += row[k * ak] * x[k * xs];
            }
#pragma unroll
            for (int s = 16; s >= 1; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
            if (lane == 0) y[m * ys] = acc;
        }
    }
}

// MatVecDot for a matrix whose faster axis is m (a transposed view): a thread owns VEC consecutive rows and one
// K-slice; slices are combined in slice order by matvec_t_final_kernel (deterministic, no atomics).
template <class T, int VEC>
__global__ void __launch_bounds__(256) matvec_t_kernel(T *out, int64_t out_ms, int64_t out_ss, const T *a, int64_t am, int64_t ak,
                                                      const T *x, int64_t xs, int64_t M, int64_t K, int64_t kchunk) {
    const int64_t mv = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (mv >= M) return;
    const int64_t k0 = (int64_t)blockIdx.y * kchunk;
    const int64_t k1 = k0 + kchunk < K ? k0 + kchunk : K;
    T acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = 0;
    const T *col = a + mv * am;
    int64_t k = k0;
    for (; k + 4 <= k1; k += 4) {
        Pack<T, VEC> av[4];
        T xv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if constexpr (VEC > 1) av[u] = load_pack<Pack<T, VEC>>(reinterpret_cast<const char *>(col + (k + u) * ak));
            else av[u].v[0] = col[(k + u) * ak];
            xv[u] = x[(k + u) * xs];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int j = 0; j < VEC; ++j) acc[j] += av[u].v[j] * xv[u];
    }
    for (; k < k1; ++k) {
        const T xk = x[k * xs];
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] += col[k * ak + j * am] * xk;
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) out[(mv + j) * out_ms + (int64_t)blockIdx.y * out_ss] = acc[j];
}
template <class T>
__global__ void matvec_t_final_kernel(T *y, int64_t ys, const T *partials, int64_t M, int slices) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    T acc = 0;
    for (int s = 0; s < slices; ++s) acc += partials[(int64_t)s * M + m];
    y[m * ys] = acc;
}

template <class T>
dn_status matvec_t_run(const dn_tensor *t, const dn_tensor *a, const dn_tensor *b) {
    constexpr int V = 16 / (int)sizeof(T);
    const int64_t M = a->shape[0], K = a->shape[1];
    const bool vec = a->stride[0] == 1 && M % V == 0 && (a->stride[1] * (int64_t)sizeof(T)) % 16 == 0 &&
                     (reinterpret_cast<uintptr_t>(data_ptr(a)) & 15) == 0;
    const int64_t threads = vec ? M / V : M;
    const int64_t gx = (threads + 255) / 256;
    // enough K-slices to fill the machine, each at least 64 deep
    int64_t slices = ((int64_t)sm_count() * 8 + gx - 1) / gx;
    if (slices > (K + 63) / 64) slices = (K + 63) / 64;
    if (slices < 1) slices = 1;
    if (slices > 65535) slices = 65535;
    const int64_t kchunk = (K + slices - 1) / slices;
    slices = K > 0 ? (K + kchunk - 1) / kchunk : 1;
    T *dst = (T *)data_ptr(t);
    int64_t out_ms = t->stride[0], out_ss = 0;
    void *scratch = nullptr;
    if (slices > 1) {
        dn_status st = scratch_alloc((size_t)slices * M * sizeof(T), &scratch);
        if (st != DN_OK) return st;
        dst = (T *)scratch;
        out_ms = 1;
        out_ss = M;
    }
    const dim3 grid((unsigned)gx, (unsigned)slices);
    if (vec)
        DN_LAUNCH((matvec_t_kernel<T, V>), grid, 256, 0, dst, out_ms, out_ss, (const T *)data_ptr(a), a->stride[0], a->stride[1],
                  (const T *)data_ptr(b), b->stride[0], M, K, kchunk);
    else
        DN_LAUNCH((matvec_t_kernel<T, 1>), grid, 256, 0, dst, out_ms, out_ss, (const T *)data_ptr(a), a->stride[0], a->stride[1],
                  (const T *)data_ptr(b), b->stride[0], M, K, kchunk);
    if (slices > 1) {
        DN_LAUNCH(matvec_t_final_kernel<T>, (unsigned)((M + 255) / 256), 256, 0, (T *)data_ptr(t), t->stride[0], (const T *)scratch, M,
                  (int)slices);
        scratch_free(scratch);
    }
    return launch_status("MatVecDot kernels");
}

}  // namespace

extern "C" {

dn_status dn_set_math_mode(int32_t mode) {
    if (mode != DN_MATH_FP32 && mode != DN_MATH_TF32 && mode != DN_MATH_FP32_STRICT) return set_error(DN_ERR_INVALID_ARG, "dn_set_math_mode: bad mode %d", mode);
    g_math_mode.store(mode, std::memory_order_relaxed);
    return DN_OK;
}

dn_status dn_get_math_mode(int32_t *mode) {
    if (!mode) return set_error(DN_ERR_INVALID_ARG, "dn_get_math_mode: null argument");
    *mode = g_math_mode.load(std::memory_order_relaxed);
    return DN_OK;
}

dn_status dn_mat_mat_dot(const dn_tensor *t, const dn_tensor *a, const dn_tensor *b) {
    dn_status st = check_mm(t, a, b, 2, "MatMatDot");
    if (st != DN_OK) return st;
    return mm_one(t, a, b, 0, 0, 0);
}

dn_status dn_batched_mat_mat_dot(const dn_tensor *t, const dn_tensor *a, const dn_tensor *b) {
    if (!tensor_valid(t) || t->ndims < 2) return set_error(DN_ERR_INVALID_ARG, "BatchedMatMatDot: bad argument");
    const int nd = t->ndims;
    dn_status st = check_mm(t, a, b, nd, "BatchedMatMatDot");
    if (st != DN_OK) return st;
    int64_t nbatch = 1;
    for (int d = 0; d < nd - 2; ++d) nbatch *= t->shape[d];
    const int64_t M = a->shape[nd - 2], K = a->shape[nd - 1], N = b->shape[nd - 1];
    // float64, and float32 batches of small matrices (below 2^27 multiply-adds each — measured: 64 x 256^3 runs at
    // 3 TFLOP/s as per-element tcgen05 launches, launch- and tensor-map-bound): ONE launch of the SIMT kernel
    if (nbatch > 1 && (t->dtype == DN_F64 || M * N * K < ((int64_t)1 << 27))) {
        int64_t bshape[DN_MAX_DIMS], bt[DN_MAX_DIMS], ba[DN_MAX_DIMS], bb[DN_MAX_DIMS];
        int nbd = 0;
        for (int d = nd - 3; d >= 0; --d) {  // innermost batch dim first; broadcast batch dims have stride 0
            bshape[nbd] = t->shape[d];
            bt[nbd] = t->stride[d];
            ba[nbd] = a->stride[d];
            bb[nbd] = b->stride[d];
            ++nbd;
        }
        if (t->dtype == DN_F32)
            return gemm_simt<float>(reinterpret_cast<float *>(data_ptr(t)), t->stride[nd - 2], t->stride[nd - 1],
                                    reinterpret_cast<const float *>(data_ptr(a)), a->stride[nd - 2], a->stride[nd - 1],
                                    reinterpret_cast<const float *>(data_ptr(b)), b->stride[nd - 2], b->stride[nd - 1], M, N, K,
                                    nbd, bshape, bt, ba, bb);
        return gemm_simt<double>(reinterpret_cast<double *>(data_ptr(t)), t->stride[nd - 2], t->stride[nd - 1],
                                 reinterpret_cast<const double *>(data_ptr(a)), a->stride[nd - 2], a->stride[nd - 1],
                                 reinterpret_cast<const double *>(data_ptr(b)), b->stride[nd - 2], b->stride[nd - 1], M, N, K,
                                 nbd, bshape, bt, ba, bb);
    }
    // Large float32 batch elements in tf32 mode: ONE persistent launch of the CTA-pair kernel over all (batch, tile)
    // pairs through 3-D tensor maps (the reference: one cublasSgemmBatched call, CudaBackend.fs:426-449), when the
    // batch dims of every operand collapse to a single stride (or are all broadcast) and the matrices are TMA-ready.
    if (nbatch > 1 && t->dtype == DN_F32 && g_math_mode.load(std::memory_order_relaxed) == DN_MATH_TF32 &&
        tf32_pair_eligible(M, N) && nbatch < (1ll << 20)) {
        auto collapse = [&](const dn_tensor *x, int64_t *bs) {  // single batch stride, 0 = broadcast; false = neither
            bool all_zero = true;
            for (int d = 0; d < nd - 2; ++d)
                if (x->shape[d] > 1 && x->stride[d] != 0) all_zero = false;
            if (all_zero) { *bs = 0; return true; }
            int64_t expect = -1;
            for (int d = nd - 3; d >= 0; --d) {
                if (x->shape[d] == 1) continue;
                if (expect < 0) *bs = x->stride[d];
                else if (x->stride[d] != expect) return false;
                expect = x->stride[d] * x->shape[d];
            }
            return *bs > 0;
        };
        int64_t cbs = 0, abs_ = 0, bbs = 0;
        const float *ap = reinterpret_cast<const float *>(data_ptr(a)), *bp = reinterpret_cast<const float *>(data_ptr(b));
        Operand2D A{ap, M, K, a->stride[nd - 2], a->stride[nd - 1]}, B{bp, N, K, b->stride[nd - 1], b->stride[nd - 2]};
        const bool a_ok = tma_ready(A) || tma_ready_mn(A), b_ok = tma_ready(B) || tma_ready_mn(B);
        if (a_ok && b_ok && collapse(t, &cbs) && cbs > 0 && collapse(a, &abs_) && collapse(b, &bbs) && abs_ % 4 == 0 &&
            bbs % 4 == 0)
            return gemm_f32_tf32(reinterpret_cast<float *>(data_ptr(t)), t->stride[nd - 2], t->stride[nd - 1], ap,
                                 a->stride[nd - 2], a->stride[nd - 1], bp, b->stride[nd - 2], b->stride[nd - 1], M, N, K, nbatch,
                                 cbs, abs_, bbs);
    }
    for (int64_t bi = 0; bi < nbatch; ++bi) {
        int64_t rem = bi, to = 0, ao = 0, bo = 0;
        for (int d = nd - 3; d >= 0; --d) {
            const int64_t x = rem % t->shape[d];
            rem /= t->shape[d];
            to += x * t->stride[d];
            ao += x * a->stride[d];  // broadcast batch dims have stride 0
            bo += x * b->stride[d];
        }
        st = mm_one(t, a, b, to, ao, bo);
        if (st != DN_OK) return st;
    }
    return DN_OK;
}

dn_status dn_vec_vec_dot(const dn_tensor *t, const dn_tensor *a, const dn_tensor *b) {
    if (!tensor_valid(t) || !tensor_valid(a) || !tensor_valid(b)) return set_error(DN_ERR_INVALID_ARG, "VecVecDot: bad argument");
    if (t->ndims != 0 || a->ndims != 1 || b->ndims != 1 || a->shape[0] != b->shape[0])
        return set_error(DN_ERR_SHAPE_MISMATCH, "VecVecDot: incompatible shapes");
    if (t->dtype != a->dtype || t->dtype != b->dtype || (t->dtype != DN_F32 && t->dtype != DN_F64))
        return set_error(DN_ERR_UNSUPPORTED, "VecVecDot is only supported for single and double");
    const int64_t n = a->shape[0];
    int grid = (int)((n + 8191) / 8192);
    const int cap = sm_count() * 8;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    void *scratch = nullptr;
    dn_status st = scratch_alloc((size_t)grid * 8, &scratch);
    if (st != DN_OK) return st;
    const int esz = dtype_size(t->dtype);
    const bool vec = a->stride[0] == 1 && b->stride[0] == 1 && (reinterpret_cast<uintptr_t>(data_ptr(a)) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(data_ptr(b)) & 15) == 0 && n >= 16 / esz;
    if (t->dtype == DN_F32) {
        if (vec)
            DN_LAUNCH((dot_partial_kernel<float, 4>), grid, 256, 0, (const float *)data_ptr(a), a->stride[0],
                      (const float *)data_ptr(b), b->stride[0], n, (float *)scratch);
        else
            DN_LAUNCH((dot_partial_kernel<float, 1>), grid, 256, 0, (const float *)data_ptr(a), a->stride[0],
                      (const float *)data_ptr(b), b->stride[0], n, (float *)scratch);
        DN_LAUNCH(dot_final_kernel<float>, 1, 32, 0, (const float *)scratch, grid, (float *)data_ptr(t));
    } else {
        if (vec)
            DN_LAUNCH((dot_partial_kernel<double, 2>), grid, 256, 0, (const double *)data_ptr(a), a->stride[0],
                      (const double *)data_ptr(b), b->stride[0], n, (double *)scratch);
        else
            DN_LAUNCH((dot_partial_kernel<double, 1>), grid, 256, 0, (const double *)data_ptr(a), a->stride[0],
                      (const double *)data_ptr(b), b->stride[0], n, (double *)scratch);
        DN_LAUNCH(dot_final_kernel<double>, 1, 32, 0, (const double *)scratch, grid, (double *)data_ptr(t));
    }
    scratch_free(scratch);
    return launch_status("VecVecDot kernels");
}

dn_status dn_mat_vec_dot(const dn_tensor *t, const dn_tensor *a, const dn_tensor *b) {
    if (!tensor_valid(t) || !tensor_valid(a) || !tensor_valid(b)) return set_error(DN_ERR_INVALID_ARG, "MatVecDot: bad argument");
    if (t->ndims != 1 || a->ndims != 2 || b->ndims != 1 || a->shape[1] != b->shape[0] || t->shape[0] != a->shape[0])
        return set_error(DN_ERR_SHAPE_MISMATCH, "MatVecDot: incompatible shapes");
    if (t->dtype != a->dtype || t->dtype != b->dtype || (t->dtype != DN_F32 && t->dtype != DN_F64))
        return set_error(DN_ERR_UNSUPPORTED, "MatVecDot is only supported for single and double");
    const int64_t M = a->shape[0], K = a->shape[1];
    if (M == 0) return DN_OK;
    const int64_t abs_k = a->stride[1] < 0 ? -a->stride[1] : a->stride[1];
    const int64_t abs_m = a->stride[0] < 0 ? -a->stride[0] : a->stride[0];
    const int esz = dtype_size(t->dtype);
    int mode = abs_k <= abs_m ? 1 : 0;
    if (mode == 0) return t->dtype == DN_F32 ? matvec_t_run<float>(t, a, b) : matvec_t_run<double>(t, a, b);
    if (mode == 1 && a->stride[1] == 1 && b->stride[0] == 1 && (a->stride[0] * esz) % 16 == 0 &&
        (reinterpret_cast<uintptr_t>(data_ptr(a)) & 15) == 0 && (reinterpret_cast<uintptr_t>(data_ptr(b)) & 15) == 0)
        mode = 2;
    int64_t ctas = (M + 7) / 8;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (ctas > cap) ctas = cap;
    if (t->dtype == DN_F32)
        DN_LAUNCH(matvec_kernel<float>, (unsigned)ctas, 256, 0, (float *)data_ptr(t), t->stride[0], (const float *)data_ptr(a),
                  a->stride[0], a->stride[1], (const float *)data_ptr(b), b->stride[0], M, K, mode);
    else
        DN_LAUNCH(matvec_kernel<double>, (unsigned)ctas, 256, 0, (double *)data_ptr(t), t->stride[0], (const double *)data_ptr(a),
                  a->stride[0], a->stride[1], (const double *)data_ptr(b), b->stride[0], M, K, mode);
    return launch_status("MatVecDot kernel");
}

}  // extern "C"
