// Blackwell-native GEMM for the BART decoder/encoder linears and the lm_head:
//   C[M,N] = A[M,K] * W[N,K]^T + bias[N]   (optional exact GELU), fp32 in / fp32 out,
// computed on the 5th-gen tensor cores as an error-compensated 3xTF32 product
//   A*W ~= A_lo*W_hi + A_hi*W_lo + A_hi*W_hi,   x_hi = x with the 13 low mantissa bits cleared,
//                                                x_lo = x - x_hi (exact),
// which keeps the fp32-level accuracy the 1e-4 beam-score parity needs (a single TF32/BF16 pass
// does not).  tcgen05.mma (kind::tf32, M=128, N=BN, K=8) issued by one thread, operands staged by
// TMA into 128B-swizzled K-major shared-memory tiles, fp32 accumulators in TMEM, epilogue
// tcgen05.ld -> registers -> global.  Warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM
// allocator, 4..7 = epilogue.  One CTA per 128 x BN output tile.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "launch.cuh"

namespace sealb200 {

constexpr int UM = 128;            // tile rows  (UMMA M)
constexpr int UK = 32;             // k-block: 32 fp32 = 128 B = one swizzle row
constexpr int USTAGES = 2;
constexpr int UTHREADS = 256;

template <int BN>
struct UmmaSmem {
    static constexpr int kABytes = UM * 128;        // one A tile (hi or lo)
    static constexpr int kWBytes = BN * 128;
    static constexpr int kStageBytes = 2 * kABytes + 2 * kWBytes;
    static constexpr int kTotal = USTAGES * kStageBytes + 1024 /*alignment slack*/ + 256 /*barriers*/;
    static constexpr int kTotalStaged = kTotal + 16 * 2048;     // + per-warp epilogue transpose buffers
};

// ---- PTX wrappers --------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded spin: a protocol bug must trap (error code back to the host), never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done && spin > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp:SmemDescriptor):
// start>>4 | LBO(1)<<16 | SBO(1024 B >> 4)<<32 | version(1)<<46 | layout SWIZZLE_128B(2)<<61
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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

__device__ __forceinline__ float gelu_erf_u(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// x -> (hi, lo): hi keeps the TF32 bits (sign, exponent, 10 mantissa bits), lo = x - hi exactly.
__global__ void __launch_bounds__(256) split_tf32_kernel(int64_t n4, const float4* __restrict__ x, float4* __restrict__ hi,
                                                         float4* __restrict__ lo) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = x[i];
        float4 h, l;
        h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
        h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
        h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
        h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
        hi[i] = h; lo[i] = l;
    }
}

// Accumulation note (measured on B200, profiles/r01_decode_mode1_v1_fail.log): the tensor core adds
// into its fp32 TMEM accumulator with truncation, so a long K loop (K/8 x 3 sequential adds) drifts
// by ~N_acc * 2^-24 relative, in one direction — 7e-6 at K = 1024, enough to push summed beam
// scores past the 1e-4 parity bound.  Therefore the K loop is cut into chunks of UKC k-blocks
// (K_c = 64): each chunk accumulates in one of two TMEM buffers (24 adds), and the epilogue warps
// drain finished chunks into fp32 REGISTER accumulators with round-to-nearest adds while the next
// chunk's MMAs run into the other buffer.
constexpr int UKC = 2;                                       // k-blocks per TMEM chunk (K_c = 64: 24 adds per chunk)
constexpr int UEPI_WARPS = 16;                               // 4 lane quarters x 4 column groups of 64
constexpr int UTHREADS2 = (4 + UEPI_WARPS) * 32;             // 640

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

template <int BN, bool GELU>
__global__ void __launch_bounds__(UTHREADS2, 1)
umma_gemm_tf32x3_persistent_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                        const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                        int M, int N, int K, const float* __restrict__ bias, float* __restrict__ C,
                        float* __restrict__ C_hi, float* __restrict__ C_lo, int ldc, int n_fastest) {
    static_assert(BN == 256, "epilogue mapping assumes a 256-column tile (2 TMEM buffers = 512 columns)");
    using SM = UmmaSmem<BN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;          // SWIZZLE_128B wants 1024 B alignment
    const uint32_t bars = base + USTAGES * SM::kStageBytes;
    const uint32_t full0 = bars, empty0 = bars + 8 * USTAGES;             // smem stage barriers
    const uint32_t tfull0 = bars + 16 * USTAGES, tempty0 = tfull0 + 16;   // 2 TMEM buffers
    const uint32_t slot = tempty0 + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // persistent: CTA b walks tiles b, b + gridDim.x, ...  Tile order is chosen by the host so that
    // the LARGER operand is streamed from HBM once: n fastest when the activations dominate (the
    // concurrently running CTAs then share A tiles and all of W stays in L2), m fastest when the
    // weights dominate (lm_head).
    const int m_tiles = (M + UM - 1) / UM, n_tiles = (N + BN - 1) / BN;
    const int total_tiles = m_tiles * n_tiles;
    const int num_k = K / UK;
    const int num_chunks = (num_k + UKC - 1) / UKC;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < USTAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull0 + 8 * b, 1); mbar_init(tempty0 + 8 * b, UEPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    } else if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(slot));

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;                                   // k-blocks issued so far (all tiles)
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int m_tile = n_fastest ? tile / n_tiles : tile % m_tiles, n_tile = n_fastest ? tile % n_tiles : tile / m_tiles;
            for (int kb = 0; kb < num_k; ++kb, ++it) {
                const int s = it % USTAGES;
                const uint32_t ph = (it / USTAGES) & 1;
                mbar_wait(empty0 + 8 * s, ph ^ 1);
                const uint32_t st = base + s * SM::kStageBytes;
                mbar_expect_tx(full0 + 8 * s, SM::kStageBytes);
                tma_load_2d(st, &tmA_hi, full0 + 8 * s, kb * UK, m_tile * UM);
                tma_load_2d(st + SM::kABytes, &tmA_lo, full0 + 8 * s, kb * UK, m_tile * UM);
                tma_load_2d(st + 2 * SM::kABytes, &tmW_hi, full0 + 8 * s, kb * UK, n_tile * BN);
                tma_load_2d(st + 2 * SM::kABytes + SM::kWBytes, &tmW_lo, full0 + 8 * s, kb * UK, n_tile * BN);
            }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor (cute/arch/mma_sm100_desc.hpp:InstrDescriptor): D=F32, A=B=TF32, K-major both,
            // N>>3 at bit 17, M>>4 at bit 24
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(UM >> 4) << 24);
            uint32_t it = 0, ch = 0;                           // k-blocks / chunks consumed so far (all tiles)
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            int kb = 0;
            for (int c = 0; c < num_chunks; ++c, ++ch) {
                const int buf = ch & 1;
                mbar_wait(tempty0 + 8 * buf, ((ch >> 1) & 1) ^ 1);    // epilogue has drained this buffer's previous use
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(buf * BN);
                const int kend = (kb + UKC < num_k) ? kb + UKC : num_k;
                for (int k0 = kb; kb < kend; ++kb, ++it) {
                    const int s = it % USTAGES;
                    const uint32_t ph = (it / USTAGES) & 1;
                    mbar_wait(full0 + 8 * s, ph);
                    tc_fence_after();
                    const uint32_t st = base + s * SM::kStageBytes;
                    const uint64_t a_hi = umma_desc_sw128(st), a_lo = umma_desc_sw128(st + SM::kABytes);
                    const uint64_t w_hi = umma_desc_sw128(st + 2 * SM::kABytes), w_lo = umma_desc_sw128(st + 2 * SM::kABytes + SM::kWBytes);
#pragma unroll
                    for (int k = 0; k < UK / 8; ++k) {         // UMMA_K = 8 tf32 = 32 B -> +2 in the >>4 address field
                        umma_tf32(tacc, a_lo + 2 * k, w_hi + 2 * k, idesc, (kb != k0) || (k != 0));
                        umma_tf32(tacc, a_hi + 2 * k, w_lo + 2 * k, idesc, 1);
                        umma_tf32(tacc, a_hi + 2 * k, w_hi + 2 * k, idesc, 1);
                    }
                    umma_commit(empty0 + 8 * s);               // frees the smem stage when these MMAs retire
                }
                umma_commit(tfull0 + 8 * buf);                 // this chunk's partial sums are complete
            }
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;                                // TMEM lane quarter this warp may touch (warp % 4)
        const int cg = (warp - 4) >> 2;                        // column group: 64 columns
        uint32_t ch = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_tile = n_fastest ? tile / n_tiles : tile % m_tiles, n_tile = n_fastest ? tile % n_tiles : tile / m_tiles;
        float acc[64];
#pragma unroll
        for (int j = 0; j < 64; ++j) acc[j] = 0.f;
        for (int c = 0; c < num_chunks; ++c, ++ch) {
            const int buf = ch & 1;
            mbar_wait(tfull0 + 8 * buf, (ch >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + cg * 64 + h * 32), r);
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[h * 32 + j] += __uint_as_float(r[j]);      // round-to-nearest promotion
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
        }
        const int row = m_tile * UM + q * 32 + lane;
        const int nb = n_tile * BN + cg * 64;
        if (row < M && nb < N) {
            const int64_t roff = (int64_t)row * ldc;
#pragma unroll
            for (int j = 0; j < 64; j += 4) {
                const int n = nb + j;
                float v[4], vh[4], vl[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float x = acc[j + u] + ((bias && n + u < N) ? bias[n + u] : 0.f);
                    v[u] = GELU ? gelu_erf_u(x) : x;
                    vh[u] = __uint_as_float(__float_as_uint(v[u]) & 0xFFFFE000u);    // TF32 split for the next GEMM
                    vl[u] = v[u] - vh[u];
                }
                if (n + 3 < N) {
                    if (C) *reinterpret_cast<float4*>(C + roff + n) = make_float4(v[0], v[1], v[2], v[3]);
                    if (C_hi) {
                        *reinterpret_cast<float4*>(C_hi + roff + n) = make_float4(vh[0], vh[1], vh[2], vh[3]);
                        *reinterpret_cast<float4*>(C_lo + roff + n) = make_float4(vl[0], vl[1], vl[2], vl[3]);
                    }
                } else {
                    for (int u = 0; u < 4; ++u) if (n + u < N) {
                        if (C) C[roff + n + u] = v[u];
                        if (C_hi) { C_hi[roff + n + u] = vh[u]; C_lo[roff + n + u] = vl[u]; }
                    }
                }
            }
        }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- 3xFP16 variant (gemm_mode 3) -------------------------------------------------------------------
// Same persistent skeleton, operands split into two IEEE half values instead of two TF32 values:
//   x = h1 + h2 (+ <= 2^-22 |x|),  h1 = rn_half(x), h2 = rn_half(x - h1);   A*W ~= A2*W1 + A1*W2 + A1*W1.
// fp16 products are exact in the fp32 accumulator, so the accuracy class is the same as 3xTF32, but an
// element costs 4 bytes instead of 8 (the TF32 kernel is bound by L2->shared-memory operand traffic,
// profiles/r01_ncu_umma_fc1_raw.csv: tensor pipe 65 % active) and kind::f16 runs at twice the TF32
// rate.  Range: activations are used unscaled (|x| must stay below 65504; the producers saturate and
// raise a sticky flag otherwise, and the absolute floor of a subnormal h2, 3e-8, is far below the
// 1e-5 logit budget); each weight matrix is pre-multiplied by a power of two 2^s so that
// max|W| ~ 2^14 and the epilogue multiplies the accumulator by 2^-s (exact).
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

constexpr int UK16 = 64;            // k-block: 64 halves = 128 B = one swizzle row
constexpr int UKC16 = 4;            // k-blocks (of 64 halves) per TMEM chunk
constexpr float kHalfMax = 65504.f;

__device__ __forceinline__ void split_half(float x, __half& h1, __half& h2, int* overflow) {
    if (fabsf(x) > kHalfMax) { if (overflow) *overflow = 1; x = copysignf(kHalfMax, x); }
    h1 = __float2half_rn(x);
    h2 = __float2half_rn(x - __half2float(h1));
}

// x * scale -> (h1, h2) halves; used for weights (once) and for activations whose producer did not split
__global__ void __launch_bounds__(256) split_half_kernel(int64_t n, const float* __restrict__ x, float scale,
                                                         __half* __restrict__ h1, __half* __restrict__ h2,
                                                         int* __restrict__ overflow) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        __half a, b;
        int ov = 0;
        split_half(x[i] * scale, a, b, &ov);
        if (ov) atomicExch(overflow, 1);
        h1[i] = a; h2[i] = b;
    }
}

__global__ void __launch_bounds__(256) absmax_kernel(int64_t n, const float* __restrict__ x, unsigned int* __restrict__ out) {
    float m = 0.f;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = fabsf(x[i]);
        if (v == v && v != INFINITY) m = fmaxf(m, v);
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));       // non-negative floats order like uints
}

// K-major shared-memory matrix descriptor for 128-byte (SWIZZLE_128B) or 64-byte (SWIZZLE_64B) rows:
// 8-row groups are ROWB*8 bytes apart (SBO); layout code 2 / 4 (cute/arch/mma_sm100_desc.hpp).
template <int ROWB>
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    static_assert(ROWB == 128 || ROWB == 64, "row bytes");
    constexpr uint64_t sbo = (ROWB * 8) >> 4;
    constexpr uint64_t layout = ROWB == 128 ? 2 : 4;
    return static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}

// Optional in-kernel timeline of CTA 0 (debug ABI sealdec_debug_gemm_trace): SM cycle counter at
// 0 entry, 1 prologue done, 2 first operands landed, 3 last MMA issued, 4 last chunk complete,
// 5 tile stored, 6 exit; 7/8 = %globaltimer (ns) at entry / exit; 9.. = epilogue sub-steps of warp 4
// (after staging and after the stores of each of the 4 passes).
__device__ long long g_gemm_trace[20];
__device__ int g_gemm_trace_on;
__device__ __forceinline__ void gemm_trace(int i) {
    if (blockIdx.x == 0 && g_gemm_trace_on) {
        g_gemm_trace[i] = clock64();
        if (i == 0 || i == 6) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); g_gemm_trace[i == 0 ? 7 : 8] = (long long)t; }
    }
}

template <int BN, bool GELU, int ROWB>
__global__ void __launch_bounds__(UTHREADS2, 1)
umma_gemm_f16x3_persistent_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                        const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                        int M, int N, int K, const float* __restrict__ bias, float w_unscale, float* __restrict__ C,
                        __half* __restrict__ C_h1, __half* __restrict__ C_h2, int ldc, int n_fastest,
                        int* __restrict__ overflow, int k_slices, int64_t slice_stride) {
    static_assert(BN == 256, "epilogue mapping assumes a 256-column tile (2 TMEM buffers = 512 columns)");
    // ROWB = bytes of K per shared-memory row: 128 -> SWIZZLE_128B, 64 K-halves per k-block, 2 stages of
    // 96 KB; 64 -> SWIZZLE_64B, 32 K-halves per k-block, 4 stages of 48 KB (three loads in flight
    // instead of one: the 2-stage kernel leaves the tensor pipe idle ~35 % of the time waiting for TMA,
    // profiles/r01_ncu_umma_fc1_raw.csv).
    constexpr int NST = (ROWB == 128) ? 2 : 4;
    constexpr int KE = ROWB / 2;                                           // K halves per k-block
    constexpr int kAB = UM * ROWB, kWB = BN * ROWB, kStage = 2 * kAB + 2 * kWB;
    // K per TMEM chunk: 256 (48 truncating tensor-core adds per chunk; measured error still below the fp32
    // SIMT kernel's) so that the MMA warp can run two chunks = half a K=1024 tile ahead while the
    // epilogue warps are busy with the previous tile's bias/GELU/split/stores
    constexpr int kChunkBlocks = (ROWB == 128) ? UKC16 : 2 * UKC16;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;          // swizzle atoms want 1024 B alignment
    const uint32_t bars = base + NST * kStage;
    const uint32_t full0 = bars, empty0 = bars + 8 * NST;                 // smem stage barriers
    const uint32_t tfull0 = bars + 16 * NST, tempty0 = tfull0 + 16;       // 2 TMEM buffers
    const uint32_t slot = tempty0 + 16;
    // per-epilogue-warp 2 KB transpose buffers (32 rows x 16 floats, XOR-swizzled 16-byte chunks): the
    // TMEM load gives lane = row, but rows are ldc floats apart in memory, so storing straight from the
    // registers touches 32 cache lines per instruction (measured: the store phase of a tile stalled the
    // MMAs of the next tile for ~15 us, profiles/r01_ncu_f16_fc1_raw.csv: tensor pipe 49 % active);
    // after the transpose 4 lanes cover 64 contiguous bytes of a row, 8 rows per instruction.
    float* stage_base = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)) + NST * kStage + 256);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) gemm_trace(0);
    // persistent: CTA b walks tiles b, b + gridDim.x, ...  Tile order is chosen by the host so that
    // the LARGER operand is streamed from HBM once: n fastest when the activations dominate (the
    // concurrently running CTAs then share A tiles and all of W stays in L2), m fastest when the
    // weights dominate (lm_head).
    const int m_tiles = (M + UM - 1) / UM, n_tiles = (N + BN - 1) / BN;
    // split-K (skinny M: a handful of tiles would leave most SMs idle and each tile's K loop is a serial
    // chain of TMA round trips): work item = (tile, K slice); slice s accumulates k-blocks
    // [s*num_k, (s+1)*num_k) and stores its raw fp32 partial tile at C + s*slice_stride (the caller
    // passes bias = nullptr, w_unscale = 1, no half outputs; umma_splitk_finish_kernel sums the slices
    // in a fixed order and applies scale / bias / GELU / split).
    const int total_tiles = m_tiles * n_tiles * k_slices;
    const int num_k = (K / KE) / k_slices;
    const int num_chunks = (num_k + kChunkBlocks - 1) / kChunkBlocks;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull0 + 8 * b, 1); mbar_init(tempty0 + 8 * b, UEPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    } else if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(slot));
    if (threadIdx.x == 0) gemm_trace(1);

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;                                   // k-blocks issued so far (all tiles)
            for (int item = blockIdx.x; item < total_tiles; item += gridDim.x) {
            const int tile = item / k_slices, kb0 = (item % k_slices) * num_k;
            const int m_tile = n_fastest ? tile / n_tiles : tile % m_tiles, n_tile = n_fastest ? tile % n_tiles : tile / m_tiles;
            for (int kb = 0; kb < num_k; ++kb, ++it) {
                const int s = it % NST;
                const uint32_t ph = (it / NST) & 1;
                mbar_wait(empty0 + 8 * s, ph ^ 1);
                const uint32_t st = base + s * kStage;
                mbar_expect_tx(full0 + 8 * s, kStage);
                tma_load_2d(st, &tmA_hi, full0 + 8 * s, (kb0 + kb) * KE, m_tile * UM);
                tma_load_2d(st + kAB, &tmA_lo, full0 + 8 * s, (kb0 + kb) * KE, m_tile * UM);
                tma_load_2d(st + 2 * kAB, &tmW_hi, full0 + 8 * s, (kb0 + kb) * KE, n_tile * BN);
                tma_load_2d(st + 2 * kAB + kWB, &tmW_lo, full0 + 8 * s, (kb0 + kb) * KE, n_tile * BN);
            }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor: D=F32 (1<<4), A=B=F16 (format 0), K-major both, N>>3 at bit 17, M>>4 at bit 24
            const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(UM >> 4) << 24);
            uint32_t it = 0, ch = 0;                           // k-blocks / chunks consumed so far (all tiles)
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            int kb = 0;
            for (int c = 0; c < num_chunks; ++c, ++ch) {
                const int buf = ch & 1;
                mbar_wait(tempty0 + 8 * buf, ((ch >> 1) & 1) ^ 1);    // epilogue has drained this buffer's previous use
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(buf * BN);
                const int kend = (kb + kChunkBlocks < num_k) ? kb + kChunkBlocks : num_k;
                for (int k0 = kb; kb < kend; ++kb, ++it) {
                    const int s = it % NST;
                    const uint32_t ph = (it / NST) & 1;
                    mbar_wait(full0 + 8 * s, ph);
                    if (it == 0) gemm_trace(2);
                    tc_fence_after();
                    const uint32_t st = base + s * kStage;
                    const uint64_t a_hi = umma_desc<ROWB>(st), a_lo = umma_desc<ROWB>(st + kAB);
                    const uint64_t w_hi = umma_desc<ROWB>(st + 2 * kAB), w_lo = umma_desc<ROWB>(st + 2 * kAB + kWB);
#pragma unroll
                    for (int k = 0; k < KE / 16; ++k) {        // UMMA_K = 16 halves = 32 B -> +2 in the >>4 address field
                        umma_f16(tacc, a_lo + 2 * k, w_hi + 2 * k, idesc, (kb != k0) || (k != 0));
                        umma_f16(tacc, a_hi + 2 * k, w_lo + 2 * k, idesc, 1);
                        umma_f16(tacc, a_hi + 2 * k, w_hi + 2 * k, idesc, 1);
                    }
                    umma_commit(empty0 + 8 * s);               // frees the smem stage when these MMAs retire
                }
                umma_commit(tfull0 + 8 * buf);                 // this chunk's partial sums are complete
            }
            if (tile == 0) gemm_trace(3);
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;                                // TMEM lane quarter this warp may touch (warp % 4)
        const int cg = (warp - 4) >> 2;                        // column group: 64 columns
        uint32_t ch = 0;
        for (int item = blockIdx.x; item < total_tiles; item += gridDim.x) {
        const int tile = item / k_slices;
        float* Cs = C ? C + (int64_t)(item % k_slices) * slice_stride : nullptr;
        const int m_tile = n_fastest ? tile / n_tiles : tile % m_tiles, n_tile = n_fastest ? tile % n_tiles : tile / m_tiles;
        float acc[64];
#pragma unroll
        for (int j = 0; j < 64; ++j) acc[j] = 0.f;
        for (int c = 0; c < num_chunks; ++c, ++ch) {
            const int buf = ch & 1;
            mbar_wait(tfull0 + 8 * buf, (ch >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + cg * 64 + h * 32), r);
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[h * 32 + j] += __uint_as_float(r[j]);      // round-to-nearest promotion
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
        }
        const bool tr = item == 0 && warp == 4 && lane == 0;
        if (tr) gemm_trace(4);
        const int row0 = m_tile * UM + q * 32;                 // this warp's 32 rows
        const int nb = n_tile * BN + cg * 64;                  // ... and 64 columns
        float* stg = stage_base + (warp - 4) * 512;            // 32 x 16 floats
        if (row0 < M && nb < N) {
#pragma unroll
            for (int pass = 0; pass < 4; ++pass) {             // 16 columns per pass
                // finish the values in registers (lane = row), park them transposed-friendly in smem
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    float v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int n = nb + pass * 16 + j4 * 4 + u;
                        const float x = acc[pass * 16 + j4 * 4 + u] * w_unscale + ((bias && n < N) ? bias[n] : 0.f);
                        v[u] = GELU ? gelu_erf_u(x) : x;
                    }
                    const int phys = j4 ^ ((lane >> 1) & 3);
                    *reinterpret_cast<float4*>(stg + lane * 16 + phys * 4) = make_float4(v[0], v[1], v[2], v[3]);
                }
                __syncwarp();
                if (tr) gemm_trace(9 + 2 * pass);
#pragma unroll
                for (int i = 0; i < 4; ++i) {                  // 8 rows x 64 B per instruction
                    const int rr = i * 8 + (lane >> 2), ch = lane & 3;
                    const float4 o = *reinterpret_cast<const float4*>(stg + rr * 16 + (ch ^ ((rr >> 1) & 3)) * 4);
                    const int row = row0 + rr;
                    const int n = nb + pass * 16 + ch * 4;
                    if (row < M && n < N) {
                        const int64_t off = (int64_t)row * ldc + n;
                        if (n + 3 < N) {
                            if (Cs) *reinterpret_cast<float4*>(Cs + off) = o;
                            if (C_h1) {
                                __half h1[4], h2[4];
                                int ov = 0;
                                split_half(o.x, h1[0], h2[0], &ov); split_half(o.y, h1[1], h2[1], &ov);
                                split_half(o.z, h1[2], h2[2], &ov); split_half(o.w, h1[3], h2[3], &ov);
                                if (ov) atomicExch(overflow, 1);
                                *reinterpret_cast<uint2*>(C_h1 + off) = make_uint2(
                                    (uint32_t)__half_as_ushort(h1[0]) | ((uint32_t)__half_as_ushort(h1[1]) << 16),
                                    (uint32_t)__half_as_ushort(h1[2]) | ((uint32_t)__half_as_ushort(h1[3]) << 16));
                                *reinterpret_cast<uint2*>(C_h2 + off) = make_uint2(
                                    (uint32_t)__half_as_ushort(h2[0]) | ((uint32_t)__half_as_ushort(h2[1]) << 16),
                                    (uint32_t)__half_as_ushort(h2[2]) | ((uint32_t)__half_as_ushort(h2[3]) << 16));
                            }
                        } else {
                            const float vv[4] = {o.x, o.y, o.z, o.w};
                            for (int u = 0; u < 4; ++u) if (n + u < N) {
                                if (Cs) Cs[off + u] = vv[u];
                                if (C_h1) { __half a, bh; int ov = 0; split_half(vv[u], a, bh, &ov); if (ov) atomicExch(overflow, 1); C_h1[off + u] = a; C_h2[off + u] = bh; }
                            }
                        }
                    }
                }
                __syncwarp();
                if (tr) gemm_trace(10 + 2 * pass);
            }
        }
        if (tr) gemm_trace(5);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
    if (threadIdx.x == 0) gemm_trace(6);
}


// Finishes a split-K GEMM: out = act((sum_s part[s]) * w_unscale + bias), slices summed in index order
// (deterministic), written as fp32 and/or as the half split the next GEMM consumes.
template <bool GELU>
__global__ void __launch_bounds__(256) umma_splitk_finish_kernel(int64_t M, int N, int ldc, int k_slices, int64_t slice_stride,
                                                                 const float* __restrict__ part, const float* __restrict__ bias,
                                                                 float w_unscale, float* __restrict__ C, __half* __restrict__ C_h1,
                                                                 __half* __restrict__ C_h2, int* __restrict__ overflow) {
    const int64_t total = M * (int64_t)(ldc / 4);
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = e / (ldc / 4);
        const int n = (int)(e % (ldc / 4)) * 4;
        if (n >= N) continue;
        const int64_t off = row * ldc + n;
        float4 acc = *reinterpret_cast<const float4*>(part + off);
        for (int sl = 1; sl < k_slices; ++sl) {
            const float4 p = *reinterpret_cast<const float4*>(part + sl * slice_stride + off);
            acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
        }
        float v[4] = {acc.x, acc.y, acc.z, acc.w};
        int ov = 0;
        for (int u = 0; u < 4; ++u) {
            if (n + u >= N) continue;
            float x = v[u] * w_unscale + (bias ? bias[n + u] : 0.f);
            if (GELU) x = gelu_erf_u(x);
            if (C) C[off + u] = x;
            if (C_h1) { __half a, b; split_half(x, a, b, &ov); C_h1[off + u] = a; C_h2[off + u] = b; }
        }
        if (ov) atomicExch(overflow, 1);
    }
}

}  // namespace sealb200
