// 3xFP16 tcgen05 GEMM for SKINNY problems (a few hundred rows: the decoder at the reference's batch of 20,
// M = 300; the last GPUs' shards under strong scaling): 128 x 64 output tiles instead of 128 x 256.
//
// Why a second tile shape.  At M = 300 a 128 x 256 tiling gives 12-48 tiles for 148 SMs, so round 1 cut every tile's
// K loop into up to 8 slices and summed them in a second kernel: two launches per linear layer, a 6 us TMEM-drain
// epilogue per 128 x 256 slice tile (profiles/r01_gemm_trace_smallM.log: 10-12 us per GEMM kernel + ~3 us finish +
// two launch gaps = 24-35 us per layer, 70 % of a batch-20 generate).  With N = 64 tiles the same problem has 48-192
// tiles: every SM gets a whole-K tile, the accumulator tile is 4x smaller (4 epilogue warps drain 64 columns), the
// operand ring is 4 stages deep (48 KB per stage) instead of 2, and bias / GELU / the half split happen in the GEMM's
// own epilogue -- ONE launch.  K is still sliced for K = 4096 (fc2), where a tile's K loop alone is 64 k-blocks.
// Same arithmetic as umma_gemm_f16x3_persistent_kernel (same 256-K TMEM chunks promoted to fp32 registers with
// round-to-nearest adds), so an unsliced tile is bit-identical to the wide kernel's.
//
// Warp roles: 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 3 idle, 4..7 epilogue (TMEM lane quarter = warp % 4).
#pragma once
#include "pdl.cuh"
#include "umma_gemm.cuh"

namespace sealb200 {

constexpr int SK_BN = 64;
constexpr int SK_NST = 4;
constexpr int SK_THREADS = 256;
constexpr int SK_AB = UM * 128, SK_WB = SK_BN * 128, SK_STAGE = 2 * SK_AB + 2 * SK_WB;      // 32 + 32... = 48 KB per stage
constexpr int SK_SMEM = SK_NST * SK_STAGE + 1024 /*alignment*/ + 256 /*barriers*/ + 4 * 2048 /*epilogue transpose*/;

template <bool GELU>
__global__ void __launch_bounds__(SK_THREADS, 1)
umma_gemm_f16x3_skinny_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                              const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                              int M, int N, int K, const float* __restrict__ bias, float w_unscale, float* __restrict__ C,
                              __half* __restrict__ C_h1, __half* __restrict__ C_h2, int ldc,
                              int* __restrict__ overflow, int k_slices, int64_t slice_stride) {
    constexpr int KE = 64;                                                 // K halves per k-block (128-byte rows)
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + SK_NST * SK_STAGE;
    const uint32_t full0 = bars, empty0 = bars + 8 * SK_NST;
    const uint32_t tfull0 = bars + 16 * SK_NST, tempty0 = tfull0 + 16;
    const uint32_t slot = tempty0 + 16;
    float* stage_base = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)) + SK_NST * SK_STAGE + 256);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = (M + UM - 1) / UM, n_tiles = (N + SK_BN - 1) / SK_BN;
    const int total_items = m_tiles * n_tiles * k_slices;
    const int num_k = (K / KE) / k_slices;
    const int num_chunks = (num_k + UKC16 - 1) / UKC16;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < SK_NST; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull0 + 8 * b, 1); mbar_init(tempty0 + 8 * b, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    } else if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(slot));
    pdl_enter();                                               // barriers / TMEM are set up while the previous kernel drains

    // work item = (m tile fastest, n tile, K slice): the CTAs running together share the W tile columns in L2
    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
                const int tile = item / k_slices, kb0 = (item % k_slices) * num_k;
                const int m_tile = tile % m_tiles, n_tile = tile / m_tiles;
                for (int kb = 0; kb < num_k; ++kb, ++it) {
                    const int s = it % SK_NST;
                    const uint32_t ph = (it / SK_NST) & 1;
                    mbar_wait(empty0 + 8 * s, ph ^ 1);
                    const uint32_t st = base + s * SK_STAGE;
                    mbar_expect_tx(full0 + 8 * s, SK_STAGE);
                    tma_load_2d(st, &tmA_hi, full0 + 8 * s, (kb0 + kb) * KE, m_tile * UM);
                    tma_load_2d(st + SK_AB, &tmA_lo, full0 + 8 * s, (kb0 + kb) * KE, m_tile * UM);
                    tma_load_2d(st + 2 * SK_AB, &tmW_hi, full0 + 8 * s, (kb0 + kb) * KE, n_tile * SK_BN);
                    tma_load_2d(st + 2 * SK_AB + SK_WB, &tmW_lo, full0 + 8 * s, (kb0 + kb) * KE, n_tile * SK_BN);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor: D=F32 (1<<4), A=B=F16, K-major both, N>>3 at bit 17, M>>4 at bit 24
            const uint32_t idesc = (1u << 4) | ((uint32_t)(SK_BN >> 3) << 17) | ((uint32_t)(UM >> 4) << 24);
            uint32_t it = 0, ch = 0;
            for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
                int kb = 0;
                for (int c = 0; c < num_chunks; ++c, ++ch) {
                    const int buf = ch & 1;
                    mbar_wait(tempty0 + 8 * buf, ((ch >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t tacc = tmem_base + (uint32_t)(buf * SK_BN);
                    const int kend = (kb + UKC16 < num_k) ? kb + UKC16 : num_k;
                    for (int k0 = kb; kb < kend; ++kb, ++it) {
                        const int s = it % SK_NST;
                        const uint32_t ph = (it / SK_NST) & 1;
                        mbar_wait(full0 + 8 * s, ph);
                        tc_fence_after();
                        const uint32_t st = base + s * SK_STAGE;
                        const uint64_t a_hi = umma_desc<128>(st), a_lo = umma_desc<128>(st + SK_AB);
                        const uint64_t w_hi = umma_desc<128>(st + 2 * SK_AB), w_lo = umma_desc<128>(st + 2 * SK_AB + SK_WB);
#pragma unroll
                        for (int k = 0; k < KE / 16; ++k) {
                            umma_f16(tacc, a_lo + 2 * k, w_hi + 2 * k, idesc, (kb != k0) || (k != 0));
                            umma_f16(tacc, a_hi + 2 * k, w_lo + 2 * k, idesc, 1);
                            umma_f16(tacc, a_hi + 2 * k, w_hi + 2 * k, idesc, 1);
                        }
                        umma_commit(empty0 + 8 * s);
                    }
                    umma_commit(tfull0 + 8 * buf);
                }
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;                                // TMEM lane quarter of this warp
        uint32_t ch = 0;
        for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
            const int tile = item / k_slices;
            float* Cs = C ? C + (int64_t)(item % k_slices) * slice_stride : nullptr;
            const int m_tile = tile % m_tiles, n_tile = tile / m_tiles;
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
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * SK_BN + h * 32), r);
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[h * 32 + j] += __uint_as_float(r[j]);      // round-to-nearest promotion
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
            }
            const int row0 = m_tile * UM + q * 32;
            const int nb = n_tile * SK_BN;
            float* stg = stage_base + (warp - 4) * 512;            // 32 x 16 floats
            if (row0 < M && nb < N) {
#pragma unroll
                for (int pass = 0; pass < 4; ++pass) {             // 16 columns per pass
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
#pragma unroll
                    for (int i = 0; i < 4; ++i) {                  // 8 rows x 64 B per instruction
                        const int rr = i * 8 + (lane >> 2), cq = lane & 3;
                        const float4 o = *reinterpret_cast<const float4*>(stg + rr * 16 + (cq ^ ((rr >> 1) & 3)) * 4);
                        const int row = row0 + rr;
                        const int n = nb + pass * 16 + cq * 4;
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
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
    }
}

}  // namespace sealb200
