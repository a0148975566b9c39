// BART-large building blocks as hand-written CUDA kernels (fp32 arithmetic, matching the
// reference's eager fp32 forward: transformers BartForConditionalGeneration, call sites
// seal/beam_search.py:231-238,481-483).  Post-LN encoder/decoder layers, learned positions with
// offset 2, layernorm_embedding, exact-erf GELU, tied lm_head + final_logits_bias.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace sealb200 {

constexpr int kHeadDim = 64;

// ---- small helpers -----------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// LayerNorm of one row held as `per` float4 per lane (d = 128*per), torch semantics:
// biased variance, eps inside the sqrt, fp32.
template <int MAXV>
__device__ __forceinline__ void warp_layernorm(float4 (&v)[MAXV], int nv, int d, const float* __restrict__ gamma,
                                               const float* __restrict__ beta, float eps, float* __restrict__ out,
                                               int lane) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) if (i < nv) s += v[i].x + v[i].y + v[i].z + v[i].w;
    const float mean = warp_sum(s) / (float)d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) if (i < nv) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
        q += a * a + b * b + c * c + e * e;
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)d + eps);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) if (i < nv) {
        const int col = (i * 32 + lane) * 4;
        const float4 g = *reinterpret_cast<const float4*>(gamma + col);
        const float4 b = *reinterpret_cast<const float4*>(beta + col);
        float4 o;
        o.x = (v[i].x - mean) * rstd * g.x + b.x;
        o.y = (v[i].y - mean) * rstd * g.y + b.y;
        o.z = (v[i].z - mean) * rstd * g.z + b.z;
        o.w = (v[i].w - mean) * rstd * g.w + b.w;
        *reinterpret_cast<float4*>(out + col) = o;
    }
}

constexpr int kLnMaxVec = 8;     // d_model <= 1024

// out[r] = LN(embed[tok[r]] * scale + pos_table[pos(r) + 2])      (BartEncoder/BartDecoder embedding)
// tok: int32, row r reads tok[r * tok_stride].  pos(r) = pos_const if pos_per_row == nullptr else pos_per_row[r].
__global__ void __launch_bounds__(128) embed_ln_kernel(int64_t rows, int d, const int32_t* __restrict__ tok,
                                                       int64_t tok_stride,
                                                       const int32_t* __restrict__ pos_per_row, int pos_const,
                                                       const float* __restrict__ embed, float scale,
                                                       const float* __restrict__ pos_table,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t r = blockIdx.x * 4LL + (threadIdx.x >> 5);
    if (r >= rows) return;
    const int nv = d / 128;
    const float* e = embed + (int64_t)tok[r * tok_stride] * d;
    const int p = (pos_per_row ? pos_per_row[r] : pos_const) + 2;
    const float* pe = pos_table + (int64_t)p * d;
    float4 v[kLnMaxVec];
#pragma unroll
    for (int i = 0; i < kLnMaxVec; ++i) if (i < nv) {
        const int col = (i * 32 + lane) * 4;
        const float4 a = *reinterpret_cast<const float4*>(e + col);
        const float4 b = *reinterpret_cast<const float4*>(pe + col);
        v[i] = make_float4(a.x * scale + b.x, a.y * scale + b.y, a.z * scale + b.z, a.w * scale + b.w);
    }
    warp_layernorm<kLnMaxVec>(v, nv, d, gamma, beta, 1e-5f, out + r * d, lane);
}

// out[r] = LN(a[r] + b[r])     (residual + sub-layer output, post-LN)
__global__ void __launch_bounds__(128) add_ln_kernel(int64_t rows, int d, const float* __restrict__ a,
                                                     const float* __restrict__ b, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t r = blockIdx.x * 4LL + (threadIdx.x >> 5);
    if (r >= rows) return;
    const int nv = d / 128;
    float4 v[kLnMaxVec];
#pragma unroll
    for (int i = 0; i < kLnMaxVec; ++i) if (i < nv) {
        const int col = (i * 32 + lane) * 4;
        const float4 x = *reinterpret_cast<const float4*>(a + r * d + col);
        const float4 y = *reinterpret_cast<const float4*>(b + r * d + col);
        v[i] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
    }
    warp_layernorm<kLnMaxVec>(v, nv, d, gamma, beta, 1e-5f, out + r * d, lane);
}

// ---- fp32 SIMT GEMM:  C[M,N] = A[M,K] * W[N,K]^T + bias[N]  (optionally GELU) ---------------------
// A, W row-major with K contiguous (nn.Linear layout).  128x128x16 tiles, 256 threads, 8x8 per thread.
constexpr int GBM = 128, GBN = 128, GBK = 16, GTHREADS = 256;

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

template <bool GELU>
__global__ void __launch_bounds__(GTHREADS, 2) sgemm_tn_kernel(int M, int N, int K, const float* __restrict__ A, int lda,
                                                            const float* __restrict__ W, int ldw,
                                                            const float* __restrict__ bias, float* __restrict__ C,
                                                            int ldc) {
    __shared__ __align__(16) float As[2][GBK][GBM + 4];
    __shared__ __align__(16) float Bs[2][GBK][GBN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
    // global -> smem mapping: each thread loads two float4 of A and two of W per k-tile
    const int lrow = tid >> 2;            // 0..63
    const int lk = (tid & 3) * 4;         // 0,4,8,12
    float4 ra[2], rb[2];
    auto gload = [&](int k0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int am = m0 + lrow + 64 * h;
            ra[h] = am < M ? *reinterpret_cast<const float4*>(A + (int64_t)am * lda + k0 + lk) : make_float4(0, 0, 0, 0);
            const int bn = n0 + lrow + 64 * h;
            rb[h] = bn < N ? *reinterpret_cast<const float4*>(W + (int64_t)bn * ldw + k0 + lk) : make_float4(0, 0, 0, 0);
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = lrow + 64 * h;
            As[buf][lk + 0][r] = ra[h].x; As[buf][lk + 1][r] = ra[h].y; As[buf][lk + 2][r] = ra[h].z; As[buf][lk + 3][r] = ra[h].w;
            Bs[buf][lk + 0][r] = rb[h].x; Bs[buf][lk + 1][r] = rb[h].y; Bs[buf][lk + 2][r] = rb[h].z; Bs[buf][lk + 3][r] = rb[h].w;
        }
    };
    const int ty = tid >> 4, tx = tid & 15;       // 16 x 16 thread grid, each 8x8 (two 4-wide halves)
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    gload(0);
    sstore(0);
    __syncthreads();
    const int nk = K / GBK;
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) gload((kt + 1) * GBK);
#pragma unroll
        for (int k = 0; k < GBK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) sstore(buf ^ 1);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= M) continue;
#pragma unroll
        for (int jh = 0; jh < 2; ++jh) {
            const int nb = n0 + (jh ? 64 + tx * 4 : tx * 4);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = nb + j;
                if (n < N) {
                    float v = acc[i][jh * 4 + j] + (bias ? bias[n] : 0.f);
                    if (GELU) v = gelu_erf(v);
                    C[(int64_t)m * ldc + n] = v;
                }
            }
        }
    }
}

// ---- attention ---------------------------------------------------------------------------------
// One warp per (row, head); lane holds dims {2*lane, 2*lane+1} of the 64-wide head.
// Online softmax (running max / sum) in fp32; scores scaled by head_dim^-0.5 = 0.125.

struct OnlineSoftmax {
    float m = -INFINITY, l = 0.f, ax = 0.f, ay = 0.f;
    __device__ __forceinline__ void push(float s, float vx, float vy) {
        const float mn = fmaxf(m, s);
        const float c = (m == -INFINITY) ? 0.f : expf(m - mn);
        const float p = expf(s - mn);
        l = l * c + p; ax = ax * c + p * vx; ay = ay * c + p * vy; m = mn;
    }
};

// Decoder self-attention for one new token per row with beam-ancestry indirection instead of a
// cache reorder (the reference index_selects 24 cache tensors per step, seal/beam_search.py:331-332).
// qkv: [R][3d] (q | k | v) of the current position; kc/vc: [T][R][d] per layer; anc: [R][T] source row
// of every earlier position.  Writes this position's k,v into the cache.
__global__ void __launch_bounds__(512) dec_self_attn_kernel(int64_t R, int d, int heads, int cur_pos, int T,
                                                            const float* __restrict__ qkv, float* __restrict__ kc,
                                                            float* __restrict__ vc, const int32_t* __restrict__ anc,
                                                            float* __restrict__ out) {
    const int64_t r = blockIdx.x;
    const int lane = threadIdx.x & 31;
    for (int h = threadIdx.x >> 5; h < heads; h += blockDim.x >> 5) {
        const int col = h * kHeadDim + lane * 2;
        const float2 q = *reinterpret_cast<const float2*>(qkv + r * 3 * d + col);
        const float2 k = *reinterpret_cast<const float2*>(qkv + r * 3 * d + d + col);
        const float2 v = *reinterpret_cast<const float2*>(qkv + r * 3 * d + 2 * d + col);
        *reinterpret_cast<float2*>(kc + ((int64_t)cur_pos * R + r) * d + col) = k;
        *reinterpret_cast<float2*>(vc + ((int64_t)cur_pos * R + r) * d + col) = v;
        OnlineSoftmax sm;
        for (int s = 0; s < cur_pos; ++s) {
            const int64_t src = anc[r * T + s];
            const float2 ks = *reinterpret_cast<const float2*>(kc + ((int64_t)s * R + src) * d + col);
            const float2 vs = *reinterpret_cast<const float2*>(vc + ((int64_t)s * R + src) * d + col);
            const float sc = warp_sum(q.x * ks.x + q.y * ks.y) * 0.125f;
            sm.push(sc, vs.x, vs.y);
        }
        const float sc = warp_sum(q.x * k.x + q.y * k.y) * 0.125f;
        sm.push(sc, v.x, v.y);
        *reinterpret_cast<float2*>(out + r * d + col) = make_float2(sm.ax / sm.l, sm.ay / sm.l);
    }
}

// Cross attention: q [R][d]; ckv [Q*S][2d] (k | v) of the encoder states; row r belongs to query
// r / beams.  Padded source positions (mask == 0) are excluded.
__global__ void __launch_bounds__(512) cross_attn_kernel(int64_t R, int d, int heads, int beams, int S,
                                                         const float* __restrict__ q, const float* __restrict__ ckv,
                                                         const int32_t* __restrict__ src_mask, float* __restrict__ out) {
    const int64_t r = blockIdx.x;
    const int64_t qi = r / beams;
    const int lane = threadIdx.x & 31;
    for (int h = threadIdx.x >> 5; h < heads; h += blockDim.x >> 5) {
        const int col = h * kHeadDim + lane * 2;
        const float2 qq = *reinterpret_cast<const float2*>(q + r * d + col);
        OnlineSoftmax sm;
        for (int s = 0; s < S; ++s) {
            if (!src_mask[qi * S + s]) continue;
            const float* base = ckv + (qi * S + s) * 2 * d;
            const float2 ks = *reinterpret_cast<const float2*>(base + col);
            const float2 vs = *reinterpret_cast<const float2*>(base + d + col);
            sm.push(warp_sum(qq.x * ks.x + qq.y * ks.y) * 0.125f, vs.x, vs.y);
        }
        *reinterpret_cast<float2*>(out + r * d + col) = make_float2(sm.ax / sm.l, sm.ay / sm.l);
    }
}

// Encoder self attention over the S positions of the same query (bidirectional, key padding mask).
// qkv [Q*S][3d].
__global__ void __launch_bounds__(512) enc_self_attn_kernel(int64_t tokens, int d, int heads, int S,
                                                            const float* __restrict__ qkv,
                                                            const int32_t* __restrict__ src_mask,
                                                            float* __restrict__ out) {
    const int64_t t = blockIdx.x;
    const int64_t qi = t / S;
    const int lane = threadIdx.x & 31;
    for (int h = threadIdx.x >> 5; h < heads; h += blockDim.x >> 5) {
        const int col = h * kHeadDim + lane * 2;
        const float2 qq = *reinterpret_cast<const float2*>(qkv + t * 3 * d + col);
        OnlineSoftmax sm;
        for (int s = 0; s < S; ++s) {
            if (!src_mask[qi * S + s]) continue;
            const float* base = qkv + (qi * S + s) * 3 * d;
            const float2 ks = *reinterpret_cast<const float2*>(base + d + col);
            const float2 vs = *reinterpret_cast<const float2*>(base + 2 * d + col);
            sm.push(warp_sum(qq.x * ks.x + qq.y * ks.y) * 0.125f, vs.x, vs.y);
        }
        *reinterpret_cast<float2*>(out + t * d + col) = make_float2(sm.ax / sm.l, sm.ay / sm.l);
    }
}

}  // namespace sealb200
