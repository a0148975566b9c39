// CUDA kernels + C ABI (include/sealfm.h) of the FM-index path.  sm_100a only.
#include "../../include/sealfm.h"
#include "fm_device.cuh"
#include "fm_host.hpp"
#include "fm_layout.hpp"
#include "fm_expand.cuh"
#include "fm_handle.hpp"
#include "common.cuh"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <memory>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

using namespace sealb200;

// ------------------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------------------
struct sealfm {
    HostIndex host;
    int device = -1;
    FmView view{};                 // device pointers
    void* d_blocks = nullptr;
    uint64_t* d_csym = nullptr;
    NodeEntry* d_node_tab = nullptr;
    uint64_t* d_sa = nullptr;
    uint64_t* d_isa = nullptr;
    uint64_t* d_beginnings = nullptr;
    uint64_t device_bytes = 0;
    std::vector<uint64_t> beginnings;
};

namespace {

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------

// Batched LF step (FMIndex::backward_search_step, fm_index.cpp:67-76): one thread per triple.
__global__ void __launch_bounds__(256) lf_step_kernel(FmView v, uint64_t n, const uint64_t* __restrict__ sym,
                                                      const uint64_t* __restrict__ lo,
                                                      const uint64_t* __restrict__ hi,
                                                      uint64_t* __restrict__ out_lo,
                                                      uint64_t* __restrict__ out_hi) {
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < n; t += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t l, r;
        lf_step(v, sym[t], lo[t], hi[t], l, r);
        out_lo[t] = l;
        out_hi[t] = r;
    }
}

// FMIndex::backward_search_multi (fm_index.cpp:55-65): fold from (0, size()), return {l, r+1}.
__global__ void __launch_bounds__(128) lf_fold_kernel(FmView v, uint64_t nq, const uint64_t* __restrict__ symbols,
                                                      const uint64_t* __restrict__ offsets,
                                                      uint64_t* __restrict__ out_lo,
                                                      uint64_t* __restrict__ out_hi) {
    for (uint64_t q = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; q < nq; q += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t l = 0, r = v.m;
        for (uint64_t t = offsets[q]; t < offsets[q + 1]; ++t) lf_step(v, symbols[t], l, r, l, r);
        out_lo[q] = l;
        out_hi[q] = r + 1;
    }
}

constexpr int kExpandWarps = 4;
constexpr int kWideThreads = 256;

// dynamic shared memory of the two kernel shapes for a tree of height L
size_t narrow_smem(uint32_t L) { return kExpandWarps * warp_expand_smem(L); }
size_t wide_smem(uint32_t L) { return sizeof(BlockFrontier) + 2ull * L * kWideThreads * 8; }

__device__ __forceinline__ WarpFrontier& warp_frontier(unsigned char* smem, uint32_t L, uint32_t warp, uint64_t*& stk) {
    unsigned char* base = smem + warp * warp_expand_smem(L);
    stk = reinterpret_cast<uint64_t*>(base + sizeof(WarpFrontier));
    return *reinterpret_cast<WarpFrontier*>(base);
}

// Allowed-token bitmask rows for R ranges (seal/beam_search.py:107,131-135).  mask is zeroed here.
// wide_list[0] = number of wide rows found, [1] = work cursor of the wide kernel, [2..] = their indices
// (nullptr: expand everything here)
__global__ void __launch_bounds__(kExpandWarps * 32) expand_mask_kernel(
    FmView v, uint64_t R, const uint64_t* __restrict__ lo, const uint64_t* __restrict__ hi,
    uint32_t* __restrict__ mask, uint32_t ld_words, uint32_t vocab, uint32_t shift, unsigned long long* wide_list) {
    extern __shared__ __align__(16) unsigned char expand_smem[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t* stk;
    WarpFrontier& F = warp_frontier(expand_smem, v.L, warp, stk);
    for (uint64_t r = blockIdx.x * (uint64_t)kExpandWarps + warp; r < R; r += (uint64_t)gridDim.x * kExpandWarps) {
        uint32_t* row = mask + r * ld_words;
        for (uint32_t w = lane; w < ld_words; w += 32) row[w] = 0;
        __syncwarp();
        const uint64_t l = lo[r], h = hi[r];
        if (wide_list && h > l && h - l >= kWideRange) {       // defer to the block-cooperative kernel
            if (lane == 0) { const unsigned long long k = atomicAdd(wide_list, 1ULL); wide_list[2 + k] = r; }
            continue;
        }
        MaskSink sink{row, vocab, shift};
        warp_expand(v, l, h, sink, F, stk);
        __syncwarp();
    }
}

// Persistent CTAs pull the wide rows found by expand_mask_kernel (device-side list, no host sync).
__global__ void __launch_bounds__(kWideThreads) expand_mask_wide_kernel(
    FmView v, const uint64_t* __restrict__ lo, const uint64_t* __restrict__ hi, uint32_t* __restrict__ mask,
    uint32_t ld_words, uint32_t vocab, uint32_t shift, unsigned long long* wide_list) {
    extern __shared__ __align__(16) unsigned char wide_smem_raw[];
    BlockFrontier& F = *reinterpret_cast<BlockFrontier*>(wide_smem_raw);
    uint64_t* stk = reinterpret_cast<uint64_t*>(wide_smem_raw + sizeof(BlockFrontier));
    __shared__ unsigned long long pick;
    const unsigned long long n = wide_list[0];
    for (;;) {
        if (threadIdx.x == 0) pick = atomicAdd(wide_list + 1, 1ULL);
        __syncthreads();
        const unsigned long long k = pick;
        __syncthreads();
        if (k >= n) break;
        const uint64_t r = wide_list[2 + k];
        MaskSink sink{mask + r * ld_words, vocab, shift};
        block_expand(v, lo[r], hi[r], sink, F, stk, stk + (size_t)v.L * kWideThreads);
    }
}

// Dense per-range symbol counts (scratch for the ordered (symbol,count) API output).
__global__ void __launch_bounds__(kExpandWarps * 32) expand_dense_kernel(
    FmView v, uint64_t R, const uint64_t* __restrict__ lo, const uint64_t* __restrict__ hi,
    uint64_t* __restrict__ dense, unsigned long long* wide_list) {
    extern __shared__ __align__(16) unsigned char expand_smem[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t* stk;
    WarpFrontier& F = warp_frontier(expand_smem, v.L, warp, stk);
    const uint64_t stride = 1ULL << v.L;
    for (uint64_t r = blockIdx.x * (uint64_t)kExpandWarps + warp; r < R; r += (uint64_t)gridDim.x * kExpandWarps) {
        const uint64_t l = lo[r], h = hi[r];
        if (wide_list && h > l && h - l >= kWideRange) {
            if (lane == 0) { const unsigned long long k = atomicAdd(wide_list, 1ULL); wide_list[2 + k] = r; }
            continue;
        }
        DenseSink sink{dense + r * stride};
        warp_expand(v, l, h, sink, F, stk);
        __syncwarp();
    }
}

__global__ void __launch_bounds__(kWideThreads) expand_dense_wide_kernel(
    FmView v, const uint64_t* __restrict__ lo, const uint64_t* __restrict__ hi, uint64_t* __restrict__ dense,
    unsigned long long* wide_list) {
    extern __shared__ __align__(16) unsigned char wide_smem_raw[];
    BlockFrontier& F = *reinterpret_cast<BlockFrontier*>(wide_smem_raw);
    uint64_t* stk = reinterpret_cast<uint64_t*>(wide_smem_raw + sizeof(BlockFrontier));
    __shared__ unsigned long long pick;
    const unsigned long long n = wide_list[0];
    const uint64_t stride = 1ULL << v.L;
    for (;;) {
        if (threadIdx.x == 0) pick = atomicAdd(wide_list + 1, 1ULL);
        __syncthreads();
        const unsigned long long k = pick;
        __syncthreads();
        if (k >= n) break;
        const uint64_t r = wide_list[2 + k];
        DenseSink sink{dense + r * stride};
        block_expand(v, lo[r], hi[r], sink, F, stk, stk + (size_t)v.L * kWideThreads);
    }
}

// Ordered compaction of one dense count row into interleaved (symbol,count) pairs — one block per
// range; out_off[r] is where range r's pairs start (in u64 units), out_len[r] receives 2k.
__global__ void __launch_bounds__(256) compact_pairs_kernel(uint32_t L, const uint64_t* __restrict__ dense,
                                                            const uint64_t* __restrict__ out_off,
                                                            uint64_t* __restrict__ out,
                                                            uint64_t* __restrict__ out_len) {
    __shared__ uint32_t warp_tot[8];
    __shared__ uint32_t carry;
    const uint64_t r = blockIdx.x;
    const uint64_t nsym = 1ULL << L;
    const uint64_t* row = dense + r * nsym;
    uint64_t* dst = out + out_off[r];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint64_t basei = 0; basei < nsym; basei += blockDim.x) {
        const uint64_t s = basei + threadIdx.x;
        const uint64_t c = s < nsym ? row[s] : 0;
        const uint32_t flag = c != 0;
        const uint32_t ball = __ballot_sync(0xffffffffu, flag);
        const uint32_t pre = __popc(ball & ((1u << lane) - 1));
        if (lane == 0) warp_tot[warp] = __popc(ball);
        __syncthreads();
        uint32_t woff = 0, tot = 0;
        for (uint32_t w = 0; w < blockDim.x / 32; ++w) { if (w < warp) woff += warp_tot[w]; tot += warp_tot[w]; }
        const uint32_t pos = carry + woff + pre;
        if (flag) { dst[2ULL * pos] = s; dst[2ULL * pos + 1] = c; }
        __syncthreads();
        if (threadIdx.x == 0) carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) out_len[r] = 2ULL * carry;
}

__global__ void __launch_bounds__(128) locate_kernel(FmView v, uint64_t n, const uint64_t* __restrict__ rows,
                                                     uint64_t* __restrict__ out, int want_doc) {
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < n; t += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t pos = locate_row(v, rows[t]);
        if (want_doc) pos = doc_of_pos(v, pos);
        out[t] = pos;
    }
}

__global__ void __launch_bounds__(64) extract_kernel(FmView v, uint64_t n, const uint64_t* __restrict__ begins,
                                                     const uint64_t* __restrict__ ends,
                                                     const uint64_t* __restrict__ offs, uint64_t* __restrict__ out) {
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < n; t += (uint64_t)gridDim.x * blockDim.x)
        extract_text(v, begins[t], ends[t], out + offs[t]);
}

// ------------------------------------------------------------------------------------------------
// host helpers
// ------------------------------------------------------------------------------------------------
int grid_for(uint64_t work_items, int per_block, int max_waves = 8) {
    uint64_t blocks = (work_items + per_block - 1) / per_block;
    uint64_t cap = (uint64_t)sm_count() * max_waves;       // multiples of the SM count (148 on B200)
    if (blocks > cap) blocks = cap;
    if (blocks == 0) blocks = 1;
    return (int)blocks;
}

template <typename T>
struct DevBuf {
    T* p = nullptr;
    explicit DevBuf(uint64_t n) { if (n) CUDA_CHECK(cudaMalloc(&p, n * sizeof(T))); }
    ~DevBuf() { if (p) cudaFree(p); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
};

void require_device(const sealfm_t* h) {
    if (!h) throw ApiError(SEALFM_EINVAL, "null handle");
    if (h->device < 0) throw ApiError(SEALFM_ENODEVICE, "index not bound to a CUDA device (call sealfm_to_device)");
    CUDA_CHECK(cudaSetDevice(h->device));
}

void upload(sealfm_t* h, int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        throw ApiError(SEALFM_ENODEVICE, std::string("no CUDA device available: ") + cudaGetErrorString(e));
    }
    if (device < 0 || device >= count) throw ApiError(SEALFM_EINVAL, "bad device id");
    CUDA_CHECK(cudaSetDevice(device));
    const HostIndex& H = h->host;
    const uint32_t L = H.max_level;
    const uint64_t m = H.size;
    DeviceArrays A;
    make_device_arrays(H, A);
    const std::vector<uint64_t>& blk = A.blocks;
    const std::vector<uint64_t>& csym = A.csym;
    const std::vector<NodeEntry>& node_tab = A.node_tab;

    auto put = [&](const void* src, uint64_t bytes) -> void* {
        void* d = nullptr;
        CUDA_CHECK(cudaMalloc(&d, bytes ? bytes : 8));
        if (bytes) CUDA_CHECK(cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice));
        h->device_bytes += bytes;
        return d;
    };
    h->device_bytes = 0;
    h->d_blocks = put(blk.data(), blk.size() * 8);
    h->d_csym = (uint64_t*)put(csym.data(), csym.size() * 8);
    h->d_node_tab = (NodeEntry*)put(node_tab.data(), node_tab.size() * sizeof(NodeEntry));
    h->d_sa = (uint64_t*)put(H.sa_samples.data(), H.sa_samples.size() * 8);
    h->d_isa = (uint64_t*)put(H.isa_samples.data(), H.isa_samples.size() * 8);
    FmView& v = h->view;
    v.blocks = (const uint4*)h->d_blocks;
    v.csym = h->d_csym; v.node_tab = h->d_node_tab;
    v.sa_samples = h->d_sa; v.isa_samples = h->d_isa;
    v.n_isa = H.isa_samples.size();
    v.beginnings = nullptr; v.n_beginnings = 0;
    v.m = m; v.L = L;
    h->device = device;
    if (!h->beginnings.empty()) {
        h->d_beginnings = (uint64_t*)put(h->beginnings.data(), h->beginnings.size() * 8);
        v.beginnings = h->d_beginnings; v.n_beginnings = h->beginnings.size();
    }
}

void release_device(sealfm_t* h) {
    if (h->device < 0) return;
    cudaSetDevice(h->device);
    cudaFree(h->d_blocks); cudaFree(h->d_csym); cudaFree(h->d_node_tab);
    cudaFree(h->d_sa); cudaFree(h->d_isa); cudaFree(h->d_beginnings);
    h->device = -1;
}

}  // namespace

namespace sealb200 {
// Bitmask rows of R SA ranges: narrow ranges by one warp each, wide ones (>= kWideRange rows) by whole CTAs pulling
// from a device-side work list.  `wide`: R + 2 u64 of scratch.  Stream-ordered, no host synchronisation; also the
// tail of every decode step (decode.cu).
void launch_expand_masks(const FmView& v, cudaStream_t s, uint64_t R, const uint64_t* lo_d, const uint64_t* hi_d, uint32_t* mask_d,
                         uint32_t ld_words, uint32_t vocab, uint32_t shift, unsigned long long* wide) {
    if (v.L > kMaxLevels) throw ApiError(SEALFM_EINVAL, "wavelet tree higher than kMaxLevels");
    CUDA_CHECK(cudaMemsetAsync(wide, 0, 2 * sizeof(unsigned long long), s));
    const int ns = (int)narrow_smem(v.L), ws = (int)wide_smem(v.L);
    static int ns_set = 0, ws_set = 0;
    if (ns > ns_set) { CUDA_CHECK(cudaFuncSetAttribute(expand_mask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ns)); ns_set = ns; }
    if (ws > ws_set) { CUDA_CHECK(cudaFuncSetAttribute(expand_mask_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ws)); ws_set = ws; }
    expand_mask_kernel<<<grid_for(R, kExpandWarps, 16), kExpandWarps * 32, ns, s>>>(v, R, lo_d, hi_d, mask_d, ld_words, vocab, shift, wide);
    CUDA_CHECK(cudaGetLastError());
    const int wide_ctas = (int)std::min<uint64_t>(R, (uint64_t)sm_count() * 2);
    expand_mask_wide_kernel<<<wide_ctas, kWideThreads, ws, s>>>(v, lo_d, hi_d, mask_d, ld_words, vocab, shift, wide);
    CUDA_CHECK(cudaGetLastError());
}

FmView sealfm_view(const sealfm_t* h) {
    if (!h) throw ApiError(SEALFM_EINVAL, "null handle");
    if (h->device < 0) throw ApiError(SEALFM_ENODEVICE, "index not bound to a CUDA device (call sealfm_to_device)");
    return h->view;
}
}  // namespace sealb200

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

const char* sealfm_last_error(void) { return last_error().c_str(); }
int sealfm_abi_version(void) { return 1; }

int sealfm_build(const uint64_t* symbols, uint64_t n, sealfm_t** out) {
    return guarded([&] {
        if (!out || (!symbols && n)) throw ApiError(SEALFM_EINVAL, "null argument");
        std::unique_ptr<sealfm> h(new sealfm());
        build_index(symbols, n, h->host);
        *out = h.release();
    });
}
int sealfm_build_gpu(const uint64_t* symbols, uint64_t n, int device, sealfm_t** out) {
    return guarded([&] {
        if (!out || (!symbols && n)) throw ApiError(SEALFM_EINVAL, "null argument");
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
            cudaGetLastError();
            throw ApiError(SEALFM_ENODEVICE, "no such CUDA device");
        }
        std::unique_ptr<sealfm> h(new sealfm());
        build_index_gpu(symbols, n, device, h->host);
        *out = h.release();
    });
}
int sealfm_build_from_file(const char* path, int width_bytes, sealfm_t** out) {
    return guarded([&] {
        if (!out || !path) throw ApiError(SEALFM_EINVAL, "null argument");
        std::unique_ptr<sealfm> h(new sealfm());
        try { build_index_from_file(path, width_bytes, h->host); }
        catch (const std::runtime_error& e) { throw ApiError(SEALFM_EIO, e.what()); }
        *out = h.release();
    });
}
int sealfm_load(const char* path, sealfm_t** out) {
    return guarded([&] {
        if (!out || !path) throw ApiError(SEALFM_EINVAL, "null argument");
        std::unique_ptr<sealfm> h(new sealfm());
        try { load_index(path, h->host); }
        catch (const std::runtime_error& e) { throw ApiError(SEALFM_EIO, e.what()); }
        *out = h.release();
    });
}
int sealfm_save(const sealfm_t* h, const char* path) {
    return guarded([&] {
        if (!h || !path) throw ApiError(SEALFM_EINVAL, "null argument");
        try { save_index_native(h->host, path); }
        catch (const std::runtime_error& e) { throw ApiError(SEALFM_EIO, e.what()); }
    });
}
int sealfm_save_sdsl(const sealfm_t* h, const char* path) {
    return guarded([&] {
        if (!h || !path) throw ApiError(SEALFM_EINVAL, "null argument");
        try { save_index_sdsl(h->host, path); }
        catch (const std::runtime_error& e) { throw ApiError(SEALFM_EIO, e.what()); }
    });
}
void sealfm_free(sealfm_t* h) {
    if (!h) return;
    release_device(h);
    delete h;
}
uint64_t sealfm_size(const sealfm_t* h) { return h ? h->host.size : 0; }
uint64_t sealfm_sigma(const sealfm_t* h) { return h ? h->host.sigma : 0; }
uint32_t sealfm_max_level(const sealfm_t* h) { return h ? h->host.max_level : 0; }

int sealfm_section(const sealfm_t* h, int which, const uint64_t** ptr, uint64_t* n_words) {
    return guarded([&] {
        if (!h || !ptr || !n_words) throw ApiError(SEALFM_EINVAL, "null argument");
        const std::vector<uint64_t>* v = nullptr;
        switch (which) {
            case 0: v = &h->host.tree; break;
            case 1: v = &h->host.alphabet; break;
            case 2: v = &h->host.C; break;
            case 3: v = &h->host.sa_samples; break;
            case 4: v = &h->host.isa_samples; break;
            default: throw ApiError(SEALFM_EINVAL, "unknown section");
        }
        *ptr = v->data(); *n_words = v->size();
    });
}

int sealfm_to_device(sealfm_t* h, int device) {
    return guarded([&] {
        if (!h) throw ApiError(SEALFM_EINVAL, "null handle");
        if (h->device >= 0) release_device(h);
        upload(h, device);
    });
}
int sealfm_device(const sealfm_t* h) { return h ? h->device : -1; }
uint64_t sealfm_device_bytes(const sealfm_t* h) { return h ? h->device_bytes : 0; }

int sealfm_set_beginnings(sealfm_t* h, const uint64_t* beginnings, uint64_t n) {
    return guarded([&] {
        if (!h || (!beginnings && n)) throw ApiError(SEALFM_EINVAL, "null argument");
        h->beginnings.assign(beginnings, beginnings + n);
        if (h->device >= 0) {
            CUDA_CHECK(cudaSetDevice(h->device));
            if (h->d_beginnings) { cudaFree(h->d_beginnings); h->d_beginnings = nullptr; }
            CUDA_CHECK(cudaMalloc(&h->d_beginnings, (n ? n : 1) * 8));
            CUDA_CHECK(cudaMemcpy(h->d_beginnings, beginnings, n * 8, cudaMemcpyHostToDevice));
            h->view.beginnings = h->d_beginnings; h->view.n_beginnings = n;
        }
    });
}

int sealfm_backward_search_step_d(const sealfm_t* h, sealfm_stream_t stream, uint64_t n,
                                  const uint64_t* sym_d, const uint64_t* lo_d, const uint64_t* hi_d,
                                  uint64_t* out_lo_d, uint64_t* out_hi_d) {
    return guarded([&] {
        require_device(h);
        if (!n) return;
        lf_step_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(h->view, n, sym_d, lo_d, hi_d, out_lo_d, out_hi_d);
        CUDA_CHECK(cudaGetLastError());
    });
}

int sealfm_expand_mask_d(const sealfm_t* h, sealfm_stream_t stream, uint64_t R, const uint64_t* lo_d,
                         const uint64_t* hi_d, uint32_t* mask_d, uint32_t ld_words, uint32_t vocab,
                         uint32_t shift) {
    return guarded([&] {
        require_device(h);
        if (!R) return;
        if ((uint64_t)ld_words * 32 < vocab) throw ApiError(SEALFM_EINVAL, "ld_words too small for vocab");
        cudaStream_t s = (cudaStream_t)stream;
        unsigned long long* wide = nullptr;                    // [count, cursor, rows...]
        CUDA_CHECK(cudaMallocAsync(&wide, (R + 2) * sizeof(unsigned long long), s));
        struct Free { unsigned long long* p; cudaStream_t s; ~Free() { cudaFreeAsync(p, s); } } guard{wide, s};
        launch_expand_masks(h->view, s, R, lo_d, hi_d, mask_d, ld_words, vocab, shift, wide);
    });
}

int sealfm_backward_search_step(const sealfm_t* h, uint64_t n, const uint64_t* sym, const uint64_t* lo,
                                const uint64_t* hi, uint64_t* out_lo, uint64_t* out_hi) {
    return guarded([&] {
        require_device(h);
        if (!n) return;
        DevBuf<uint64_t> d(5 * n);
        CUDA_CHECK(cudaMemcpy(d.p, sym, n * 8, cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemcpy(d.p + n, lo, n * 8, cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemcpy(d.p + 2 * n, hi, n * 8, cudaMemcpyHostToDevice));
        lf_step_kernel<<<grid_for(n, 256), 256>>>(h->view, n, d.p, d.p + n, d.p + 2 * n, d.p + 3 * n, d.p + 4 * n);
        CUDA_CHECK(cudaGetLastError());
        CUDA_CHECK(cudaMemcpy(out_lo, d.p + 3 * n, n * 8, cudaMemcpyDeviceToHost));
        CUDA_CHECK(cudaMemcpy(out_hi, d.p + 4 * n, n * 8, cudaMemcpyDeviceToHost));
    });
}

int sealfm_backward_search_multi(const sealfm_t* h, uint64_t nq, const uint64_t* symbols,
                                 const uint64_t* offsets, uint64_t* out_lo, uint64_t* out_hi) {
    return guarded([&] {
        require_device(h);
        if (!nq) return;
        const uint64_t tot = offsets[nq];
        DevBuf<uint64_t> ds(tot), doff(nq + 1), dout(2 * nq);
        if (tot) CUDA_CHECK(cudaMemcpy(ds.p, symbols, tot * 8, cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemcpy(doff.p, offsets, (nq + 1) * 8, cudaMemcpyHostToDevice));
        lf_fold_kernel<<<grid_for(nq, 128), 128>>>(h->view, nq, ds.p, doff.p, dout.p, dout.p + nq);
        CUDA_CHECK(cudaGetLastError());
        CUDA_CHECK(cudaMemcpy(out_lo, dout.p, nq * 8, cudaMemcpyDeviceToHost));
        CUDA_CHECK(cudaMemcpy(out_hi, dout.p + nq, nq * 8, cudaMemcpyDeviceToHost));
    });
}

int sealfm_distinct_count_multi(const sealfm_t* h, uint64_t n, const uint64_t* lows, const uint64_t* highs,
                                uint64_t* out_offsets, uint64_t* out, uint64_t out_cap) {
    return guarded([&] {
        require_device(h);
        if (!out_offsets || (!lows && n) || (!highs && n)) throw ApiError(SEALFM_EINVAL, "null argument");
        const uint32_t L = h->host.max_level;
        const uint64_t nsym = 1ULL << L;
        const uint64_t chunk = std::max<uint64_t>(1, std::min<uint64_t>(64, (1ULL << 25) / (nsym * 8)));
        std::vector<uint64_t> lens(n, 0);
        std::vector<std::vector<uint64_t>> pieces;      // per chunk: packed results (upper-bound layout)
        std::vector<std::vector<uint64_t>> piece_off;
        pieces.reserve((n + chunk - 1) / chunk);
        for (uint64_t c0 = 0; c0 < n; c0 += chunk) {
            const uint64_t cn = std::min(chunk, n - c0);
            std::vector<uint64_t> ub(cn + 1, 0);        // upper bound on 2k per range
            for (uint64_t i = 0; i < cn; ++i) {
                uint64_t lo = lows[c0 + i], hi = highs[c0 + i];
                // hi == size()+1 is reachable through the reference's first-step quirk;
                // the arithmetic below is the reference's own for such a range.
                if (hi > h->host.size + 1) throw ApiError(SEALFM_EINVAL, "range end beyond size()+1");
                uint64_t k = hi > lo ? std::min<uint64_t>(hi - lo, nsym) : 0;
                ub[i + 1] = ub[i] + 2 * k;
            }
            DevBuf<uint64_t> dlo(cn), dhi(cn), ddense(cn * nsym), doff(cn + 1), dout(ub[cn]), dlen(cn);
            CUDA_CHECK(cudaMemcpy(dlo.p, lows + c0, cn * 8, cudaMemcpyHostToDevice));
            CUDA_CHECK(cudaMemcpy(dhi.p, highs + c0, cn * 8, cudaMemcpyHostToDevice));
            CUDA_CHECK(cudaMemcpy(doff.p, ub.data(), (cn + 1) * 8, cudaMemcpyHostToDevice));
            CUDA_CHECK(cudaMemset(ddense.p, 0, cn * nsym * 8));
            DevBuf<unsigned long long> dwide(cn + 2);
            CUDA_CHECK(cudaMemset(dwide.p, 0, (cn + 2) * sizeof(unsigned long long)));
            const int ns = (int)narrow_smem(L), ws = (int)wide_smem(L);
            CUDA_CHECK(cudaFuncSetAttribute(expand_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ns));
            expand_dense_kernel<<<grid_for(cn, kExpandWarps, 16), kExpandWarps * 32, ns>>>(h->view, cn, dlo.p, dhi.p, ddense.p, dwide.p);
            CUDA_CHECK(cudaGetLastError());
            CUDA_CHECK(cudaFuncSetAttribute(expand_dense_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ws));
            expand_dense_wide_kernel<<<sm_count() * 2, kWideThreads, ws>>>(h->view, dlo.p, dhi.p, ddense.p, dwide.p);
            CUDA_CHECK(cudaGetLastError());
            compact_pairs_kernel<<<(unsigned)cn, 256>>>(L, ddense.p, doff.p, dout.p, dlen.p);
            CUDA_CHECK(cudaGetLastError());
            pieces.emplace_back(ub[cn]);
            if (ub[cn]) CUDA_CHECK(cudaMemcpy(pieces.back().data(), dout.p, ub[cn] * 8, cudaMemcpyDeviceToHost));
            CUDA_CHECK(cudaMemcpy(lens.data() + c0, dlen.p, cn * 8, cudaMemcpyDeviceToHost));
            piece_off.push_back(std::move(ub));
        }
        out_offsets[0] = 0;
        for (uint64_t i = 0; i < n; ++i) out_offsets[i + 1] = out_offsets[i] + lens[i];
        if (!out) return;
        if (out_offsets[n] > out_cap) throw ApiError(SEALFM_ECAPACITY, "output buffer too small");
        for (uint64_t c0 = 0, pc = 0; c0 < n; c0 += chunk, ++pc) {
            const uint64_t cn = std::min(chunk, n - c0);
            for (uint64_t i = 0; i < cn; ++i)
                if (lens[c0 + i])
                    std::memcpy(out + out_offsets[c0 + i], pieces[pc].data() + piece_off[pc][i], lens[c0 + i] * 8);
        }
    });
}

static int locate_impl(const sealfm_t* h, uint64_t n, const uint64_t* rows, uint64_t* out, int want_doc) {
    return guarded([&] {
        require_device(h);
        if (!n) return;
        if (want_doc && !h->view.beginnings) throw ApiError(SEALFM_EINVAL, "sealfm_set_beginnings not called");
        DevBuf<uint64_t> d(2 * n);
        CUDA_CHECK(cudaMemcpy(d.p, rows, n * 8, cudaMemcpyHostToDevice));
        locate_kernel<<<grid_for(n, 128), 128>>>(h->view, n, d.p, d.p + n, want_doc);
        CUDA_CHECK(cudaGetLastError());
        CUDA_CHECK(cudaMemcpy(out, d.p + n, n * 8, cudaMemcpyDeviceToHost));
    });
}
int sealfm_locate(const sealfm_t* h, uint64_t n, const uint64_t* rows, uint64_t* out_pos) {
    return locate_impl(h, n, rows, out_pos, 0);
}
int sealfm_doc_index_from_rows(const sealfm_t* h, uint64_t n, const uint64_t* rows, uint64_t* out_doc) {
    return locate_impl(h, n, rows, out_doc, 1);
}

int sealfm_extract_text(const sealfm_t* h, uint64_t n, const uint64_t* begins, const uint64_t* ends,
                        uint64_t* out_offsets, uint64_t* out, uint64_t out_cap) {
    return guarded([&] {
        require_device(h);
        if (!out_offsets) throw ApiError(SEALFM_EINVAL, "null argument");
        out_offsets[0] = 0;
        for (uint64_t i = 0; i < n; ++i) {
            if (ends[i] >= h->host.size || begins[i] > ends[i]) throw ApiError(SEALFM_EINVAL, "bad text interval");
            out_offsets[i + 1] = out_offsets[i] + (ends[i] - begins[i]);
        }
        if (!out || !n) return;
        const uint64_t tot = out_offsets[n];
        if (tot > out_cap) throw ApiError(SEALFM_ECAPACITY, "output buffer too small");
        DevBuf<uint64_t> db(n), de(n), doff(n + 1), dout(tot);
        CUDA_CHECK(cudaMemcpy(db.p, begins, n * 8, cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemcpy(de.p, ends, n * 8, cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemcpy(doff.p, out_offsets, (n + 1) * 8, cudaMemcpyHostToDevice));
        extract_kernel<<<grid_for(n, 64), 64>>>(h->view, n, db.p, de.p, doff.p, dout.p);
        CUDA_CHECK(cudaGetLastError());
        if (tot) CUDA_CHECK(cudaMemcpy(out, dout.p, tot * 8, cudaMemcpyDeviceToHost));
    });
}

}  // extern "C"
