// Host-side derivation of the device layout (see fm_device.cuh) from a HostIndex.  Used by the
// uploader (fm_kernels.cu) and by tests/hostcheck.cpp.
#pragma once
#include "fm_device.cuh"
#include "fm_host.hpp"

#include <vector>

namespace sealb200 {

struct DeviceArrays {
    std::vector<uint64_t> blocks;      // 4 u64 per 32-byte block
    std::vector<uint64_t> csym;        // 2^L + 1
    std::vector<NodeEntry> node_tab;   // 2^L, heap order
};

inline void make_device_arrays(const HostIndex& H, DeviceArrays& A) {
    const uint32_t L = H.max_level;
    const uint64_t m = H.size;
    const uint64_t words = H.tree.size();
    // two padding blocks: positions one past the tree must read zero bits, exactly
    // like sdsl's zeroed padding word (sdsl/memory_management.hpp:351-366)
    const uint64_t nblk = (words + 2) / 3 + 2;
    A.blocks.assign(nblk * 4, 0);
    uint64_t ones = 0;
    for (uint64_t b = 0; b < nblk; ++b) {
        A.blocks[4 * b] = ones;
        for (int s = 0; s < 3; ++s) {
            const uint64_t w = 3 * b + s;
            const uint64_t x = w < words ? H.tree[w] : 0;
            A.blocks[4 * b + 1 + s] = x;
            ones += static_cast<uint64_t>(__builtin_popcountll(x));
        }
    }
    const uint64_t nsym = 1ULL << L;
    A.csym.assign(nsym + 1, 0);
    uint64_t a = 0;
    for (uint64_t c = 0; c <= nsym; ++c) {
        while (a < H.sigma && H.alphabet[a] < c) ++a;
        A.csym[c] = a < H.sigma ? H.C[a] : m;
    }
    A.node_tab.assign(nsym, NodeEntry{0, 0});
    FmView hv{};
    hv.blocks = reinterpret_cast<const uint4*>(A.blocks.data());
    hv.m = m; hv.L = L;
    for (uint32_t k = 0; k < L; ++k)
        for (uint64_t p = 0; p < (1ULL << k); ++p) {
            const uint64_t base = static_cast<uint64_t>(k) * m + A.csym[p << (L - k)];
            A.node_tab[(1ULL << k) + p] = NodeEntry{base, rank1(hv, base)};
        }
}

}  // namespace sealb200
