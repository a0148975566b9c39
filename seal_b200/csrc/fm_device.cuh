// Device-side FM-index: layout + per-thread primitives.
//
// Layout in HBM (all read-only after upload):
//   blocks     32-byte records {u64 ones_before, u64 w0, u64 w1, u64 w2}: the level-concatenated
//              wavelet-tree bit-vector (same logical bit order as sdsl's m_tree, so positions one
//              past a node behave exactly as in the reference — SURVEY.md §H1) cut into 192-bit
//              payloads with the absolute rank in front.  One rank query = ONE 32-byte sector
//              (sdsl: 16 B of counts + 8 B of bits in two different arrays, rank_support_v.hpp:114-124).
//   csym       u64[2^L + 1]: number of BWT symbols < c (C array over the dense symbol space).
//   node_tab   {u64 base, u64 ones}[2^L], heap order (index (1<<k)+p): absolute bit position of the
//              first bit of node (level k, prefix p) = k*m + csym[p << (L-k)], and rank1 there.
//              Removes the two node-boundary ranks sdsl performs per level (wt_int.hpp:365-367) —
//              they are static per node — and is one 16-byte load; the two children of a node sit
//              in one 32-byte sector.
//   sa_samples u64[ceil(m/32)], SA[32 i];   isa_samples u64[(m-1)/64+1], ISA[64 i]
//
// Everything here is per-thread code marked SEAL_HD so tests/hostcheck can compile the very same
// functions with g++ and compare them with the oracle where no GPU exists (test-only; the library
// never runs them on the host).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define SEAL_HD __host__ __device__ __forceinline__
#else
#define SEAL_HD inline
struct uint4 { unsigned int x, y, z, w; };
#endif

namespace sealb200 {

struct NodeEntry { uint64_t base, ones; };

struct FmView {
    const uint4* blocks;
    const uint64_t* csym;
    const NodeEntry* node_tab;
    const uint64_t* sa_samples;
    const uint64_t* isa_samples;
    const uint64_t* beginnings;     // optional (doc offsets), may be null
    uint64_t n_beginnings;
    uint64_t n_isa;
    uint64_t m;                     // BWT length (n+1)
    uint32_t L;                     // wavelet tree height
};

SEAL_HD int popc64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}

SEAL_HD NodeEntry load_node(const FmView& v, uint32_t heap) {
#if defined(__CUDA_ARCH__)
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(v.node_tab + heap));
    NodeEntry e;
    e.base = (static_cast<uint64_t>(q.y) << 32) | q.x;
    e.ones = (static_cast<uint64_t>(q.w) << 32) | q.z;
    return e;
#else
    return v.node_tab[heap];
#endif
}

// rank1(p) over the level-concatenated bit-vector; *bit (optional) receives bit p itself.
SEAL_HD uint64_t rank1(const FmView& v, uint64_t p, int* bit = nullptr) {
    const uint64_t w = p >> 6;
    const uint64_t b = w / 3;
    const uint32_t s = static_cast<uint32_t>(w - 3 * b);
    const uint32_t o = static_cast<uint32_t>(p & 63);
#if defined(__CUDA_ARCH__)
    const uint4 q0 = __ldg(v.blocks + 2 * b);
    const uint4 q1 = __ldg(v.blocks + 2 * b + 1);
#else
    const uint4 q0 = v.blocks[2 * b];
    const uint4 q1 = v.blocks[2 * b + 1];
#endif
    const uint64_t abs = (static_cast<uint64_t>(q0.y) << 32) | q0.x;
    const uint64_t w0 = (static_cast<uint64_t>(q0.w) << 32) | q0.z;
    const uint64_t w1 = (static_cast<uint64_t>(q1.y) << 32) | q1.x;
    const uint64_t w2 = (static_cast<uint64_t>(q1.w) << 32) | q1.z;
    uint64_t r = abs;
    uint64_t cur = w0;
    if (s >= 1) { r += popc64(w0); cur = w1; }
    if (s >= 2) { r += popc64(w1); cur = w2; }
    r += popc64(cur & ((1ULL << o) - 1));
    if (bit) *bit = static_cast<int>((cur >> o) & 1);
    return r;
}

// Occurrences of symbol c (< 2^L) in BWT[0,i) and BWT[0,j): wt_int::rank (sdsl/wt_int.hpp:356-380)
// for two positions at once so the two dependent load chains overlap.
SEAL_HD void wt_rank2(const FmView& v, uint64_t i, uint64_t j, uint32_t c, uint64_t& ri, uint64_t& rj) {
    const uint32_t L = v.L;
    NodeEntry e = load_node(v, 1);                          // root: base 0, ones 0
    for (uint32_t k = 0; k < L && (i | j); ++k) {
        // the next node on c's path is known without looking at the data: fetch its entry now so
        // that each level costs one dependent memory round trip (the rank sector), not two
        NodeEntry nx = e;
        if (k + 1 < L) nx = load_node(v, (2u << k) + (c >> (L - 1 - k)));
        const uint64_t a = rank1(v, e.base + i) - e.ones;
        const uint64_t b = rank1(v, e.base + j) - e.ones;
        if ((c >> (L - 1 - k)) & 1) { i = a; j = b; }
        else { i -= a; j -= b; }
        e = nx;
    }
    ri = i; rj = j;
}

SEAL_HD bool sym_present(const FmView& v, uint64_t c) {
    if (c >= (1ULL << v.L)) return false;
    return c == 0 || v.csym[c + 1] > v.csym[c];
}

// sdsl::backward_search (sdsl/suffix_array_algorithm.hpp:163-191) as wrapped by
// FMIndex::backward_search_step (fm_index.cpp:67-76): inclusive r in, inclusive r out.
SEAL_HD void lf_step(const FmView& v, uint64_t c, uint64_t l, uint64_t r, uint64_t& l_res, uint64_t& r_res) {
    if (!sym_present(v, c)) { l_res = 1; r_res = 0; return; }     // unknown symbol (:176-178)
    const uint64_t cb = v.csym[c];
    if (l == 0 && r + 1 == v.m) {                                  // whole-range shortcut (:181-183)
        l_res = cb; r_res = v.csym[c + 1] - 1; return;
    }
    uint64_t a, b;
    wt_rank2(v, l, r + 1, static_cast<uint32_t>(c), a, b);
    l_res = cb + a;
    r_res = cb + b - 1;
}

// wt_int::inverse_select (sdsl/wt_int.hpp:391-414): symbol at BWT[i] and its rank.
SEAL_HD uint64_t inverse_select(const FmView& v, uint64_t i, uint32_t& c_out) {
    const uint32_t L = v.L;
    uint32_t prefix = 0;
    for (uint32_t k = 0; k < L; ++k) {
        const NodeEntry e = load_node(v, (1u << k) + prefix);
        int bit;
        const uint64_t a = rank1(v, e.base + i, &bit) - e.ones;
        i = bit ? a : i - a;
        prefix = (prefix << 1) | static_cast<uint32_t>(bit);
    }
    c_out = prefix;
    return i;
}

// lf[i] (sdsl/suffix_array_helper.hpp:337-348)
SEAL_HD uint64_t lf_row(const FmView& v, uint64_t i) {
    uint32_t c;
    const uint64_t j = inverse_select(v, i, c);
    return v.csym[c] + j;
}

// csa_wt::operator[] (sdsl/csa_wt.hpp:335-348) behind FMIndex::locate (fm_index.cpp:163-167)
SEAL_HD uint64_t locate_row(const FmView& v, uint64_t row) {
    if (row >= v.m) return ~0ULL;
    uint64_t off = 0;
    while (row & 31) { row = lf_row(v, row); ++off; }
    const uint64_t res = v.sa_samples[row >> 5] + off;
    return res < v.m ? res : res - v.m;
}

// bisect_right(beginnings, pos) - 1  (seal/index.py:77-82)
SEAL_HD uint64_t doc_of_pos(const FmView& v, uint64_t pos) {
    uint64_t lo = 0, hi = v.n_beginnings;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (pos < v.beginnings[mid]) hi = mid; else lo = mid + 1;
    }
    return lo - 1;
}

// isa_of_csa_wt::operator[] (sdsl/suffix_array_helper.hpp:500-514)
SEAL_HD uint64_t isa_at(const FmView& v, uint64_t i) {
    const uint64_t ci = (i / 64 + 1) % v.n_isa;
    uint64_t res = v.isa_samples[ci];
    const uint64_t pos = ci * 64;
    uint64_t steps = pos < i ? pos + v.m - i : pos - i;
    while (steps--) res = lf_row(v, res);
    return res;
}

// FMIndex::extract_text (fm_index.cpp:169-184); writes end-begin symbols.
SEAL_HD void extract_text(const FmView& v, uint64_t begin, uint64_t end, uint64_t* out) {
    if (end <= begin) return;
    uint64_t start = isa_at(v, end);
    uint32_t c;
    (void)inverse_select(v, start, c);
    out[0] = c;
    for (uint64_t t = 1; t < end - begin; ++t) {
        uint64_t l, r;
        lf_step(v, c, start, start + 1, l, r);
        start = l;
        (void)inverse_select(v, start, c);
        out[t] = c;
    }
}

// Depth-first expansion of the distinct symbols of BWT[i,j) below node (level, prefix):
// wt_int::_interval_symbols (sdsl/wt_int.hpp:108-147) with the node-boundary ranks taken from the
// node tables.  sink(symbol, rank_i, rank_j) is called in ascending symbol order.
template <typename Sink>
SEAL_HD void expand_dfs(const FmView& v, uint32_t level, uint32_t prefix, uint64_t i, uint64_t j, Sink& sink) {
    struct Frame { uint64_t i, j; NodeEntry e; uint32_t prefix, level; };
    Frame stk[36];
    int sp = 0;
    const uint32_t L = v.L;
    {
        NodeEntry e0{0, 0};
        if (level < L) e0 = load_node(v, (1u << level) + prefix);
        stk[sp++] = Frame{i, j, e0, prefix, level};
    }
    while (sp) {
        const Frame f = stk[--sp];
        if (f.level == L) { sink(f.prefix, f.i, f.j); continue; }
        // children's table entries share one 32-byte sector; fetched together with the rank sectors
        NodeEntry c0{0, 0}, c1{0, 0};
        if (f.level + 1 < L) {
            const uint32_t h = (2u << f.level) + 2u * f.prefix;
            c0 = load_node(v, h); c1 = load_node(v, h + 1);
        }
        uint64_t a, b;
        if (f.j == f.i + 1) {                      // single position: one sector, take the bit
            int bit;
            a = rank1(v, f.e.base + f.i, &bit) - f.e.ones;
            b = a + static_cast<uint64_t>(bit);
        } else {
            a = rank1(v, f.e.base + f.i) - f.e.ones;
            b = rank1(v, f.e.base + f.j) - f.e.ones;
        }
        const uint64_t ones = b - a;
        const uint64_t zeros = (f.j - f.i) - ones;
        if (ones) stk[sp++] = Frame{a, b, c1, (f.prefix << 1) | 1u, f.level + 1};
        if (zeros) stk[sp++] = Frame{f.i - a, f.j - b, c0, f.prefix << 1, f.level + 1};
    }
}

}  // namespace sealb200
