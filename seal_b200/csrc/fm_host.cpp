// Host-side FM-index construction and (de)serialisation.  See fm_host.hpp.
#include "fm_host.hpp"
#include "sais.hpp"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>

namespace sealb200 {

namespace {

inline uint32_t hi_bit(uint64_t x) { uint32_t r = 0; while (x >>= 1) ++r; return r; }

struct FileCloser { void operator()(FILE* f) const { if (f) fclose(f); } };
using FilePtr = std::unique_ptr<FILE, FileCloser>;

struct Reader {
    FilePtr f;
    std::string path;
    explicit Reader(const std::string& p) : f(fopen(p.c_str(), "rb")), path(p) {
        if (!f) throw std::runtime_error("cannot open " + p);
    }
    void read(void* dst, size_t bytes) {
        if (bytes && fread(dst, 1, bytes, f.get()) != bytes)
            throw std::runtime_error("truncated index file " + path);
    }
    void skip(uint64_t bytes) {
        if (fseeko(f.get(), static_cast<off_t>(bytes), SEEK_CUR) != 0)
            throw std::runtime_error("seek failed in " + path);
    }
    uint64_t u64() { uint64_t v; read(&v, 8); return v; }
    uint32_t u32() { uint32_t v; read(&v, 4); return v; }
    uint8_t u8() { uint8_t v; read(&v, 1); return v; }
    bool at_eof() { int c = fgetc(f.get()); if (c == EOF) return true; ungetc(c, f.get()); return false; }
};

// sdsl int_vector<w> stream: u64 bit_size, [u8 width iff w == 0], ceil(bit_size/64) words
// (sdsl/int_vector.hpp:593-609,1546-1595)
struct PackedVec {
    uint64_t bit_size = 0;
    uint8_t width = 0;
    std::vector<uint64_t> words;
    uint64_t count() const { return width ? bit_size / width : 0; }
    uint64_t get(uint64_t i) const {
        if (!width) return 0;
        uint64_t p = i * width, w = p >> 6, o = p & 63;
        uint64_t v = words[w] >> o;
        if (o + width > 64) v |= words[w + 1] << (64 - o);
        return width == 64 ? v : (v & ((1ULL << width) - 1));
    }
};
void read_iv(Reader& r, int fixed_width, PackedVec& v, bool keep = true) {
    v.bit_size = r.u64();
    v.width = fixed_width ? static_cast<uint8_t>(fixed_width) : r.u8();
    uint64_t nw = (v.bit_size + 63) >> 6;
    if (keep) { v.words.resize(nw + 1); v.words[nw] = 0; r.read(v.words.data(), nw * 8); }
    else r.skip(nw * 8);
}
// sdsl select_support_mcl stream (sdsl/select_support_mcl.hpp:425-493): skipped, never used by SEAL.
void skip_mcl(Reader& r) {
    uint64_t arg_cnt = r.u64();
    if (!arg_cnt) return;
    PackedVec tmp;
    read_iv(r, 0, tmp, false);                 // superblock
    read_iv(r, 1, tmp, false);                 // mini_or_long
    uint64_t sb = (arg_cnt + 4095) >> 12;
    for (uint64_t i = 0; i < sb; ++i) read_iv(r, 0, tmp, false);
}

const char kMagic[8] = {'S', 'E', 'A', 'L', 'B', '2', 'F', 'M'};

void load_native(Reader& r, HostIndex& o) {
    uint32_t version = r.u32();
    if (version != 1) throw std::runtime_error("unsupported native index version");
    o.max_level = r.u32();
    o.size = r.u64();
    o.sigma = r.u64();
    uint64_t tw = r.u64(), nsa = r.u64(), nisa = r.u64();
    if (o.max_level == 0 || o.max_level > 32 || tw != ((o.size * o.max_level + 63) >> 6))
        throw std::runtime_error("corrupt native index header");
    o.tree.resize(tw); r.read(o.tree.data(), tw * 8);
    o.alphabet.resize(o.sigma); r.read(o.alphabet.data(), o.sigma * 8);
    o.C.resize(o.sigma + 1); r.read(o.C.data(), (o.sigma + 1) * 8);
    o.sa_samples.resize(nsa); r.read(o.sa_samples.data(), nsa * 8);
    o.isa_samples.resize(nisa); r.read(o.isa_samples.data(), nisa * 8);
}

// sdsl store_to_file(csa_wt_int<>) layout: SURVEY.md Appendix A; sdsl/csa_wt.hpp:374-393,
// sdsl/wt_int.hpp:693-716, sdsl/csa_alphabet_strategy.hpp:582-605, sdsl/sd_vector.hpp:404-427.
void load_sdsl(Reader& r, HostIndex& o) {
    o.size = r.u64();
    uint64_t wt_sigma = r.u64();
    PackedVec tree; read_iv(r, 1, tree);
    PackedVec bb; read_iv(r, 64, bb, false);    // rank_support_v blocks: rebuilt in our own layout
    skip_mcl(r); skip_mcl(r);                   // select1 / select0 supports of the tree
    o.max_level = r.u32();
    if (o.size == 0 || o.max_level == 0 || o.max_level > 32 || tree.bit_size != o.size * o.max_level)
        throw std::runtime_error("not an sdsl csa_wt_int<> index (wavelet-tree header mismatch): " + r.path);
    tree.words.pop_back();
    o.tree.swap(tree.words);

    PackedVec sa; read_iv(r, 0, sa);
    o.sa_samples.resize(sa.count());
    for (uint64_t i = 0; i < o.sa_samples.size(); ++i) o.sa_samples[i] = sa.get(i);
    PackedVec isa; read_iv(r, 0, isa);
    o.isa_samples.resize(isa.count());
    for (uint64_t i = 0; i < o.isa_samples.size(); ++i) o.isa_samples[i] = isa.get(i);

    // int_alphabet: sd_vector m_char | (empty rank/select) | m_C | m_sigma
    uint64_t char_size = r.u64();
    uint8_t wl = r.u8();
    PackedVec low; read_iv(r, 0, low);
    PackedVec high; read_iv(r, 1, high);
    skip_mcl(r); skip_mcl(r);
    PackedVec Cv; read_iv(r, 0, Cv);
    o.sigma = r.u64();
    if (Cv.count() != o.sigma + 1 || o.sigma != wt_sigma)
        throw std::runtime_error("sdsl index: alphabet section inconsistent: " + r.path);
    o.C.resize(o.sigma + 1);
    for (uint64_t i = 0; i <= o.sigma; ++i) o.C[i] = Cv.get(i);
    o.alphabet.clear(); o.alphabet.reserve(o.sigma);
    if (char_size == 0) {
        // contiguous alphabet: char2comp is the identity (csa_alphabet_strategy.hpp:426-430)
        for (uint64_t c = 0; c < o.sigma; ++c) o.alphabet.push_back(c);
    } else {
        // i-th 1 in `high` preceded by z zeros encodes (z << wl) | low[i]   (sd_vector.hpp:547-551)
        uint64_t ones = (wl == 0) ? 0 : low.count();
        uint64_t z = 0, i = 0;
        for (uint64_t p = 0; p < high.bit_size; ++p) {
            if ((high.words[p >> 6] >> (p & 63)) & 1) {
                uint64_t lowv = wl ? low.get(i) : 0;
                o.alphabet.push_back((z << wl) | lowv);
                ++i;
            } else ++z;
        }
        (void)ones;
        if (o.alphabet.size() != o.sigma)
            throw std::runtime_error("sdsl index: cannot decode alphabet bitmap: " + r.path);
    }
    if (!r.at_eof()) throw std::runtime_error("sdsl index: trailing bytes: " + r.path);
}

template <typename Idx>
void suffix_array(const std::vector<uint32_t>& s, uint64_t sigma, std::vector<Idx>& sa) {
    sa.resize(s.size());
    SaIs<uint32_t, Idx>::run(s.data(), sa.data(), static_cast<Idx>(s.size()), static_cast<Idx>(sigma));
}

template <typename Idx>
void finish_build(const std::vector<uint32_t>& s, HostIndex& o) {
    const uint64_t m = o.size;
    std::vector<Idx> sa;
    suffix_array<Idx>(s, o.sigma, sa);

    // samples
    o.sa_samples.resize((m + 31) / 32);
    for (uint64_t i = 0; i < m; i += 32) o.sa_samples[i / 32] = static_cast<uint64_t>(sa[i]);
    o.isa_samples.assign((m - 1) / 64 + 1, 0);
    for (uint64_t i = 0; i < m; ++i) {
        uint64_t p = static_cast<uint64_t>(sa[i]);
        if ((p & 63) == 0) o.isa_samples[p >> 6] = i;
    }
    // BWT in real symbols (all < 2^32, checked by the caller)
    std::vector<uint32_t> cur(m);
    for (uint64_t i = 0; i < m; ++i) {
        uint64_t p = static_cast<uint64_t>(sa[i]);
        uint32_t comp = p ? s[p - 1] : s[m - 1];
        cur[i] = static_cast<uint32_t>(o.alphabet[comp]);
    }
    std::vector<Idx>().swap(sa);

    // level-wise wavelet tree bits; same bit order as sdsl/wt_int.hpp:202-242
    const uint32_t L = o.max_level;
    o.tree.assign((m * L + 63) >> 6, 0);
    std::vector<uint32_t> ones(m);
    uint64_t pos = 0;
    for (uint32_t k = 0; k < L; ++k) {
        const uint32_t bit_shift = L - k - 1;
        const uint32_t node_shift = L - k;        // node id = x >> node_shift (k leading bits)
        uint64_t start = 0;
        while (start < m) {
            const uint64_t node = (node_shift >= 32) ? 0 : (cur[start] >> node_shift);
            uint64_t i = start, c0 = 0, c1 = 0;
            while (i < m && ((node_shift >= 32) ? 0 : (cur[i] >> node_shift)) == node) {
                uint32_t x = cur[i];
                if ((x >> bit_shift) & 1) { o.tree[pos >> 6] |= 1ULL << (pos & 63); ones[c1++] = x; }
                else cur[start + c0++] = x;
                ++pos; ++i;
            }
            std::memcpy(cur.data() + start + c0, ones.data(), c1 * sizeof(uint32_t));
            start = i;
        }
    }
}

void build_from_symbols(const uint64_t* sym, uint64_t n, HostIndex& o) {
    o = HostIndex();
    const uint64_t m = n + 1;
    uint64_t maxsym = 0;
    for (uint64_t i = 0; i < n; ++i) {
        if (sym[i] == 0) throw std::runtime_error("symbol 0 is reserved for the sentinel");
        maxsym = std::max(maxsym, sym[i]);
    }
    if (maxsym >= (1ULL << 32)) throw std::runtime_error("symbols must be < 2^32");
    o.size = m;
    o.max_level = hi_bit(std::max<uint64_t>(maxsym, 1)) + 1;      // sdsl/wt_int.hpp:182-193
    // alphabet + C (sdsl/csa_alphabet_strategy.hpp:494-534)
    std::vector<uint64_t> cnt(maxsym + 1, 0);
    cnt[0] = 1;
    for (uint64_t i = 0; i < n; ++i) cnt[sym[i]]++;
    std::vector<uint32_t> comp(maxsym + 1, 0);
    uint64_t sum = 0;
    for (uint64_t c = 0; c <= maxsym; ++c) {
        if (!cnt[c]) continue;
        comp[c] = static_cast<uint32_t>(o.alphabet.size());
        o.alphabet.push_back(c);
        o.C.push_back(sum);
        sum += cnt[c];
    }
    o.C.push_back(sum);
    o.sigma = o.alphabet.size();
    std::vector<uint32_t> s(m);
    for (uint64_t i = 0; i < n; ++i) s[i] = comp[sym[i]];
    s[n] = 0;
    if (m < (1ULL << 31) - 8) finish_build<int32_t>(s, o);
    else finish_build<int64_t>(s, o);
}

}  // namespace

void build_index(const uint64_t* symbols, uint64_t n, HostIndex& out) {
    build_from_symbols(symbols, n, out);
}

void build_index_from_file(const std::string& path, int width_bytes, HostIndex& out) {
    if (width_bytes != 1 && width_bytes != 2 && width_bytes != 4 && width_bytes != 8)
        throw std::runtime_error("width must be 1, 2, 4 or 8 bytes");
    Reader r(path);
    if (fseeko(r.f.get(), 0, SEEK_END) != 0) throw std::runtime_error("seek failed: " + path);
    uint64_t bytes = static_cast<uint64_t>(ftello(r.f.get()));
    fseeko(r.f.get(), 0, SEEK_SET);
    uint64_t n = bytes / width_bytes;
    std::vector<uint64_t> sym(n);
    std::vector<uint8_t> raw(std::min<uint64_t>(bytes, 1 << 24));
    uint64_t done = 0;
    while (done < n) {
        uint64_t chunk = std::min<uint64_t>(n - done, raw.size() / width_bytes);
        r.read(raw.data(), chunk * width_bytes);
        for (uint64_t i = 0; i < chunk; ++i) {
            uint64_t v = 0;
            std::memcpy(&v, raw.data() + i * width_bytes, width_bytes);   // little-endian
            sym[done + i] = v;
        }
        done += chunk;
    }
    build_from_symbols(sym.data(), n, out);
}

void load_index(const std::string& path, HostIndex& out) {
    out = HostIndex();
    Reader r(path);
    char head[8];
    r.read(head, 8);
    if (std::memcmp(head, kMagic, 8) == 0) { load_native(r, out); return; }
    fseeko(r.f.get(), 0, SEEK_SET);
    load_sdsl(r, out);
}

void save_index_native(const HostIndex& o, const std::string& path) {
    FilePtr f(fopen(path.c_str(), "wb"));
    if (!f) throw std::runtime_error("cannot open for writing: " + path);
    auto w = [&](const void* p, size_t b) {
        if (b && fwrite(p, 1, b, f.get()) != b) throw std::runtime_error("short write: " + path);
    };
    uint32_t version = 1;
    uint64_t tw = o.tree.size(), nsa = o.sa_samples.size(), nisa = o.isa_samples.size();
    w(kMagic, 8); w(&version, 4); w(&o.max_level, 4); w(&o.size, 8); w(&o.sigma, 8);
    w(&tw, 8); w(&nsa, 8); w(&nisa, 8);
    w(o.tree.data(), tw * 8);
    w(o.alphabet.data(), o.sigma * 8);
    w(o.C.data(), (o.sigma + 1) * 8);
    w(o.sa_samples.data(), nsa * 8);
    w(o.isa_samples.data(), nisa * 8);
}

// ------------------------------------------------------------------------------------------------
// sdsl-compatible writer: the byte stream sdsl::store_to_file(csa_wt_int<>) produces for this index
// (SURVEY.md Appendix A), so that an index built here -- on the GPU in tens of milliseconds -- loads in
// the unmodified reference (load_FMIndex, fm_index.cpp:191-199).  Everything the reader above skips has to
// be constructed: the rank_support_v blocks (sdsl/rank_support_v.hpp:67-106), the select_support_mcl
// tables of the tree and of the alphabet's sd_vector (sdsl/select_support_mcl.hpp:209-345, both of its
// construction paths -- they differ in how the last, partial superblock is stored) and the sd_vector
// itself (sdsl/sd_vector.hpp:192-228).  tests/test_host_logic.py compares the files byte for byte with
// the reference's own FMIndex::save where oracle/_ref is available.
// ------------------------------------------------------------------------------------------------
namespace {

struct Writer {
    FilePtr f; std::string path;
    explicit Writer(const std::string& p) : f(fopen(p.c_str(), "wb")), path(p) {
        if (!f) throw std::runtime_error("cannot open for writing: " + p);
    }
    void bytes(const void* d, size_t n) { if (n && fwrite(d, 1, n, f.get()) != n) throw std::runtime_error("short write: " + path); }
    void u64(uint64_t v) { bytes(&v, 8); }
    void u32(uint32_t v) { bytes(&v, 4); }
    void u8(uint8_t v) { bytes(&v, 1); }
};

// sdsl int_vector<0> under construction: n entries of `width` bits
struct Packed {
    uint64_t n = 0; uint8_t width = 64; std::vector<uint64_t> w;
    Packed() = default;
    Packed(uint64_t count, uint8_t wd) : n(count), width(wd), w((count * wd + 63) / 64 + 1, 0) {}
    bool empty() const { return n == 0; }
    void set(uint64_t i, uint64_t v) {
        if (width < 64) v &= (1ULL << width) - 1;                   // int_vector truncates
        const uint64_t p = i * width, k = p >> 6, o = p & 63;
        w[k] |= v << o;
        if (o + width > 64) w[k + 1] |= v >> (64 - o);
    }
    void write(Writer& out, bool with_width_byte) const {
        const uint64_t bits = n * width;
        out.u64(bits);
        if (with_width_byte) out.u8(width);
        out.bytes(w.data(), ((bits + 63) >> 6) * 8);
    }
};

struct BitView {                                                     // a bit_vector: `size` bits in 64-bit words
    const uint64_t* w; uint64_t size;
    uint64_t capacity() const { return ((size + 63) >> 6) << 6; }
    bool bit(uint64_t i) const { return (w[i >> 6] >> (i & 63)) & 1; }
};

inline int nth_set_bit(uint64_t x, uint64_t i) {                     // position of the i-th (1-based) set bit
    for (uint64_t k = 1; k < i; ++k) x &= x - 1;
    return __builtin_ctzll(x);
}

// select_support_mcl<b,1> over `v`, serialised (sdsl/select_support_mcl.hpp:425-463)
void write_mcl(Writer& out, const BitView& v, int b) {
    constexpr uint64_t SB = 4096;
    auto is_arg = [&](uint64_t i) { return v.bit(i) == (b != 0); };
    uint64_t arg_cnt = 0;
    for (uint64_t i = 0; i < (v.size >> 6); ++i) arg_cnt += __builtin_popcountll(b ? v.w[i] : ~v.w[i]);
    for (uint64_t i = (v.size >> 6) << 6; i < v.size; ++i) arg_cnt += is_arg(i);
    out.u64(arg_cnt);
    if (!arg_cnt) return;
    const uint64_t logn = hi_bit(v.capacity()) + 1, logn4 = logn * logn * logn * logn;
    const uint64_t sb = (arg_cnt + SB - 1) / SB;
    Packed super(sb, (uint8_t)logn);
    std::vector<Packed> mini(sb), lng(sb);
    bool any_long = false;
    std::vector<uint64_t> pos(SB);
    if (v.size < 100000) {                                           // init_slow (:209-258)
        uint64_t cnt = 0, sbi = 0;
        for (uint64_t i = 0; i < v.size; ++i) {
            if (!is_arg(i)) continue;
            pos[cnt % SB] = i;
            ++cnt;
            if (cnt % SB == 0 || cnt == arg_cnt) {
                const uint64_t last = (cnt - 1) % SB;
                super.set(sbi, pos[0]);
                const uint64_t diff = pos[last] - pos[0];
                if (diff > logn4) {
                    any_long = true;
                    lng[sbi] = Packed(SB, (uint8_t)(hi_bit(pos[last]) + 1));
                    for (uint64_t j = 0; j <= last; ++j) lng[sbi].set(j, pos[j]);
                } else {
                    mini[sbi] = Packed(64, (uint8_t)(hi_bit(diff) + 1));
                    for (uint64_t j = 0; j <= last; j += 64) mini[sbi].set(j / 64, pos[j] - pos[0]);
                }
                ++sbi;
            }
        }
    } else {
        // Word-wise construction (the path sdsl takes from 100 000 bits on, :261-345).  Only the position of every
        // 64th argument of the current superblock is recorded while scanning (`marks`); a superblock is closed as
        // soon as its 4033rd argument (mark 63) has been seen, by walking on to its last argument.  What is left
        // after the last closed superblock is always stored as a long block of width hi(size-1)+1.
        uint64_t mark = 0;                 // marks recorded in the open superblock
        uint64_t want = 1;                 // rank (1-based, over the whole vector) of the next argument to mark
        uint64_t seen_before = 0, seen = 0, sbi = 0;
        std::vector<uint64_t>& marks = pos;                          // marks[64 * k] = position of argument 64 k of the block
        const uint64_t words = v.capacity() >> 6;
        for (uint64_t wi = 0; wi < words; ++wi) {
            const uint64_t word = b ? v.w[wi] : ~v.w[wi];
            seen += __builtin_popcountll(word);
            if (seen >= want) {
                marks[mark * 64] = wi * 64 + nth_set_bit(word, want - seen_before);
                ++mark; want += 64;
                if (mark == 64) {
                    const uint64_t first = marks[0];
                    uint64_t last = marks[63 * 64];
                    for (uint64_t i = last + 1, have = 63 * 64; i < v.size && have < SB; ++i)
                        if (is_arg(i)) { last = i; ++have; }
                    super.set(sbi, first);
                    if (last - first > logn4) {
                        any_long = true;
                        lng[sbi] = Packed(SB, (uint8_t)(hi_bit(last) + 1));
                        for (uint64_t i = first, k = 0; k < SB && i <= last; ++i) if (is_arg(i)) lng[sbi].set(k++, i);
                    } else {
                        mini[sbi] = Packed(64, (uint8_t)(hi_bit(last - first) + 1));
                        for (uint64_t k = 0; k < 64; ++k) mini[sbi].set(k, marks[64 * k] - first);
                    }
                    ++sbi;
                    mark = 0;
                }
            }
            seen_before = seen;
        }
        if (mark > 0) {
            any_long = true;
            lng[sbi] = Packed(SB, (uint8_t)(hi_bit(v.size - 1) + 1));
            for (uint64_t i = marks[0], k = 0; i < v.size; ++i) if (is_arg(i)) lng[sbi].set(k++, i);
            ++sbi;
        }
    }
    super.write(out, true);
    Packed mol;                                                      // bit i = superblock i is a mini block (:441-446)
    if (any_long) { mol = Packed(sb, 1); for (uint64_t i = 0; i < sb; ++i) if (!mini[i].empty()) mol.set(i, 1); }
    else { mol.n = 0; mol.width = 1; mol.w.assign(1, 0); }
    mol.write(out, false);
    for (uint64_t i = 0; i < sb; ++i) {
        const bool use_long = any_long && mini[i].empty();
        (use_long ? lng[i] : mini[i]).write(out, true);
    }
}

}  // namespace

void save_index_sdsl(const HostIndex& o, const std::string& path) {
    Writer out(path);
    const uint64_t m = o.size, L = o.max_level, tree_bits = m * L;
    // ---- wt_int (sdsl/wt_int.hpp:693-705) ----
    out.u64(m); out.u64(o.sigma);
    out.u64(tree_bits); out.bytes(o.tree.data(), ((tree_bits + 63) >> 6) * 8);
    {   // rank_support_v<1,1> (:67-106): per 512 bits [absolute count][seven 9-bit in-block counts]
        const uint64_t cap = ((tree_bits + 63) >> 6) << 6, W = cap >> 6;
        std::vector<uint64_t> bb(((cap >> 9) + 1) << 1, 0);
        uint64_t sum = __builtin_popcountll(o.tree[0]), second = 0, j = 0, i = 1;
        for (; i < W; ++i) {
            if (!(i & 7)) { j += 2; bb[j - 1] = second; bb[j] = bb[j - 2] + sum; second = sum = 0; }
            else second |= sum << (63 - 9 * (i & 7));
            sum += __builtin_popcountll(o.tree[i]);
        }
        if (i & 7) { second |= sum << (63 - 9 * (i & 7)); bb[j + 1] = second; }
        else { j += 2; bb[j - 1] = second; bb[j] = bb[j - 2] + sum; bb[j + 1] = 0; }
        out.u64(bb.size() * 64); out.bytes(bb.data(), bb.size() * 8);
    }
    const BitView tree{o.tree.data(), tree_bits};
    write_mcl(out, tree, 1);
    write_mcl(out, tree, 0);
    out.u32((uint32_t)L);
    // ---- SA / ISA samples (sdsl/csa_sampling_strategy.hpp:85-99, :626-641): width hi(n)+1 ----
    const uint8_t wn = (uint8_t)(hi_bit(m) + 1);
    Packed sa(o.sa_samples.size(), wn), isa(o.isa_samples.size(), wn);
    for (uint64_t i = 0; i < o.sa_samples.size(); ++i) sa.set(i, o.sa_samples[i]);
    for (uint64_t i = 0; i < o.isa_samples.size(); ++i) isa.set(i, o.isa_samples[i]);
    sa.write(out, true); isa.write(out, true);
    // ---- int_alphabet (sdsl/csa_alphabet_strategy.hpp:494-534, :582-594) ----
    bool continuous = true;
    for (uint64_t c = 0; c < o.sigma; ++c) continuous = continuous && o.alphabet[c] == c;
    if (continuous) {                                                // default-constructed sd_vector + supports
        out.u64(0); out.u8(0);
        Packed().write(out, true);                                   // m_low: empty int_vector<0>, width 64
        out.u64(0);                                                  // m_high: empty bit_vector
        out.u64(0); out.u64(0);                                      // the two select supports: no arguments
    } else {                                                         // sd_vector over the "symbol present" bitmap (sd_vector.hpp:192-228)
        const uint64_t n = o.alphabet.back() + 1, ones = o.sigma;
        uint8_t logm = (uint8_t)(hi_bit(ones) + 1); const uint8_t logn = (uint8_t)(hi_bit(n) + 1);
        if (logm == logn) --logm;
        const uint8_t wl = logn - logm;
        Packed low(ones, wl);
        const uint64_t high_bits = ones + (1ULL << logm);
        std::vector<uint64_t> high((high_bits + 63) / 64 + 1, 0);
        uint64_t last_high = 0, highpos = 0;
        for (uint64_t k = 0; k < ones; ++k) {
            const uint64_t p = o.alphabet[k], cur_high = p >> wl;
            highpos += cur_high - last_high; last_high = cur_high;
            low.set(k, p);
            high[highpos >> 6] |= 1ULL << (highpos & 63); ++highpos;
        }
        out.u64(n); out.u8(wl);
        low.write(out, true);
        out.u64(high_bits); out.bytes(high.data(), ((high_bits + 63) >> 6) * 8);
        const BitView hv{high.data(), high_bits};
        write_mcl(out, hv, 1);
        write_mcl(out, hv, 0);
    }
    Packed C(o.sigma + 1, wn);
    for (uint64_t i = 0; i <= o.sigma; ++i) C.set(i, o.C[i]);
    C.write(out, true);
    out.u64(o.sigma);
}

}  // namespace sealb200
