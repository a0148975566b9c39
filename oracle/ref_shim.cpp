// TEST INFRASTRUCTURE ONLY — never linked into, imported by, or called from the product path.
//
// extern "C" shim around the UNMODIFIED reference class `FMIndex`
// (/root/reference/seal/cpp_modules/fm_index.{hpp,cpp}) so that tests and bench.py's reference arm
// can drive the reference's own sdsl-lite path through ctypes.  SWIG (the reference's binding
// generator, seal/cpp_modules/fm_index.i) is not installed in this image, so this file plays the
// role of the SWIG wrapper; it adds no logic.  Built by oracle/Makefile into oracle/_ref/.
#include "fm_index.hpp"   // from /root/reference/seal/cpp_modules (via -I)
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

extern "C" {

void* ref_new() { return new FMIndex(); }
void ref_free(void* h) { delete static_cast<FMIndex*>(h); }

void ref_initialize(void* h, const uint64_t* data, uint64_t n) {
    std::vector<char_type> v(data, data + n);
    static_cast<FMIndex*>(h)->initialize(v);
}
void ref_initialize_from_file(void* h, const char* path, int width) {
    static_cast<FMIndex*>(h)->initialize_from_file(std::string(path), width);
}
void* ref_load(const char* path) { return new FMIndex(load_FMIndex(std::string(path))); }
void ref_save(void* h, const char* path) { static_cast<FMIndex*>(h)->save(std::string(path)); }
uint64_t ref_size(void* h) { return static_cast<FMIndex*>(h)->size(); }
uint64_t ref_sigma(void* h) { return static_cast<FMIndex*>(h)->index.wavelet_tree.sigma; }
uint32_t ref_max_level(void* h) { return static_cast<FMIndex*>(h)->index.wavelet_tree.max_level; }

void ref_backward_search_step(void* h, uint64_t sym, uint64_t lo, uint64_t hi, uint64_t* out2) {
    const std::vector<size_type> r = static_cast<FMIndex*>(h)->backward_search_step(sym, lo, hi);
    out2[0] = r[0]; out2[1] = r[1];
}
// n triples in, n pairs out; one reference call per triple (what seal/index.py:108 does per token)
void ref_backward_search_step_batch(void* h, uint64_t n, const uint64_t* sym, const uint64_t* lo,
                                    const uint64_t* hi, uint64_t* out_lo, uint64_t* out_hi) {
    FMIndex* fm = static_cast<FMIndex*>(h);
    for (uint64_t i = 0; i < n; i++) {
        const std::vector<size_type> r = fm->backward_search_step(sym[i], lo[i], hi[i]);
        out_lo[i] = r[0]; out_hi[i] = r[1];
    }
}
void ref_backward_search_multi(void* h, const uint64_t* q, uint64_t n, uint64_t* out2) {
    std::vector<char_type> v(q, q + n);
    const std::vector<size_type> r = static_cast<FMIndex*>(h)->backward_search_multi(v);
    out2[0] = r[0]; out2[1] = r[1];
}
// returns the length of the reference's return vector; copies min(len, cap) entries
uint64_t ref_distinct(void* h, uint64_t lo, uint64_t hi, uint64_t* out, uint64_t cap) {
    const std::vector<char_type> r = static_cast<FMIndex*>(h)->distinct(lo, hi);
    uint64_t m = r.size() < cap ? r.size() : cap;
    if (m) std::memcpy(out, r.data(), m * sizeof(uint64_t));
    return r.size();
}
uint64_t ref_distinct_count(void* h, uint64_t lo, uint64_t hi, uint64_t* out, uint64_t cap) {
    const std::vector<char_type> r = static_cast<FMIndex*>(h)->distinct_count(lo, hi);
    uint64_t m = r.size() < cap ? r.size() : cap;
    if (m) std::memcpy(out, r.data(), m * sizeof(uint64_t));
    return r.size();
}
// The reference's threaded path (fm_index.cpp:111-131, one std::async per range).  Flattened:
// offsets[i]..offsets[i+1] index into out (interleaved sym,count).  Returns total length; if
// out == nullptr only runs the reference call (timing) and fills offsets when non-null.
uint64_t ref_distinct_count_multi(void* h, uint64_t n, const uint64_t* lows, const uint64_t* highs,
                                  uint64_t* offsets, uint64_t* out, uint64_t cap) {
    std::vector<size_type> lo(lows, lows + n), hi(highs, highs + n);
    const std::vector<std::vector<char_type>> r = static_cast<FMIndex*>(h)->distinct_count_multi(lo, hi);
    uint64_t tot = 0;
    for (uint64_t i = 0; i < n; i++) {
        if (offsets) offsets[i] = tot;
        if (out) for (size_t j = 0; j < r[i].size() && tot + j < cap; j++) out[tot + j] = r[i][j];
        tot += r[i].size();
    }
    if (offsets) offsets[n] = tot;
    return tot;
}
uint64_t ref_locate(void* h, uint64_t row) { return static_cast<FMIndex*>(h)->locate(row); }
uint64_t ref_extract_text(void* h, uint64_t begin, uint64_t end, uint64_t* out, uint64_t cap) {
    const std::vector<char_type> r = static_cast<FMIndex*>(h)->extract_text(begin, end);
    uint64_t m = r.size() < cap ? r.size() : cap;
    if (m) std::memcpy(out, r.data(), m * sizeof(uint64_t));
    return r.size();
}

}  // extern "C"
