// TEST-ONLY harness (never part of libsealb200.so): compiles the per-thread device primitives of
// seal_b200/csrc/fm_device.cuh with g++ so their arithmetic can be compared with the oracle in a
// container that has no GPU.  The shipped library runs these functions on the GPU only.
//
// C interface used by tests/test_hostcheck.py through ctypes.
#include "../../seal_b200/csrc/fm_device.cuh"
#include "../../seal_b200/csrc/fm_host.hpp"
#include "../../seal_b200/csrc/fm_layout.hpp"

#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

using namespace sealb200;

struct HC {
    HostIndex H;
    DeviceArrays A;
    FmView v{};
    std::vector<uint64_t> beginnings;
};

static std::string g_err;

static void bind(HC* h) {
    make_device_arrays(h->H, h->A);
    h->v.blocks = reinterpret_cast<const uint4*>(h->A.blocks.data());
    h->v.csym = h->A.csym.data();
    h->v.node_tab = h->A.node_tab.data();
    h->v.sa_samples = h->H.sa_samples.data();
    h->v.isa_samples = h->H.isa_samples.data();
    h->v.n_isa = h->H.isa_samples.size();
    h->v.m = h->H.size;
    h->v.L = h->H.max_level;
}

extern "C" {
const char* hc_error() { return g_err.c_str(); }
HC* hc_build(const uint64_t* sym, uint64_t n) {
    try { HC* h = new HC(); build_index(sym, n, h->H); bind(h); return h; }
    catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
HC* hc_load(const char* path) {
    try { HC* h = new HC(); load_index(path, h->H); bind(h); return h; }
    catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
void hc_free(HC* h) { delete h; }
uint64_t hc_size(HC* h) { return h->H.size; }
void hc_lf_step(HC* h, uint64_t n, const uint64_t* sym, const uint64_t* lo, const uint64_t* hi,
                uint64_t* ol, uint64_t* oh) {
    for (uint64_t i = 0; i < n; ++i) lf_step(h->v, sym[i], lo[i], hi[i], ol[i], oh[i]);
}
struct VecSink {
    std::vector<uint64_t>* out;
    void operator()(uint32_t s, uint64_t ri, uint64_t rj) const { out->push_back(s); out->push_back(rj - ri); }
};
// per-thread DFS from the root == what one lane does in phase 2 of the warp expansion
uint64_t hc_distinct_count(HC* h, uint64_t lo, uint64_t hi, uint64_t* out, uint64_t cap) {
    std::vector<uint64_t> r;
    if (lo < hi) { VecSink s{&r}; expand_dfs(h->v, 0, 0, lo, hi, s); }
    uint64_t m = r.size() < cap ? r.size() : cap;
    if (m) std::memcpy(out, r.data(), m * 8);
    return r.size();
}
uint64_t hc_locate(HC* h, uint64_t row) { return locate_row(h->v, row); }
void hc_extract(HC* h, uint64_t b, uint64_t e, uint64_t* out) { extract_text(h->v, b, e, out); }
}
