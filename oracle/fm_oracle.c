/* TEST INFRASTRUCTURE ONLY — see fm_oracle.h.  Plain-C restatement of the sdsl-lite 2.1.0
 * csa_wt_int<> arithmetic SEAL's FM-index runs on.  "sdsl/" = /root/reference/res/external/
 * sdsl-lite/include/sdsl/.  Written for clarity, not speed; the suffix sorter is a qsort-based
 * prefix doubling that is fine up to a few million symbols. */
#define _GNU_SOURCE
#include "fm_oracle.h"
#include <stdlib.h>
#include <string.h>

struct fmo {
    uint64_t size;        /* n+1 (with sentinel)                         sdsl/csa_wt.hpp size()   */
    uint32_t max_level;   /* L                                           sdsl/wt_int.hpp:189-193  */
    uint64_t sigma;       /* number of distinct symbols incl. sentinel                            */
    uint64_t* sa;         /* explicit suffix array, size entries                                  */
    uint64_t* bwt;        /* explicit BWT, size entries                                           */
    uint64_t* tree;       /* level-concatenated wavelet-tree bits + 1 zero padding word           */
    uint64_t tree_words;  /* ceil(size*L/64)                                                      */
    uint64_t* bb;         /* rank_support_v basic blocks                                          */
    uint64_t bb_words;
    uint64_t* alpha;      /* ascending symbols; index = comp char                                 */
    uint64_t* C;          /* sigma+1 cumulative counts                                            */
    uint64_t* sa_s;  uint64_t n_sa_s;    /* SA[i] for i % 32 == 0                                 */
    uint64_t* isa_s; uint64_t n_isa_s;   /* ISA[j] for j % 64 == 0                                */
};

/* ---- suffix sorting (what sdsl does with qsufsort, sdsl/construct_sa.hpp:162-166; any correct
 * sorter yields the same SA because the sentinel makes all suffixes distinct) ------------------ */
static const uint64_t* g_rk; static uint64_t g_k, g_n;
static int cmp_sfx(const void* pa, const void* pb) {
    uint64_t a = *(const uint64_t*)pa, b = *(const uint64_t*)pb;
    if (g_rk[a] != g_rk[b]) return g_rk[a] < g_rk[b] ? -1 : 1;
    uint64_t ra = a + g_k < g_n ? g_rk[a + g_k] + 1 : 0;
    uint64_t rb = b + g_k < g_n ? g_rk[b + g_k] + 1 : 0;
    if (ra != rb) return ra < rb ? -1 : 1;
    return 0;
}
static int suffix_sort(const uint64_t* t, uint64_t n, uint64_t* sa) {
    uint64_t* rk = (uint64_t*)malloc(n * 8), *tmp = (uint64_t*)malloc(n * 8);
    if (!rk || !tmp) { free(rk); free(tmp); return -1; }
    for (uint64_t i = 0; i < n; i++) { sa[i] = i; rk[i] = t[i]; }
    for (uint64_t k = 0;; k = k ? 2 * k : 1) {
        /* k == 0: sort by first symbol only */
        g_rk = rk; g_k = k ? k : n; g_n = n;
        qsort(sa, n, 8, cmp_sfx);
        tmp[sa[0]] = 0;
        uint64_t r = 0;
        for (uint64_t i = 1; i < n; i++) {
            if (cmp_sfx(&sa[i - 1], &sa[i]) != 0) r++;
            tmp[sa[i]] = r;
        }
        memcpy(rk, tmp, n * 8);
        if (r == n - 1) break;
    }
    free(rk); free(tmp);
    return 0;
}

static inline uint32_t hi_bit(uint64_t x) { /* sdsl bits::hi: index of the highest set bit, hi(0)=0 */
    uint32_t r = 0; while (x >>= 1) r++; return r;
}
static inline uint64_t popcnt64(uint64_t x) { return (uint64_t)__builtin_popcountll(x); }

/* ---- rank_support_v<1,1>: construction sdsl/rank_support_v.hpp:67-106 -------------------------- */
static void build_rank_blocks(fmo_t* o) {
    uint64_t capacity = o->tree_words << 6;               /* int_vector::capacity() in bits */
    o->bb_words = ((capacity >> 9) + 1) << 1;
    o->bb = (uint64_t*)calloc(o->bb_words + 2, 8);
    const uint64_t* data = o->tree;
    uint64_t i, j = 0;
    o->bb[0] = o->bb[1] = 0;
    uint64_t sum = popcnt64(data[0]);
    uint64_t second = 0;
    for (i = 1; i < (capacity >> 6); ++i) {
        if (!(i & 0x7)) {
            j += 2;
            o->bb[j - 1] = second;
            o->bb[j] = o->bb[j - 2] + sum;
            second = sum = 0;
        } else {
            second |= sum << (63 - 9 * (i & 0x7));
        }
        sum += popcnt64(data[i]);
    }
    if (i & 0x7) {
        second |= sum << (63 - 9 * (i & 0x7));
        o->bb[j + 1] = second;
    } else {
        j += 2;
        o->bb[j - 1] = second;
        o->bb[j] = o->bb[j - 2] + sum;
        o->bb[j + 1] = 0;
    }
}

/* rank_support_v::rank, sdsl/rank_support_v.hpp:114-124.  idx may be one past a node / the tree
 * (SEAL's first-step quirk): the padding word after the tree is zero
 * (sdsl/memory_management.hpp:351-366). */
uint64_t fmo_bv_rank(const fmo_t* o, uint64_t idx) {
    const uint64_t* p = o->bb + ((idx >> 8) & 0xFFFFFFFFFFFFFFFEULL);
    uint64_t r = p[0] + ((p[1] >> (63 - 9 * ((idx & 0x1FF) >> 6))) & 0x1FF);
    if (idx & 0x3F) r += popcnt64(o->tree[idx >> 6] & ((1ULL << (idx & 0x3F)) - 1));
    return r;
}
static inline int tree_bit(const fmo_t* o, uint64_t p) { return (int)((o->tree[p >> 6] >> (p & 63)) & 1); }

/* ---- wavelet tree construction, sdsl/wt_int.hpp:169-256 -------------------------------------- */
static int build_wt(fmo_t* o) {
    uint64_t m = o->size;
    uint64_t x = 1;
    for (uint64_t i = 0; i < m; i++) if (o->bwt[i] > x) x = o->bwt[i];
    o->max_level = hi_bit(x) + 1;
    uint32_t L = o->max_level;
    uint64_t bit_size = m * L;
    o->tree_words = (bit_size + 63) >> 6;
    o->tree = (uint64_t*)calloc(o->tree_words + 1, 8);    /* + zero padding word */
    uint64_t* rac = (uint64_t*)malloc(m * 8), *buf1 = (uint64_t*)malloc(m * 8);
    if (!o->tree || !rac || !buf1) { free(rac); free(buf1); return -1; }
    memcpy(rac, o->bwt, m * 8);
    o->sigma = 0;
    uint64_t tree_pos = 0;
    uint64_t mask_old = 1ULL << L;
    for (uint32_t k = 0; k < L; ++k) {
        uint64_t start = 0;
        const uint64_t mask_new = 1ULL << (L - k - 1);
        do {
            uint64_t i = start, cnt0 = 0, cnt1 = 0;
            uint64_t start_value = rac[i] & mask_old, v;
            while (i < m && ((v = rac[i]) & mask_old) == start_value) {
                if (v & mask_new) {
                    o->tree[tree_pos >> 6] |= 1ULL << (tree_pos & 63);
                    buf1[cnt1++] = v;
                } else {
                    rac[start + cnt0++] = v;
                }
                ++tree_pos; ++i;
            }
            if (k + 1 < L) {
                for (uint64_t j = 0; j < cnt1; ++j) rac[start + cnt0 + j] = buf1[j];
            } else {
                o->sigma += (cnt0 > 0) + (cnt1 > 0);
            }
            start += cnt0 + cnt1;
        } while (start < m);
        mask_old += mask_new;
    }
    free(rac); free(buf1);
    build_rank_blocks(o);
    return 0;
}

static int cmp_u64(const void* a, const void* b) {
    uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b; return x < y ? -1 : x > y;
}

fmo_t* fmo_build(const uint64_t* text, uint64_t n) {
    fmo_t* o = (fmo_t*)calloc(1, sizeof(fmo_t));
    if (!o) return NULL;
    uint64_t m = n + 1;
    o->size = m;
    uint64_t* t = (uint64_t*)malloc(m * 8);
    o->sa = (uint64_t*)malloc(m * 8);
    o->bwt = (uint64_t*)malloc(m * 8);
    if (!t || !o->sa || !o->bwt) { free(t); fmo_free(o); return NULL; }
    memcpy(t, text, n * 8); t[n] = 0;                      /* sdsl/construct.hpp:48-52 */
    if (suffix_sort(t, m, o->sa)) { free(t); fmo_free(o); return NULL; }
    for (uint64_t i = 0; i < m; i++) o->bwt[i] = o->sa[i] ? t[o->sa[i] - 1] : t[m - 1];
    /* alphabet + C, sdsl/csa_alphabet_strategy.hpp:494-534 */
    uint64_t* s = (uint64_t*)malloc(m * 8);
    memcpy(s, t, m * 8); qsort(s, m, 8, cmp_u64);
    uint64_t sg = 0;
    for (uint64_t i = 0; i < m; i++) if (i == 0 || s[i] != s[i - 1]) sg++;
    o->alpha = (uint64_t*)malloc(sg * 8);
    o->C = (uint64_t*)malloc((sg + 1) * 8);
    uint64_t idx = 0;
    for (uint64_t i = 0; i < m; i++) if (i == 0 || s[i] != s[i - 1]) { o->alpha[idx] = s[i]; o->C[idx] = i; idx++; }
    o->C[sg] = m;
    free(s); free(t);
    if (build_wt(o)) { fmo_free(o); return NULL; }
    /* SA samples every 32nd row, sdsl/csa_sampling_strategy.hpp:85-99 */
    o->n_sa_s = (m + 31) / 32;
    o->sa_s = (uint64_t*)malloc(o->n_sa_s * 8);
    for (uint64_t i = 0; i < m; i += 32) o->sa_s[i / 32] = o->sa[i];
    /* ISA samples every 64th text position, sdsl/csa_sampling_strategy.hpp:626-641 */
    o->n_isa_s = (m - 1) / 64 + 1;
    o->isa_s = (uint64_t*)calloc(o->n_isa_s, 8);
    for (uint64_t i = 0; i < m; i++) if (o->sa[i] % 64 == 0) o->isa_s[o->sa[i] / 64] = i;
    return o;
}

void fmo_free(fmo_t* o) {
    if (!o) return;
    free(o->sa); free(o->bwt); free(o->tree); free(o->bb); free(o->alpha); free(o->C);
    free(o->sa_s); free(o->isa_s); free(o);
}

uint64_t fmo_size(const fmo_t* o) { return o->size; }
uint64_t fmo_sigma(const fmo_t* o) { return o->sigma; }
uint32_t fmo_max_level(const fmo_t* o) { return o->max_level; }

/* wt_int::rank, sdsl/wt_int.hpp:356-380 */
uint64_t fmo_wt_rank(const fmo_t* o, uint64_t i, uint64_t c) {
    if ((1ULL << o->max_level) <= c) return 0;
    uint64_t offset = 0, mask = 1ULL << (o->max_level - 1), node_size = o->size;
    for (uint32_t k = 0; k < o->max_level && i; ++k) {
        uint64_t ones_before_o = fmo_bv_rank(o, offset);
        uint64_t ones_before_i = fmo_bv_rank(o, offset + i) - ones_before_o;
        uint64_t ones_before_end = fmo_bv_rank(o, offset + node_size) - ones_before_o;
        if (c & mask) {
            offset += node_size - ones_before_end;
            node_size = ones_before_end;
            i = ones_before_i;
        } else {
            node_size = node_size - ones_before_end;
            i = i - ones_before_i;
        }
        offset += o->size;
        mask >>= 1;
    }
    return i;
}

/* wt_int::inverse_select, sdsl/wt_int.hpp:391-414: returns rank(i, wt[i]) and wt[i] */
static uint64_t inverse_select(const fmo_t* o, uint64_t i, uint64_t* c_out) {
    uint64_t c = 0, node_size = o->size, offset = 0;
    for (uint32_t k = 0; k < o->max_level; ++k) {
        uint64_t ones_before_o = fmo_bv_rank(o, offset);
        uint64_t ones_before_i = fmo_bv_rank(o, offset + i) - ones_before_o;
        uint64_t ones_before_end = fmo_bv_rank(o, offset + node_size) - ones_before_o;
        c <<= 1;
        if (tree_bit(o, offset + i)) {
            offset += node_size - ones_before_end;
            node_size = ones_before_end;
            i = ones_before_i;
            c |= 1;
        } else {
            node_size = node_size - ones_before_end;
            i = i - ones_before_i;
        }
        offset += o->size;
    }
    *c_out = c;
    return i;
}

/* int_alphabet::char2comp, sdsl/csa_alphabet_strategy.hpp:420-432 (0 when absent) */
static uint64_t char2comp(const fmo_t* o, uint64_t c) {
    uint64_t lo = 0, hi = o->sigma;
    while (lo < hi) { uint64_t mid = (lo + hi) / 2; if (o->alpha[mid] < c) lo = mid + 1; else hi = mid; }
    return (lo < o->sigma && o->alpha[lo] == c) ? lo : 0;
}

/* sdsl::backward_search, sdsl/suffix_array_algorithm.hpp:163-191, as wrapped by
 * FMIndex::backward_search_step, seal/cpp_modules/fm_index.cpp:67-76 */
void fmo_backward_search_step(const fmo_t* o, uint64_t c, uint64_t l, uint64_t r, uint64_t out[2]) {
    uint64_t cc = char2comp(o, c);
    if (cc == 0 && c > 0) { out[0] = 1; out[1] = 0; return; }
    uint64_t c_begin = o->C[cc];
    if (l == 0 && r + 1 == o->size) {
        out[0] = c_begin; out[1] = o->C[cc + 1] - 1;
    } else {
        out[0] = c_begin + fmo_wt_rank(o, l, c);
        out[1] = c_begin + fmo_wt_rank(o, r + 1, c) - 1;
    }
}

/* FMIndex::backward_search_multi, fm_index.cpp:55-65 (starts from r = size()) */
void fmo_backward_search_multi(const fmo_t* o, const uint64_t* q, uint64_t n, uint64_t out[2]) {
    uint64_t lr[2] = {0, o->size};
    for (uint64_t i = 0; i < n; i++) fmo_backward_search_step(o, q[i], lr[0], lr[1], lr);
    out[0] = lr[0]; out[1] = lr[1] + 1;
}

/* wt_int::_interval_symbols, sdsl/wt_int.hpp:108-147 */
typedef struct { uint64_t* out; uint64_t cap, len, nodes; } isym_ctx;
static void isym_rec(const fmo_t* o, uint64_t i, uint64_t j, uint32_t level, uint64_t path,
                     uint64_t node_size, uint64_t offset, isym_ctx* cx) {
    if (level >= o->max_level) {
        if (cx->len + 1 < cx->cap) { cx->out[cx->len] = path; cx->out[cx->len + 1] = j - i; }
        cx->len += 2;
        return;
    }
    cx->nodes++;
    uint64_t ones_before_o = fmo_bv_rank(o, offset);
    uint64_t ones_before_i = fmo_bv_rank(o, offset + i) - ones_before_o;
    uint64_t ones_before_j = fmo_bv_rank(o, offset + j) - ones_before_o;
    uint64_t ones_before_end = fmo_bv_rank(o, offset + node_size) - ones_before_o;
    if ((j - i) - (ones_before_j - ones_before_i) > 0)
        isym_rec(o, i - ones_before_i, j - ones_before_j, level + 1, path << 1,
                 node_size - ones_before_end, offset + o->size, cx);
    if ((ones_before_j - ones_before_i) > 0)
        isym_rec(o, ones_before_i, ones_before_j, level + 1, (path << 1) | 1,
                 ones_before_end, offset + (node_size - ones_before_end) + o->size, cx);
}

/* FMIndex::distinct_count, fm_index.cpp:91-109 → sdsl::interval_symbols, sdsl/wt_int.hpp:489-509 */
uint64_t fmo_distinct_count(const fmo_t* o, uint64_t lo, uint64_t hi, uint64_t* out, uint64_t cap) {
    if (lo == hi) return 0;
    if (lo + 1 == hi) {                                    /* sdsl/wt_int.hpp:497-504 */
        uint64_t c; (void)inverse_select(o, lo, &c);
        if (cap >= 2) { out[0] = c; out[1] = 1; }
        return 2;
    }
    isym_ctx cx = {out, cap, 0, 0};
    isym_rec(o, lo, hi, 0, 0, o->size, 0, &cx);
    return cx.len;
}

uint64_t fmo_visited_nodes(const fmo_t* o, uint64_t lo, uint64_t hi) {
    if (lo >= hi) return 0;
    if (lo + 1 == hi) return o->max_level;
    isym_ctx cx = {NULL, 0, 0, 0};
    isym_rec(o, lo, hi, 0, 0, o->size, 0, &cx);
    return cx.nodes;
}

/* lf[i], sdsl/suffix_array_helper.hpp:337-348 */
static uint64_t lf(const fmo_t* o, uint64_t i) {
    uint64_t c, j = inverse_select(o, i, &c);
    return o->C[char2comp(o, c)] + j;
}

/* FMIndex::locate, fm_index.cpp:163-167 → csa_wt::operator[], sdsl/csa_wt.hpp:335-348 */
uint64_t fmo_locate(const fmo_t* o, uint64_t row) {
    if (row >= o->size) return (uint64_t)-1;
    uint64_t off = 0, i = row;
    while (i % 32) { i = lf(o, i); ++off; }
    uint64_t result = o->sa_s[i / 32];
    return result + off < o->size ? result + off : result + off - o->size;
}

/* isa_of_csa_wt::operator[], sdsl/suffix_array_helper.hpp:500-514 */
static uint64_t isa_at(const fmo_t* o, uint64_t i) {
    uint64_t ci = (i / 64 + 1) % o->n_isa_s;               /* sample_qeq, csa_sampling_strategy.hpp:660-665 */
    uint64_t result = o->isa_s[ci], pos = ci * 64;
    uint64_t steps = pos < i ? pos + o->size - i : pos - i;
    while (steps--) result = lf(o, result);
    return result;
}

/* bwt[i] = wt_int::operator[], sdsl/wt_int.hpp:322-343 */
static uint64_t bwt_at(const fmo_t* o, uint64_t i) { uint64_t c; (void)inverse_select(o, i, &c); return c; }

/* FMIndex::extract_text, fm_index.cpp:169-184 */
uint64_t fmo_extract_text(const fmo_t* o, uint64_t begin, uint64_t end, uint64_t* out, uint64_t cap) {
    uint64_t len = 0;
    if (end - begin == 0) return 0;
    uint64_t start = isa_at(o, end);
    uint64_t symbol = bwt_at(o, start);
    if (len < cap) out[len] = symbol;
    len++;
    if (end - begin == 1) return len;
    for (uint64_t i = 0; i < end - begin - 1; i++) {
        uint64_t lr[2];
        fmo_backward_search_step(o, symbol, start, start + 1, lr);
        start = lr[0];
        symbol = bwt_at(o, start);
        if (len < cap) out[len] = symbol;
        len++;
    }
    return len;
}

const uint64_t* fmo_tree_words(const fmo_t* o, uint64_t* n) { *n = o->tree_words; return o->tree; }
const uint64_t* fmo_rank_blocks(const fmo_t* o, uint64_t* n) { *n = o->bb_words; return o->bb; }
const uint64_t* fmo_sa_samples(const fmo_t* o, uint64_t* n) { *n = o->n_sa_s; return o->sa_s; }
const uint64_t* fmo_isa_samples(const fmo_t* o, uint64_t* n) { *n = o->n_isa_s; return o->isa_s; }
const uint64_t* fmo_alphabet(const fmo_t* o, uint64_t* s) { *s = o->sigma; return o->alpha; }
const uint64_t* fmo_C(const fmo_t* o, uint64_t* n) { *n = o->sigma + 1; return o->C; }
const uint64_t* fmo_bwt(const fmo_t* o, uint64_t* n) { *n = o->size; return o->bwt; }
const uint64_t* fmo_sa(const fmo_t* o, uint64_t* n) { *n = o->size; return o->sa; }
