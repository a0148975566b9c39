/* TEST INFRASTRUCTURE ONLY — the product (seal_b200/, include/) never includes, links or calls this.
 *
 * Plain-C CPU restatement of the FM-index arithmetic the reference runs through sdsl-lite 2.1.0
 * (vendored at /root/reference/res/external/sdsl-lite, abbreviated "sdsl/" below) and
 * seal/cpp_modules/fm_index.cpp.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it, and only as the checker.
 *
 * Parity status: the reference ships no golden vectors for this path; this
 * restatement is pinned against the reference ITSELF compiled from its own sources
 * (oracle/_ref/libseal_ref.so, see oracle/Makefile) in tests/test_oracle_vs_ref.py, and against
 * fixtures that library generated (tests/golden/, generator: tests/golden/make_golden.py).
 */
#ifndef SEAL_FM_ORACLE_H
#define SEAL_FM_ORACLE_H
#include <stdint.h>

typedef struct fmo fmo_t;

/* text: n symbols, all > 0 (SEAL passes token+10, seal/index.py:16,52).  A 0 sentinel is appended
 * as sdsl::construct does (sdsl/construct.hpp:48-52).  Returns NULL on allocation failure. */
fmo_t*   fmo_build(const uint64_t* text, uint64_t n);
void     fmo_free(fmo_t* o);

uint64_t fmo_size(const fmo_t* o);        /* n+1                       fm_index.cpp:50-52        */
uint64_t fmo_sigma(const fmo_t* o);       /* distinct symbols incl. 0  sdsl/wt_int.hpp:236       */
uint32_t fmo_max_level(const fmo_t* o);   /*                           sdsl/wt_int.hpp:189-193   */

uint64_t fmo_bv_rank(const fmo_t* o, uint64_t idx);            /* sdsl/rank_support_v.hpp:114-124 */
uint64_t fmo_wt_rank(const fmo_t* o, uint64_t i, uint64_t c);  /* sdsl/wt_int.hpp:356-380         */

/* sdsl/suffix_array_algorithm.hpp:163-191 wrapped as fm_index.cpp:67-76: inclusive hi in/out */
void     fmo_backward_search_step(const fmo_t* o, uint64_t sym, uint64_t lo, uint64_t hi,
                                  uint64_t out[2]);
/* fm_index.cpp:55-65: fold over q from (0,size()); returns {l, r+1} */
void     fmo_backward_search_multi(const fmo_t* o, const uint64_t* q, uint64_t n, uint64_t out[2]);

/* fm_index.cpp:91-109: interleaved (sym,count), ascending sym; returns the vector length (2k).
 * Writes at most cap entries. */
uint64_t fmo_distinct_count(const fmo_t* o, uint64_t lo, uint64_t hi, uint64_t* out, uint64_t cap);
/* instrumentation: internal wavelet-tree nodes the expansion of [lo,hi) visits,
 * N_b = sum over levels of the number of distinct symbol prefixes (L when one symbol). */
uint64_t fmo_visited_nodes(const fmo_t* o, uint64_t lo, uint64_t hi);

uint64_t fmo_locate(const fmo_t* o, uint64_t row);             /* fm_index.cpp:163-167            */
/* fm_index.cpp:169-184; returns the vector length */
uint64_t fmo_extract_text(const fmo_t* o, uint64_t begin, uint64_t end, uint64_t* out, uint64_t cap);

/* raw sections, for byte-level comparison with the product's builder and with sdsl's .fmi */
const uint64_t* fmo_tree_words(const fmo_t* o, uint64_t* n_words);      /* excl. the padding word */
const uint64_t* fmo_rank_blocks(const fmo_t* o, uint64_t* n_words);
const uint64_t* fmo_sa_samples(const fmo_t* o, uint64_t* n);
const uint64_t* fmo_isa_samples(const fmo_t* o, uint64_t* n);
const uint64_t* fmo_alphabet(const fmo_t* o, uint64_t* sigma);          /* ascending symbols       */
const uint64_t* fmo_C(const fmo_t* o, uint64_t* n);                     /* sigma+1 entries         */
const uint64_t* fmo_bwt(const fmo_t* o, uint64_t* n);                   /* explicit BWT (n+1)      */
const uint64_t* fmo_sa(const fmo_t* o, uint64_t* n);                    /* explicit SA (n+1)       */

#endif
