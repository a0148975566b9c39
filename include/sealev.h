/* sealev.h -- native host logic of SEAL's evidence aggregation (seal/keys.py:178-497).
 *
 * The reference implements this stage in Python on top of per-call FM-index queries.  seal_b200.keys.
 * aggregate_evidence keeps the reference's signature, batches every index access onto the GPU
 * (sealfm_backward_search_multi, sealfm_locate, sealfm_extract_text) and hands the two order-defining loops
 * that remain to these functions: plain C++ on doubles in the reference's evaluation order (results are
 * bit-identical to the reference function, tests/golden/keys_golden.json).  Host-only: no CUDA calls.
 *
 * Keys are passed flattened: key k = key_tok[key_off[k] .. key_off[k+1]).  Returns 0 or a SEALFM_E* code
 * (sealev_last_error() for the message). */
#ifndef SEALEV_H
#define SEALEV_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* sealev_last_error(void);
/* multi_key_score is Python's built-in sum() over the picked keys' scores (seal/keys.py:476): Neumaier-compensated since
 * CPython 3.12, a plain left-to-right sum before (the reference pins Python 3.8/3.9).  compensated != 0 selects the
 * former (default); seal_b200.keys sets it from the running interpreter so "the reference's result" follows the
 * interpreter the reference would run under. */
void sealev_set_sum_mode(int compensated);

/* First stage (seal/keys.py:316-368): walks the located occurrences of the rare keys in key order --
 * occurrence j of key k (span_off[k] <= j < span_off[k+1]) ends at token position pos[j] in document doc[j] --
 * credits each document once per key unless the occurrence overlaps positions claimed by an earlier key
 * (allow_overlaps lifts that), damps repeated token types per document with `beta`, and returns the
 * max_docs best documents (stable order of first touch among ties).  key_count[k] = corpus count of key k,
 * empty_count = count of the empty key (len(index)); sort_mode 0 = by score, 1 = sort_by_length,
 * 2 = sort_by_freq (only the tie rule of a document's best key, which single_key weighs in). */
int sealev_first_stage(int64_t n_keys, const int64_t* key_tok, const int64_t* key_off, const double* key_score,
                       const int64_t* key_count, int64_t empty_count, const int64_t* span_off,
                       const uint64_t* pos, const int64_t* doc, int32_t sort_mode, int32_t allow_overlaps, double beta,
                       double single_key, int64_t max_docs, int64_t* out_docs, int64_t* out_n);

/* Full scoring of the shortlisted documents (seal/keys.py:378-491): document d = doc_tok[doc_off[d] ..
 * doc_off[d+1]); finds every occurrence of every scored key (trie scan in the reference's discovery order),
 * places keys greedily by (score, key, position), adds the unigram scores of uncovered token types
 * (unigram may be NULL).  Outputs per document: out_score, out_best (key index or -1) / out_best_score, and
 * the picked keys pick_key[pick_off[d] .. pick_off[d+1]) with their damped scores; a pick < 0 is the
 * unigram of token -(pick) - 1.  pick_cap = capacity of pick_key / pick_score (SEALFM_ECAPACITY if short). */
int sealev_score_docs(int64_t n_keys, const int64_t* key_tok, const int64_t* key_off, const double* key_score,
                      const int64_t* key_count, int64_t empty_count, int64_t n_docs, const int64_t* doc_tok,
                      const int64_t* doc_off, const double* unigram, int64_t n_unigram, int32_t sort_mode,
                      int32_t allow_overlaps, int32_t ignore_free_places, int32_t single_key_add_unigrams, double beta,
                      double single_key, double* out_score, int64_t* out_best, double* out_best_score,
                      int64_t* pick_off, int64_t* pick_key, double* pick_score, int64_t pick_cap);

#ifdef __cplusplus
}
#endif
#endif
