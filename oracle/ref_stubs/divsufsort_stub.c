/* TEST INFRASTRUCTURE ONLY. Byte-alphabet suffix sorters are unreachable from SEAL's
 * csa_wt_int<> path; abort loudly if that assumption is ever wrong. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
int32_t divsufsort(const uint8_t* T, int32_t* SA, int32_t n) {
    (void)T; (void)SA; (void)n; fprintf(stderr, "oracle/_ref: divsufsort stub reached\n"); abort();
}
int64_t divsufsort64(const uint8_t* T, int64_t* SA, int64_t n) {
    (void)T; (void)SA; (void)n; fprintf(stderr, "oracle/_ref: divsufsort64 stub reached\n"); abort();
}
