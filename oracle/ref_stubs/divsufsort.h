/* TEST INFRASTRUCTURE ONLY. Stand-in declaration for libdivsufsort's generated header (the real
 * one is produced by the reference's cmake from divsufsort.h.cmake, which we do not run).
 * SEAL's integer-alphabet index never reaches divsufsort (sdsl/construct_sa.hpp:150-166 uses
 * qsufsort for t_width != 8); the definitions in divsufsort_stub.c abort if ever called. */
#ifndef ORACLE_DIVSUFSORT_STUB_H
#define ORACLE_DIVSUFSORT_STUB_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
int32_t divsufsort(const uint8_t* T, int32_t* SA, int32_t n);
#ifdef __cplusplus
}
#endif
#endif
