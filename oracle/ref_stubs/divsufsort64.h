/* TEST INFRASTRUCTURE ONLY. See divsufsort.h in this directory. */
#ifndef ORACLE_DIVSUFSORT64_STUB_H
#define ORACLE_DIVSUFSORT64_STUB_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
int64_t divsufsort64(const uint8_t* T, int64_t* SA, int64_t n);
#ifdef __cplusplus
}
#endif
#endif
