/* TEST INFRASTRUCTURE (oracle build only): minimal stand-in for FFTW3's single
 * precision API used by /root/reference/src/DCTFFTW.cpp:151-206.  FFTW itself is
 * not vendored by the reference and not installed here (no version pin in
 * meson.build:185-188); oracle/shim/fftw_shim.c evaluates the published
 * REDFT10 definition directly, so dct=1..4 parity is "unpinned" (DESIGN.md). */
#ifndef FFTW3_SHIM_H
#define FFTW3_SHIM_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct fftwf_plan_s *fftwf_plan;
typedef enum { FFTW_R2HC = 0, FFTW_HC2R = 1, FFTW_DHT = 2, FFTW_REDFT00 = 3, FFTW_REDFT01 = 4, FFTW_REDFT10 = 5, FFTW_REDFT11 = 6 } fftwf_r2r_kind;
#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)
void *fftwf_malloc(size_t n);
void fftwf_free(void *p);
fftwf_plan fftwf_plan_r2r_2d(int n0, int n1, float *in, float *out, fftwf_r2r_kind kind0, fftwf_r2r_kind kind1, unsigned flags);
void fftwf_execute_r2r(const fftwf_plan p, float *in, float *out);
void fftwf_destroy_plan(fftwf_plan p);
#ifdef __cplusplus
}
#endif
#endif
