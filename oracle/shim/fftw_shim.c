/* TEST INFRASTRUCTURE (oracle build only).
 * Direct evaluation of FFTW's REDFT10 x REDFT10 (unnormalised 2-D DCT-II):
 *   Y[k1][k2] = 4 * sum_i sum_j x[i][j] cos(pi (2i+1) k1 / (2 n0)) cos(pi (2j+1) k2 / (2 n1))
 * as documented by FFTW ("1d Real-even DFTs": REDFT10 Y_k = 2 sum x_j cos(pi (j+1/2) k / n)).
 * Replaces the un-vendored fftw3f used at /root/reference/src/DCTFFTW.cpp:170,206.
 * Summation order differs from FFTW's codelets => float results may differ in
 * the last ulp; dct=1..4 parity is therefore tolerance-based ("unpinned"). */
#include <math.h>
#include <stdlib.h>
#include "fftw3.h"

struct fftwf_plan_s {
    int n0, n1;
    float *c0; /* n0 x n0: c0[k*n0+i] = 2 cos(pi (2i+1) k / (2 n0)) */
    float *c1;
    float *tmp;
};

void *fftwf_malloc(size_t n) {
    void *p = NULL;
    if (posix_memalign(&p, 64, n ? n : 64))
        return NULL;
    return p;
}

void fftwf_free(void *p) { free(p); }

static float *costab(int n) {
    float *t = (float *)malloc(sizeof(float) * (size_t)n * (size_t)n);
    for (int k = 0; k < n; k++)
        for (int i = 0; i < n; i++)
            t[k * n + i] = (float)(2.0 * cos(M_PI * (2.0 * i + 1.0) * k / (2.0 * n)));
    return t;
}

fftwf_plan fftwf_plan_r2r_2d(int n0, int n1, float *in, float *out, fftwf_r2r_kind kind0, fftwf_r2r_kind kind1, unsigned flags) {
    (void)in; (void)out; (void)flags;
    if (kind0 != FFTW_REDFT10 || kind1 != FFTW_REDFT10)
        return NULL;
    fftwf_plan p = (fftwf_plan)malloc(sizeof(*p));
    p->n0 = n0;
    p->n1 = n1;
    p->c0 = costab(n0);
    p->c1 = costab(n1);
    p->tmp = (float *)malloc(sizeof(float) * (size_t)n0 * (size_t)n1);
    return p;
}

void fftwf_execute_r2r(const fftwf_plan p, float *in, float *out) {
    int n0 = p->n0, n1 = p->n1;
    /* rows (dimension 1) */
    for (int i = 0; i < n0; i++)
        for (int k = 0; k < n1; k++) {
            float s = 0.f;
            for (int j = 0; j < n1; j++)
                s += in[i * n1 + j] * p->c1[k * n1 + j];
            p->tmp[i * n1 + k] = s;
        }
    /* columns (dimension 0) */
    for (int k = 0; k < n0; k++)
        for (int j = 0; j < n1; j++) {
            float s = 0.f;
            for (int i = 0; i < n0; i++)
                s += p->tmp[i * n1 + j] * p->c0[k * n0 + i];
            out[k * n1 + j] = s;
        }
}

void fftwf_destroy_plan(fftwf_plan p) {
    if (!p)
        return;
    free(p->c0);
    free(p->c1);
    free(p->tmp);
    free(p);
}
