/* TEST INFRASTRUCTURE (oracle build only).
 * nasm is not installed, so the x264-derived NASM entry points the reference
 * declares at /root/reference/src/SADFunctions.cpp:264-325 and CPU.h:47-49 are
 * provided here by forwarding to the reference's OWN portable C kernels
 * (selectSADFunction/selectSATDFunction with opt=0 -> sad_c / Satd_C,
 * SADFunctions.cpp:516-567,864-920).  x264's checkasm guarantees asm == C. */
#include <stdint.h>
#include <cpuid.h>
#include <immintrin.h>
#include "SADFunctions.h"

extern "C" {

void mvtools_cpu_emms() {}

void mvtools_cpu_cpuid(uint32_t op, uint32_t *eax, uint32_t *ebx, uint32_t *ecx, uint32_t *edx) {
    __cpuid_count(op, 0, *eax, *ebx, *ecx, *edx);
}

uint64_t mvtools_cpu_xgetbv(int xcr) {
    uint32_t lo, hi;
    __asm__ volatile("xgetbv" : "=a"(lo), "=d"(hi) : "c"(xcr));
    return ((uint64_t)hi << 32) | lo;
}

#define SAD_SHIM(w, h, isa) \
    unsigned int mvtools_pixel_sad_##w##x##h##_##isa(const uint8_t *pSrc, intptr_t nSrcPitch, const uint8_t *pRef, intptr_t nRefPitch) { \
        static SADFunction f = selectSADFunction(w, h, 8, 0, 0); \
        return f(pSrc, nSrcPitch, pRef, nRefPitch); \
    }
#define SATD_SHIM(w, h, isa) \
    unsigned int mvtools_pixel_satd_##w##x##h##_##isa(const uint8_t *pSrc, intptr_t nSrcPitch, const uint8_t *pRef, intptr_t nRefPitch) { \
        static SADFunction f = selectSATDFunction(w, h, 8, 0, 0); \
        return f(pSrc, nSrcPitch, pRef, nRefPitch); \
    }

SAD_SHIM(4, 4, mmx2) SAD_SHIM(4, 8, mmx2) SAD_SHIM(8, 4, mmx2) SAD_SHIM(8, 8, mmx2) SAD_SHIM(8, 16, mmx2)
SAD_SHIM(16, 8, mmx2) SAD_SHIM(16, 16, mmx2)
SAD_SHIM(8, 4, cache64_mmx2) SAD_SHIM(8, 8, cache64_mmx2) SAD_SHIM(8, 16, cache64_mmx2)
SAD_SHIM(8, 16, sse2) SAD_SHIM(16, 8, sse2) SAD_SHIM(16, 16, sse2)
SAD_SHIM(16, 8, sse3) SAD_SHIM(16, 16, sse3)
SAD_SHIM(16, 8, cache64_ssse3) SAD_SHIM(16, 16, cache64_ssse3)

SATD_SHIM(4, 4, mmx2)
SATD_SHIM(8, 4, sse2) SATD_SHIM(8, 8, sse2) SATD_SHIM(16, 8, sse2) SATD_SHIM(16, 16, sse2)
SATD_SHIM(4, 4, ssse3) SATD_SHIM(8, 4, ssse3) SATD_SHIM(8, 8, ssse3) SATD_SHIM(16, 8, ssse3) SATD_SHIM(16, 16, ssse3)
SATD_SHIM(4, 4, sse4) SATD_SHIM(8, 4, sse4) SATD_SHIM(8, 8, sse4) SATD_SHIM(16, 8, sse4) SATD_SHIM(16, 16, sse4)
SATD_SHIM(4, 4, avx) SATD_SHIM(8, 4, avx) SATD_SHIM(8, 8, avx) SATD_SHIM(16, 8, avx) SATD_SHIM(16, 16, avx)
SATD_SHIM(4, 4, xop) SATD_SHIM(8, 4, xop) SATD_SHIM(8, 8, xop) SATD_SHIM(16, 8, xop) SATD_SHIM(16, 16, xop)
SATD_SHIM(8, 8, avx2) SATD_SHIM(16, 8, avx2) SATD_SHIM(16, 16, avx2)

} // extern "C"
