/*
 * Helper functions accompanying include/vs/VapourSynth4.h (restated; see the
 * note at the top of that file).  Only what the mv.* hot-path filters use.
 */
#ifndef VSHELPER4_H
#define VSHELPER4_H

#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include "VapourSynth4.h"

#define VSH_STD_PLANE_ALIGNMENT 64

#ifdef _WIN32
#include <malloc.h>
#define VSH_ALIGNED_MALLOC(pptr, size, alignment) do { *(pptr) = _aligned_malloc((size), (alignment)); } while (0)
#define VSH_ALIGNED_FREE(ptr) do { _aligned_free((ptr)); } while (0)
#else
#define VSH_ALIGNED_MALLOC(pptr, size, alignment) do { if (posix_memalign((void **)(pptr), (alignment), (size))) *((void **)(pptr)) = NULL; } while (0)
#define VSH_ALIGNED_FREE(ptr) do { free((ptr)); } while (0)
#endif

#define VSMAX(a, b) ((a) > (b) ? (a) : (b))
#define VSMIN(a, b) ((a) > (b) ? (b) : (a))

#ifdef __cplusplus
#define VSH4_MANGLE(n) n
namespace vsh {
#else
#define VSH4_MANGLE(n) vsh_##n
#endif

static inline int VSH4_MANGLE(isConstantVideoFormat)(const VSVideoInfo *vi) {
    return vi->height > 0 && vi->width > 0 && vi->format.colorFamily != cfUndefined;
}

static inline int VSH4_MANGLE(isSameVideoFormat)(const VSVideoFormat *v1, const VSVideoFormat *v2) {
    return v1->colorFamily == v2->colorFamily && v1->sampleType == v2->sampleType && v1->bitsPerSample == v2->bitsPerSample && v1->subSamplingW == v2->subSamplingW && v1->subSamplingH == v2->subSamplingH;
}

static inline void VSH4_MANGLE(bitblt)(void *dstp, ptrdiff_t dst_stride, const void *srcp, ptrdiff_t src_stride, size_t row_size, size_t height) {
    if (height) {
        if (src_stride == dst_stride && src_stride == (ptrdiff_t)row_size) {
            memcpy(dstp, srcp, row_size * height);
        } else {
            const uint8_t *srcp8 = (const uint8_t *)srcp;
            uint8_t *dstp8 = (uint8_t *)dstp;
            size_t i;
            for (i = 0; i < height; i++) {
                memcpy(dstp8, srcp8, row_size);
                srcp8 += src_stride;
                dstp8 += dst_stride;
            }
        }
    }
}

#ifdef __cplusplus
}
#endif

#endif /* VSHELPER4_H */
