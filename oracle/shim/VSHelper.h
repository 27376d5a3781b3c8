/* TEST INFRASTRUCTURE (oracle build only): stand-in for the API-3 VSHelper.h that
 * /root/reference/src/MVFrame.cpp:22 and SimpleResize.cpp:9 include.  Only vs_bitblt and
 * VSMIN/VSMAX are needed. */
#ifndef VSHELPER_H
#define VSHELPER_H
#include <stdlib.h>
#include <math.h>
#include <limits.h>
#include <assert.h>
#include <stdint.h>
#include <string.h>
#include <stddef.h>
#ifndef VSMAX
#define VSMAX(a, b) ((a) > (b) ? (a) : (b))
#define VSMIN(a, b) ((a) > (b) ? (b) : (a))
#endif
static inline void vs_bitblt(void *dstp, int dst_stride, const void *srcp, int src_stride, int row_size, int height) {
    const uint8_t *s = (const uint8_t *)srcp;
    uint8_t *d = (uint8_t *)dstp;
    for (int i = 0; i < height; i++) {
        memcpy(d, s, (size_t)row_size);
        s += src_stride;
        d += dst_stride;
    }
}
#endif
