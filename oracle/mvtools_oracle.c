/*
 * TEST INFRASTRUCTURE -- CPU oracle: a plain-C restatement of the per-frame
 * algorithms of the mv.* hot path of dubhater/vapoursynth-mvtools (e516e90).
 * Every function cites the reference file:line it follows.  It exists to check
 * the CUDA path; nothing in the product links, loads or calls it.
 *
 * PINNING: the reference has no tests or golden vectors, so
 * this restatement is pinned against the reference ITSELF: tests/test_oracle_vs_ref.py
 * runs oracle/_ref/libmvtools_ref.so (the unmodified reference sources, compiled in
 * place by oracle/Makefile) on seeded clips and requires byte-identical super
 * frames, vector blobs and output planes; hashes of those runs are committed under
 * tests/golden/ so the check also runs where /root/reference is absent.
 * Not restated: dct=1..4 (FFTW-based DCT; "parity unpinned").
 */
#include "mvtools_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static __thread char g_err[512];

const char *mvo_last_error(void) { return g_err; }

static int fail(const char *msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return 1;
}

#define IMIN(a, b) ((a) < (b) ? (a) : (b))
#define IMAX(a, b) ((a) > (b) ? (a) : (b))

static inline int LD(const uint8_t *p, int bps, ptrdiff_t i) {
    return bps == 1 ? p[i] : ((const uint16_t *)p)[i];
}

static inline void ST(uint8_t *p, int bps, ptrdiff_t i, int v) {
    if (bps == 1)
        p[i] = (uint8_t)v;
    else
        ((uint16_t *)p)[i] = (uint16_t)v;
}

/* ======================================================================== */
/* geometry                                                                  */

/* PlaneHeightLuma / PlaneWidthLuma, MVFrame.cpp:1209-1226 */
static int plane_dim_luma(int src, int level, int ratio, int pad) {
    int d = src;
    for (int i = 1; i <= level; i++)
        d = pad >= ratio ? ((d / ratio + 1) / 2) * ratio : ((d / ratio) / 2) * ratio;
    return d;
}

/* PlaneSuperOffset, MVFrame.cpp:1229-1247, expressed in rows */
static int plane_super_row(int chroma, int src_height, int level, int pel, int vpad, int yRatioUV) {
    int height = src_height;
    int rows;
    if (level == 0)
        return 0;
    rows = pel * pel * (src_height + vpad * 2);
    for (int i = 1; i < level; i++) {
        height = chroma ? plane_dim_luma(src_height * yRatioUV, i, yRatioUV, vpad * yRatioUV) / yRatioUV
                        : plane_dim_luma(src_height, i, yRatioUV, vpad);
        rows += height + vpad * 2;
    }
    return rows;
}

typedef struct Lv {
    int w, h, hpad, vpad, pw, ph, pel, row0;
} Lv;

/* mvgofInit + mvfInit + mvgofUpdate, MVFrame.cpp:1758-1779,1852-1903 */
static void level_geom(int width, int height, int xR, int yR, int hpad, int vpad, int pel, int plane, int level, Lv *g) {
    int w0 = plane ? width / xR : width;
    int h0 = plane ? height / yR : height;
    int hp = plane ? hpad / xR : hpad;
    int vp = plane ? vpad / yR : vpad;
    if (level == 0) {
        g->w = w0;
        g->h = h0;
    } else {
        int wl = plane_dim_luma(width, level, xR, hpad);
        int hl = plane_dim_luma(height, level, yR, vpad);
        g->w = plane ? wl / xR : wl;
        g->h = plane ? hl / yR : hl;
    }
    g->hpad = hp;
    g->vpad = vp;
    g->pw = g->w + 2 * hp;
    g->ph = g->h + 2 * vp;
    g->pel = level == 0 ? pel : 1;
    g->row0 = plane_super_row(plane != 0, h0, level, pel, vp, yR);
}

/* ======================================================================== */
/* Super: reduce / pad / refine                                              */

/* RB2F_C, MVFrame.cpp:576-595 */
static void rb2f(uint8_t *d, const uint8_t *s, ptrdiff_t dp, ptrdiff_t sp, int w, int h, int bps) {
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++)
            ST(d, bps, x, (LD(s, bps, x * 2) + LD(s, bps, x * 2 + 1) + LD(s, bps, x * 2 + sp + 1) + LD(s, bps, x * 2 + sp) + 2) / 4);
        d += dp * bps;
        s += sp * 2 * bps;
    }
}

/* vertical passes: RB2FilteredVertical :598-643, RB2BilinearFilteredVertical :696-735,
 * RB2QuadraticVertical :782-836, RB2CubicVertical :885-939 */
static void reduce_vertical(uint8_t *d, const uint8_t *s, ptrdiff_t dp, ptrdiff_t sp, int w, int h, int bps, int rfilter) {
    for (int y = 0; y < h; y++) {
        int edge;
        if (rfilter == 1)
            edge = (y == 0);
        else
            edge = (y == 0) || (y >= IMAX(h - 1, 1)); /* first and last row: plain average */
        for (int x = 0; x < w; x++) {
            int v;
            if (edge) {
                v = (LD(s, bps, x) + LD(s, bps, x + sp) + 1) / 2;
            } else if (rfilter == 1) {
                v = (LD(s, bps, x - sp) + LD(s, bps, x) * 2 + LD(s, bps, x + sp) + 2) / 4;
            } else if (rfilter == 2) {
                v = (LD(s, bps, x - sp) + (LD(s, bps, x) + LD(s, bps, x + sp)) * 3 + LD(s, bps, x + sp * 2) + 4) / 8;
            } else {
                int m0 = LD(s, bps, x - sp * 2), m1 = LD(s, bps, x - sp), m2 = LD(s, bps, x);
                int m3 = LD(s, bps, x + sp), m4 = LD(s, bps, x + sp * 2), m5 = LD(s, bps, x + sp * 3);
                if (rfilter == 3) {
                    m2 = (m2 + m3) * 22;
                    m1 = (m1 + m4) * 9;
                    m0 += m5 + m2 + m1 + 32;
                    m0 >>= 6;
                } else {
                    m2 = (m2 + m3) * 10;
                    m1 = (m1 + m4) * 5;
                    m0 += m5 + m2 + m1 + 16;
                    m0 >>= 5;
                }
                v = m0;
            }
            ST(d, bps, x, v);
        }
        d += dp * bps;
        s += sp * 2 * bps;
    }
}

/* in-place horizontal passes: :646-685, :739-768, :840-881, :943-984 */
static void reduce_horizontal_inplace(uint8_t *s, ptrdiff_t sp, int w, int h, int bps, int rfilter) {
    for (int y = 0; y < h; y++) {
        int s0 = (LD(s, bps, 0) + LD(s, bps, 1) + 1) / 2;
        if (rfilter == 1) {
            for (int x = 1; x < w; x++)
                ST(s, bps, x, (LD(s, bps, x * 2 - 1) + LD(s, bps, x * 2) * 2 + LD(s, bps, x * 2 + 1) + 2) / 4);
            ST(s, bps, 0, s0);
        } else {
            for (int x = 1; x < w - 1; x++) {
                int v;
                if (rfilter == 2) {
                    v = (LD(s, bps, x * 2 - 1) + (LD(s, bps, x * 2) + LD(s, bps, x * 2 + 1)) * 3 + LD(s, bps, x * 2 + 2) + 4) / 8;
                } else {
                    int m0 = LD(s, bps, x * 2 - 2), m1 = LD(s, bps, x * 2 - 1), m2 = LD(s, bps, x * 2);
                    int m3 = LD(s, bps, x * 2 + 1), m4 = LD(s, bps, x * 2 + 2), m5 = LD(s, bps, x * 2 + 3);
                    if (rfilter == 3) {
                        m2 = (m2 + m3) * 22;
                        m1 = (m1 + m4) * 9;
                        m0 += m5 + m2 + m1 + 32;
                        m0 >>= 6;
                    } else {
                        m2 = (m2 + m3) * 10;
                        m1 = (m1 + m4) * 5;
                        m0 += m5 + m2 + m1 + 16;
                        m0 >>= 5;
                    }
                    v = m0;
                }
                ST(s, bps, x, v);
            }
            ST(s, bps, 0, s0);
            for (int x = IMAX(w - 1, 1); x < w; x++)
                ST(s, bps, x, (LD(s, bps, x * 2) + LD(s, bps, x * 2 + 1) + 1) / 2);
        }
        s += sp * bps;
    }
}

/* mvpReduceTo, MVFrame.cpp:1632-1681: vertical pass at DOUBLE width into the
 * destination, then in-place horizontal (the stale right half stays in the frame) */
static void reduce_plane(uint8_t *d, const uint8_t *s, ptrdiff_t pitch, int w, int h, int bps, int rfilter) {
    if (rfilter == 0) {
        rb2f(d, s, pitch, pitch, w, h, bps);
        return;
    }
    reduce_vertical(d, s, pitch, pitch, w * 2, h, bps, rfilter);
    reduce_horizontal_inplace(d, pitch, w, h, bps, rfilter);
}

/* PadReferenceFrame, MVFrame.cpp:1264-1318 (p = padded origin, pitch in samples) */
static void pad_plane(uint8_t *p, ptrdiff_t pitch, int hpad, int vpad, int w, int h, int bps) {
    for (int y = 0; y < h + 2 * vpad; y++) {
        int sy = IMIN(IMAX(y - vpad, 0), h - 1);
        const uint8_t *srow = p + (ptrdiff_t)(sy + vpad) * pitch * bps;
        uint8_t *drow = p + (ptrdiff_t)y * pitch * bps;
        int interior_row = (y >= vpad && y < vpad + h);
        for (int x = 0; x < w + 2 * hpad; x++) {
            int interior_col = (x >= hpad && x < hpad + w);
            if (interior_row && interior_col)
                continue;
            int sx = IMIN(IMAX(x - hpad, 0), w - 1);
            ST(drow, bps, x, LD(srow, bps, sx + hpad));
        }
    }
}

/* VerticalBilinear :508-526, VerticalBicubic :1115-1147, VerticalWiener :1019-1063 */
static void refine_vertical(uint8_t *d, const uint8_t *s, ptrdiff_t pitch, int w, int h, int bps, int bits, int sharp) {
    int pixelMax = (1 << bits) - 1;
    for (int j = 0; j < h; j++) {
        const uint8_t *sr = s + (ptrdiff_t)j * pitch * bps;
        uint8_t *dr = d + (ptrdiff_t)j * pitch * bps;
        int mode; /* 0 copy, 1 average, 2 filter */
        if (j == h - 1)
            mode = 0;
        else if (sharp == 0)
            mode = 1;
        else if (sharp == 1)
            mode = (j >= 1 && j < h - 3) ? 2 : 1;
        else
            mode = (j >= 2 && j < h - 4) ? 2 : 1;
        for (int i = 0; i < w; i++) {
            int v;
            if (mode == 0)
                v = LD(sr, bps, i);
            else if (mode == 1)
                v = (LD(sr, bps, i) + LD(sr, bps, i + pitch) + 1) >> 1;
            else if (sharp == 1) {
                v = (-LD(sr, bps, i - pitch) - LD(sr, bps, i + pitch * 2) + (LD(sr, bps, i) + LD(sr, bps, i + pitch)) * 9 + 8) >> 4;
                v = IMIN(pixelMax, IMAX(0, v));
            } else {
                int m0 = LD(sr, bps, i - pitch * 2), m1 = LD(sr, bps, i - pitch), m2 = LD(sr, bps, i);
                int m3 = LD(sr, bps, i + pitch), m4 = LD(sr, bps, i + pitch * 2), m5 = LD(sr, bps, i + pitch * 3);
                m2 = (m2 + m3) * 4;
                m2 -= m1 + m4;
                m2 *= 5;
                m0 += m5 + m2 + 16;
                m0 >>= 5;
                v = IMAX(0, IMIN(m0, pixelMax));
            }
            ST(dr, bps, i, v);
        }
    }
}

/* HorizontalBilinear :529-546, HorizontalBicubic :1150-1176, HorizontalWiener :1066-1111 */
static void refine_horizontal(uint8_t *d, const uint8_t *s, ptrdiff_t pitch, int w, int h, int bps, int bits, int sharp) {
    int pixelMax = (1 << bits) - 1;
    for (int j = 0; j < h; j++) {
        const uint8_t *sr = s + (ptrdiff_t)j * pitch * bps;
        uint8_t *dr = d + (ptrdiff_t)j * pitch * bps;
        for (int i = 0; i < w; i++) {
            int mode, v;
            if (i == w - 1)
                mode = 0;
            else if (sharp == 0)
                mode = 1;
            else if (sharp == 1)
                mode = (i >= 1 && i < w - 3) ? 2 : 1;
            else
                mode = (i >= 2 && i < w - 4) ? 2 : 1;
            if (mode == 0)
                v = LD(sr, bps, i);
            else if (mode == 1)
                v = (LD(sr, bps, i) + LD(sr, bps, i + 1) + 1) >> 1;
            else if (sharp == 1) {
                v = (-(LD(sr, bps, i - 1) + LD(sr, bps, i + 2)) + (LD(sr, bps, i) + LD(sr, bps, i + 1)) * 9 + 8) >> 4;
                v = IMIN(pixelMax, IMAX(0, v));
            } else {
                int m0 = LD(sr, bps, i - 2), m1 = LD(sr, bps, i - 1), m2 = LD(sr, bps, i);
                int m3 = LD(sr, bps, i + 1), m4 = LD(sr, bps, i + 2), m5 = LD(sr, bps, i + 3);
                m2 = (m2 + m3) * 4;
                m2 -= m1 + m4;
                m2 *= 5;
                m0 += m5 + m2 + 16;
                m0 >>= 5;
                v = IMAX(0, IMIN(m0, pixelMax));
            }
            ST(dr, bps, i, v);
        }
    }
}

/* DiagonalBilinear, MVFrame.cpp:549-572 */
static void refine_diagonal_bilinear(uint8_t *d, const uint8_t *s, ptrdiff_t pitch, int w, int h, int bps) {
    for (int j = 0; j < h; j++) {
        const uint8_t *sr = s + (ptrdiff_t)j * pitch * bps;
        uint8_t *dr = d + (ptrdiff_t)j * pitch * bps;
        for (int i = 0; i < w; i++) {
            int v;
            if (j < h - 1) {
                if (i < w - 1)
                    v = (LD(sr, bps, i) + LD(sr, bps, i + 1) + LD(sr, bps, i + pitch) + LD(sr, bps, i + pitch + 1) + 2) >> 2;
                else
                    v = (LD(sr, bps, i) + LD(sr, bps, i + pitch) + 1) >> 1;
            } else {
                if (i < w - 1)
                    v = (LD(sr, bps, i) + LD(sr, bps, i + 1) + 1) >> 1;
                else
                    v = LD(sr, bps, i);
            }
            ST(dr, bps, i, v);
        }
    }
}

/* Average2, MVFrame.cpp:1180-1197 */
static void average2(uint8_t *d, const uint8_t *a, const uint8_t *b, ptrdiff_t pitch, int w, int h, int bps) {
    for (int j = 0; j < h; j++) {
        for (int i = 0; i < w; i++)
            ST(d, bps, i, (LD(a, bps, i) + LD(b, bps, i) + 1) >> 1);
        d += pitch * bps;
        a += pitch * bps;
        b += pitch * bps;
    }
}

/* The 3/4-pel planes of an 8-bit clip, with the reference's default opt=1 on x86-64: Average2 over nPaddedWidth - 1 columns
 * (MVFrame.cpp:1519-1526) runs Average2_avx2 (MVFrame.cpp:1496-1503, MVFrame_AVX2.cpp:19-33), whose 32-byte steps also write
 * the LAST column unless nPaddedWidth - 1 is a multiple of 32: avg(the sample after plane a's row, b's last sample).  The
 * sample after the row is stride slack that mvsuper's memset and the SIMD passes leave at 0 (a frame stride wider than
 * the row: every VapourSynth stride except super widths that are multiples of the 64-byte alignment).  The C loop
 * (opt=0, and every 16-bit clip) leaves the column at the memset's 0. */
static void average2_last_column_avx2(uint8_t *d, const uint8_t *b, ptrdiff_t pitch, int w, int h) {
    if ((w - 1) % 32 == 0)
        return;
    for (int j = 0; j < h; j++)
        d[(ptrdiff_t)j * pitch + w - 1] = (uint8_t)((b[(ptrdiff_t)j * pitch + w - 1] + 1) >> 1);
}

/* mvpRefine, MVFrame.cpp:1386-1527 */
static void refine_plane(uint8_t *base, ptrdiff_t pitch, const Lv *g, int bps, int bits, int sharp) {
    int pel = g->pel;
    if (pel == 1)
        return;
    ptrdiff_t sub = (ptrdiff_t)pitch * bps * g->ph;
    uint8_t *P[16];
    for (int i = 0; i < pel * pel; i++)
        P[i] = base + sub * i;
    int w = g->pw, h = g->ph;
    int iH = pel == 2 ? 1 : 2, iV = pel == 2 ? 2 : 8, iD = pel == 2 ? 3 : 10;
    refine_horizontal(P[iH], P[0], pitch, w, h, bps, bits, sharp);
    refine_vertical(P[iV], P[0], pitch, w, h, bps, bits, sharp);
    if (sharp == 0)
        refine_diagonal_bilinear(P[iD], P[0], pitch, w, h, bps);
    else
        refine_horizontal(P[iD], P[iV], pitch, w, h, bps, bits, sharp);
    if (pel == 4) {
        ptrdiff_t rowb = pitch * bps;
        average2(P[1], P[0], P[2], pitch, w, h, bps);
        average2(P[9], P[8], P[10], pitch, w, h, bps);
        average2(P[4], P[0], P[8], pitch, w, h, bps);
        average2(P[6], P[2], P[10], pitch, w, h, bps);
        average2(P[5], P[4], P[6], pitch, w, h, bps);
        average2(P[3], P[0] + bps, P[2], pitch, w - 1, h, bps);
        average2(P[11], P[8] + bps, P[10], pitch, w - 1, h, bps);
        average2(P[12], P[0] + rowb, P[8], pitch, w, h - 1, bps);
        average2(P[14], P[2] + rowb, P[10], pitch, w, h - 1, bps);
        average2(P[13], P[12], P[14], pitch, w, h, bps);
        average2(P[7], P[4] + bps, P[6], pitch, w - 1, h, bps);
        average2(P[15], P[12] + bps, P[14], pitch, w - 1, h, bps);
        if (bps == 1) {
            average2_last_column_avx2(P[3], P[2], pitch, w, h);
            average2_last_column_avx2(P[11], P[10], pitch, w, h);
            average2_last_column_avx2(P[7], P[6], pitch, w, h);
            average2_last_column_avx2(P[15], P[14], pitch, w, h);
        }
    }
}

/* mvpRefineExtPel2 / mvpRefineExtPel4, MVFrame.cpp:1530-1614: sub-planes 1..pel^2-1 of level 0 are picked out of a
 * pel-times upsized clip.  Literal restatement, including what the reference does for an already padded pel clip:
 * it still copies only width x height samples, from the pel clip's origin to the planes' padded origin. */
static void refine_ext_plane(uint8_t *level0, ptrdiff_t pitch, const Lv *g, int bps, int pel, const uint8_t *pel_plane, ptrdiff_t pel_stride, int is_padded) {
    ptrdiff_t sub = pitch * g->ph; /* samples between sub-planes */
    ptrdiff_t offset = is_padded ? 0 : (ptrdiff_t)g->vpad * pitch + g->hpad;
    for (int i = 1; i < pel * pel; i++) {
        int sx = i % pel, sy = i / pel;
        for (int y = 0; y < g->h; y++) {
            uint8_t *d = level0 + ((ptrdiff_t)i * sub + offset + (ptrdiff_t)y * pitch) * bps;
            const uint8_t *s = pel_plane + (ptrdiff_t)(y * pel + sy) * pel_stride;
            for (int x = 0; x < g->w; x++) {
                if (bps == 1)
                    d[x] = s[x * pel + sx];
                else
                    ((uint16_t *)d)[x] = ((const uint16_t *)s)[x * pel + sx];
            }
        }
        if (!is_padded)
            pad_plane(level0 + (ptrdiff_t)i * sub * bps, pitch, g->hpad, g->vpad, g->w, g->h, bps);
    }
}

static int super_frame_impl(const mvb200_super_params *p, const uint8_t *const src[3], const ptrdiff_t src_stride[3], const uint8_t *const pel[3], const ptrdiff_t pel_stride[3],
                            int pel_padded, uint8_t *const dst[3], const ptrdiff_t dst_stride[3]);

/* mvsuperGetFrame, MVSuper.c:43-126 */
int mvo_super_frame(const mvb200_super_params *p, const uint8_t *const src[3], const ptrdiff_t src_stride[3], uint8_t *const dst[3], const ptrdiff_t dst_stride[3]) {
    return super_frame_impl(p, src, src_stride, NULL, NULL, 0, dst, dst_stride);
}

/* ... with a pelclip (MVSuper.c:91-104) */
int mvo_super_frame_pel(const mvb200_super_params *p, const uint8_t *const src[3], const ptrdiff_t src_stride[3], const uint8_t *const pel[3], const ptrdiff_t pel_stride[3],
                        int pel_padded, uint8_t *const dst[3], const ptrdiff_t dst_stride[3]) {
    return super_frame_impl(p, src, src_stride, pel, pel_stride, pel_padded, dst, dst_stride);
}

static int super_frame_impl(const mvb200_super_params *p, const uint8_t *const src[3], const ptrdiff_t src_stride[3], const uint8_t *const pel[3], const ptrdiff_t pel_stride[3],
                            int pel_padded, uint8_t *const dst[3], const ptrdiff_t dst_stride[3]) {
    int bps = (p->bits + 7) / 8;
    for (int plane = 0; plane < p->nplanes; plane++) {
        int ph = plane ? p->super_height >> p->ssh : p->super_height;
        memset(dst[plane], 0, (size_t)dst_stride[plane] * ph); /* MVSuper.c:75 */
        if (!(p->modeyuv & (1 << plane)))
            continue;
        ptrdiff_t pitch = dst_stride[plane] / bps;
        Lv g[MVB200_MAX_LEVELS];
        for (int l = 0; l < p->levels; l++)
            level_geom(p->width, p->height, p->xRatioUV, p->yRatioUV, p->hpad, p->vpad, p->pel, plane, l, &g[l]);
#define LEVEL_BASE(l) (dst[plane] + (ptrdiff_t)g[l].row0 * dst_stride[plane])
#define INTERIOR(l) (LEVEL_BASE(l) + ((ptrdiff_t)g[l].vpad * pitch + g[l].hpad) * bps)
        /* mvfFillPlane, MVFrame.cpp:1367-1371 */
        for (int y = 0; y < g[0].h; y++)
            memcpy(INTERIOR(0) + (ptrdiff_t)y * dst_stride[plane], src[plane] + (ptrdiff_t)y * src_stride[plane], (size_t)g[0].w * bps);
        /* mvgofReduce, MVFrame.cpp:1928-1933: reduce then pad each new level */
        for (int l = 0; l < p->levels - 1; l++) {
            reduce_plane(INTERIOR(l + 1), INTERIOR(l), pitch, g[l + 1].w, g[l + 1].h, bps, p->rfilter);
            pad_plane(LEVEL_BASE(l + 1), pitch, g[l + 1].hpad, g[l + 1].vpad, g[l + 1].w, g[l + 1].h, bps);
        }
        /* mvgofPad (level 0), MVSuper.c:89; mvgofRefine, MVSuper.c:103 */
        pad_plane(LEVEL_BASE(0), pitch, g[0].hpad, g[0].vpad, g[0].w, g[0].h, bps);
        if (pel && p->pel >= 2)
            refine_ext_plane(LEVEL_BASE(0), pitch, &g[0], bps, p->pel, pel[plane], pel_stride[plane], pel_padded);
        else
            refine_plane(LEVEL_BASE(0), pitch, &g[0], bps, p->bits, p->sharp);
#undef LEVEL_BASE
#undef INTERIOR
    }
    return 0;
}

/* ======================================================================== */
/* SAD / SATD / luma                                                         */

/* sad_c, SADFunctions.cpp:354-367 */
unsigned mvo_sad(const uint8_t *a, ptrdiff_t ap, const uint8_t *b, ptrdiff_t bp, int w, int h, int bps) {
    unsigned sum = 0;
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++)
            sum += (unsigned)abs(LD(a, bps, x) - LD(b, bps, x));
        a += ap;
        b += bp;
    }
    return sum;
}

static void hadamard4(int *d0, int *d1, int *d2, int *d3, int s0, int s1, int s2, int s3) {
    int t0 = s0 + s1, t1 = s0 - s1, t2 = s2 + s3, t3 = s2 - s3;
    *d0 = t0 + t2;
    *d2 = t0 - t2;
    *d1 = t1 + t3;
    *d3 = t1 - t3;
}

/* sum |H4 . D . H4^T| of one 4x4 tile of differences (un-halved).
 * Real_Satd_4x4_C / Real_Satd_8x4_C, SADFunctions.cpp:605-676, with the packed
 * "two sums in one word" trick (abs2, :596-602) undone into plain ints. */
static unsigned satd_tile4(const uint8_t *a, ptrdiff_t ap, const uint8_t *b, ptrdiff_t bp, int bps) {
    int t[4][4];
    for (int i = 0; i < 4; i++) {
        int d0 = LD(a, bps, 0) - LD(b, bps, 0), d1 = LD(a, bps, 1) - LD(b, bps, 1);
        int d2 = LD(a, bps, 2) - LD(b, bps, 2), d3 = LD(a, bps, 3) - LD(b, bps, 3);
        hadamard4(&t[i][0], &t[i][1], &t[i][2], &t[i][3], d0, d1, d2, d3);
        a += ap;
        b += bp;
    }
    unsigned sum = 0;
    for (int i = 0; i < 4; i++) {
        int a0, a1, a2, a3;
        hadamard4(&a0, &a1, &a2, &a3, t[0][i], t[1][i], t[2][i], t[3][i]);
        sum += (unsigned)(abs(a0) + abs(a1) + abs(a2) + abs(a3));
    }
    return sum;
}

/* Satd_C, SADFunctions.cpp:687-710: 4x4 alone, everything else tiled 8x4, each tile halved */
unsigned mvo_satd(const uint8_t *a, ptrdiff_t ap, const uint8_t *b, ptrdiff_t bp, int w, int h, int bps) {
    if (w == 4 && h == 4)
        return satd_tile4(a, ap, b, bp, bps) >> 1;
    unsigned sum = 0;
    for (int y = 0; y < h; y += 4)
        for (int x = 0; x < w; x += 8) {
            unsigned s = satd_tile4(a + y * ap + x * bps, ap, b + y * bp + x * bps, bp, bps) +
                         satd_tile4(a + y * ap + (x + 4) * bps, ap, b + y * bp + (x + 4) * bps, bp, bps);
            sum += s >> 1;
        }
    return sum;
}

/* luma_c, Luma.cpp:14-25 */
static unsigned luma_sum(const uint8_t *a, ptrdiff_t ap, int w, int h, int bps) {
    unsigned sum = 0;
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++)
            sum += (unsigned)LD(a, bps, x);
        a += ap;
    }
    return sum;
}

/* ======================================================================== */
/* Analyse: the search                                                       */

typedef mvb200_vector VECTOR;

typedef struct PlaneView { /* one plane of one level of a super frame (MVPlane) */
    const uint8_t *base;
    ptrdiff_t pitch; /* bytes */
    Lv g;
} PlaneView;

static const uint8_t *pel_ptr(const PlaneView *v, int bps, int logpel, int x, int y) {
    /* mvpGetAbsolutePointerPel1/2/4, MVFrame.cpp:1706-1730 */
    int idx = 0;
    if (logpel == 1) {
        idx = (x & 1) | ((y & 1) << 1);
        x >>= 1;
        y >>= 1;
    } else if (logpel == 2) {
        idx = (x & 3) | ((y & 3) << 2);
        x >>= 2;
        y >>= 2;
    }
    return v->base + (ptrdiff_t)idx * v->pitch * v->g.ph + (ptrdiff_t)y * v->pitch + (ptrdiff_t)x * bps;
}

/* ---- dct = 1..4: the FFTW DCT path (DCTFFTW.cpp).  The transform itself lives in FFTW3f, which the reference does not
 * vendor or pin (meson.build:185-188): it is restated from FFTW's documented definition of REDFT10 -- unnormalised
 * DCT-II, Y_k = 2 sum_j x_j cos(pi (j + 1/2) k / n), applied to rows then columns -- as plain float dot products in index
 * order.  FFTW's codelets add in a different order, so against a build with the real FFTW this path agrees to float
 * rounding only (PARITY UNPINNED for dct 1..4; the tests pin it against oracle/_ref, whose FFTW is the same definition,
 * oracle/shim/fftw_shim.c).  Everything around the transform is the reference's integer code, restated exactly. */
typedef struct Dct {
    int sizex, sizey, bits, dctshift, dctshift0;
    float *cx, *cy, *fsrc, *ftmp, *fdct; /* cx[k * sizex + j] = (float)(2 cos(pi (2j + 1) k / (2 sizex))) */
    uint8_t *dctSrc, *dctRef;            /* blkW x blkH pixel-typed "DCT images", pitch = blkW * bps */
} Dct;

static float *dct_costab(int n) {
    float *t = (float *)malloc(sizeof(float) * (size_t)n * (size_t)n);
    for (int k = 0; k < n; k++)
        for (int i = 0; i < n; i++)
            t[k * n + i] = (float)(2.0 * cos(M_PI * (2.0 * i + 1.0) * k / (2.0 * n)));
    return t;
}

/* dctInit, DCTFFTW.cpp:135-173 */
static void dct_init(Dct *d, int sizex, int sizey, int bits) {
    d->sizex = sizex;
    d->sizey = sizey;
    d->bits = bits;
    int size2d = sizex * sizey, cur = 1;
    d->dctshift = 0;
    while (cur < size2d) {
        d->dctshift++;
        cur <<= 1;
    }
    d->dctshift0 = d->dctshift + 2;
    d->cx = dct_costab(sizex);
    d->cy = dct_costab(sizey);
    d->fsrc = (float *)malloc(sizeof(float) * (size_t)size2d);
    d->ftmp = (float *)malloc(sizeof(float) * (size_t)size2d);
    d->fdct = (float *)malloc(sizeof(float) * (size_t)size2d);
    int bps = (bits + 7) / 8;
    d->dctSrc = (uint8_t *)malloc((size_t)size2d * bps);
    d->dctRef = (uint8_t *)malloc((size_t)size2d * bps);
}

static void dct_free(Dct *d) {
    free(d->cx);
    free(d->cy);
    free(d->fsrc);
    free(d->ftmp);
    free(d->fdct);
    free(d->dctSrc);
    free(d->dctRef);
}

/* dctBytes2D, DCTFFTW.cpp:200-208: Pixels2Float (:185-197), REDFT10 x REDFT10, Float2Pixels_C (:31-54) */
static void dct_bytes2d(const Dct *d, const uint8_t *src, ptrdiff_t src_pitch, uint8_t *dst) {
    const int nx = d->sizex, ny = d->sizey, bps = (d->bits + 7) / 8;
    for (int j = 0; j < ny; j++)
        for (int i = 0; i < nx; i++)
            d->fsrc[j * nx + i] = (float)LD(src + (ptrdiff_t)j * src_pitch, bps, i);
    for (int i = 0; i < ny; i++) /* rows */
        for (int k = 0; k < nx; k++) {
            float acc = 0.f;
            for (int j = 0; j < nx; j++)
                acc += d->fsrc[i * nx + j] * d->cx[k * nx + j];
            d->ftmp[i * nx + k] = acc;
        }
    for (int k = 0; k < ny; k++) /* columns */
        for (int j = 0; j < nx; j++) {
            float acc = 0.f;
            for (int i = 0; i < ny; i++)
                acc += d->ftmp[i * nx + j] * d->cy[k * ny + i];
            d->fdct[k * nx + j] = acc;
        }
    const int pixelMax = (1 << d->bits) - 1, pixelHalf = 1 << (d->bits - 1);
    const float sqrt_2_div_2 = 0.70710678118654752440084436210485f;
    for (int j = 0; j < ny; j++)
        for (int i = 0; i < nx; i++) {
            float f = d->fdct[j * nx + i] * sqrt_2_div_2;
            int integ = (int)nearbyintf(f);
            int v = IMIN(pixelMax, IMAX(0, (integ >> d->dctshift) + pixelHalf));
            if (bps == 1)
                dst[j * nx + i] = (uint8_t)v;
            else
                ((uint16_t *)dst)[j * nx + i] = (uint16_t)v;
        }
    float f = d->fdct[0] * 0.5f;
    int integ = (int)nearbyintf(f);
    int v = IMIN(pixelMax, IMAX(0, (integ >> d->dctshift0) + pixelHalf)); /* DC */
    if (bps == 1)
        dst[0] = (uint8_t)v;
    else
        ((uint16_t *)dst)[0] = (uint16_t)v;
}

typedef struct Search {
    Dct *dct; /* dct 1..4 */
    /* constant per level */
    int nBlkX, nBlkY, blkW, blkH, ovX, ovY, logPel, logScale, xR, yR, logxR, logyR, bps, chroma, smallest;
    int cw, ch; /* chroma block size */
    PlaneView src[3], ref[3];
    VECTOR *vectors;
    int searchType, searchParam;
    int penaltyZero, pglobal, penaltyNew, badrange, tryMany, dctmode, dctweight16;
    int64_t badSAD, LSAD, verybigSAD;
    VECTOR globalMV, zeroMVfs;
    /* per block */
    int x[3], y[3], blkx, blky, blkIdx, scanDir;
    int dxMin, dxMax, dyMin, dyMax;
    int64_t nLambda, nMinCost;
    VECTOR bestMV, predictor, predictors[5];
    const uint8_t *pSrc[3];
    int srcLuma, badcount;
    int64_t sumLumaChange;
} Search;

static int clipx(const Search *S, int v) { return IMIN(IMAX(v, S->dxMin), S->dxMax - 1); }
static int clipy(const Search *S, int v) { return IMIN(IMAX(v, S->dyMin), S->dyMax - 1); }

static VECTOR clipmv(const Search *S, VECTOR v) { /* pobClipMV, PlaneOfBlocks.cpp:293-311 */
    VECTOR r;
    r.x = clipx(S, v.x);
    r.y = clipy(S, v.y);
    r.sad = v.sad;
    return r;
}

static int median3(int a, int b, int c) { return IMAX(IMIN(a, b), IMIN(IMAX(a, b), c)); }

static const uint8_t *ref_luma(const Search *S, int vx, int vy) { /* pobGetRefBlock :35-53 */
    return pel_ptr(&S->ref[0], S->bps, S->logPel, (S->x[0] << S->logPel) + vx, (S->y[0] << S->logPel) + vy);
}

static const uint8_t *ref_chroma(const Search *S, int plane, int vx, int vy) { /* pobGetRefBlockU/V :56-101 */
    int xbias = (vx < 0) * ((1 << S->logxR) - 1);
    int ybias = (vy < 0) * ((1 << S->logyR) - 1);
    return pel_ptr(&S->ref[plane], S->bps, S->logPel, (S->x[plane] << S->logPel) + ((vx + xbias) >> S->logxR),
                   (S->y[plane] << S->logPel) + ((vy + ybias) >> S->logyR));
}

/* pobLumaSAD<dctmode>, PlaneOfBlocks.cpp:117-203 (all dct modes) */
static int64_t luma_sad(Search *S, const uint8_t *pRef) {
    ptrdiff_t sp = S->src[0].pitch, rp = S->ref[0].pitch;
    int64_t sad = mvo_sad(S->pSrc[0], sp, pRef, rp, S->blkW, S->blkH, S->bps);
    int d = S->dctmode;
    if (d == 0)
        return sad;
    if (d >= 1 && d <= 4) { /* :123-163 */
        const Dct *D = S->dct;
        const ptrdiff_t dp = (ptrdiff_t)S->blkW * S->bps;
        if (d == 1 || (d == 2 && S->dctweight16 > 0)) {
            dct_bytes2d(D, pRef, rp, D->dctRef);
            unsigned dcdiff = (unsigned)abs(LD(D->dctSrc, S->bps, 0) - LD(D->dctRef, S->bps, 0));
            int64_t dctsad = (int64_t)(unsigned)(mvo_sad(D->dctSrc, dp, D->dctRef, dp, S->blkW, S->blkH, S->bps) + dcdiff * 3u) * (int64_t)S->blkW / 2;
            if (d == 1)
                return dctsad;
            return (sad * (16 - S->dctweight16) + dctsad * S->dctweight16) / 16;
        }
        if (d == 2)
            return sad;
        int refLuma = (int)luma_sum(pRef, rp, S->blkW, S->blkH, S->bps);
        if (abs(S->srcLuma - refLuma) > ((S->srcLuma + refLuma) >> 5)) {
            dct_bytes2d(D, pRef, rp, D->dctRef);
            int64_t dctsad = (int64_t)mvo_sad(D->dctSrc, dp, D->dctRef, dp, S->blkW, S->blkH, S->bps) * (int64_t)S->blkW / 2;
            if (d == 3)
                sad = sad / 2 + dctsad / 2;
            else
                sad = sad / 4 + dctsad / 2 + dctsad / 4;
        }
        return sad;
    }
    if (d == 5)
        return mvo_satd(S->pSrc[0], sp, pRef, rp, S->blkW, S->blkH, S->bps);
    if (d == 6) {
        if (S->dctweight16 > 0) {
            int64_t dctsad = mvo_satd(S->pSrc[0], sp, pRef, rp, S->blkW, S->blkH, S->bps);
            sad = (sad * (16 - S->dctweight16) + dctsad * S->dctweight16) / 16;
        }
        return sad;
    }
    if (d == 9) {
        if (S->dctweight16 > 1) {
            int half = S->dctweight16 / 2;
            int64_t dctsad = mvo_satd(S->pSrc[0], sp, pRef, rp, S->blkW, S->blkH, S->bps);
            sad = (sad * (16 - half) + dctsad * half) / 16;
        }
        return sad;
    }
    /* 7, 8, 10: per-block luma-adaptive */
    int refLuma = (int)luma_sum(pRef, rp, S->blkW, S->blkH, S->bps);
    int shift = d == 10 ? 4 : 5;
    if (abs(S->srcLuma - refLuma) > ((S->srcLuma + refLuma) >> shift)) {
        int64_t dctsad = mvo_satd(S->pSrc[0], sp, pRef, rp, S->blkW, S->blkH, S->bps);
        if (d == 7)
            sad = sad / 2 + dctsad / 2;
        else if (d == 8)
            sad = sad / 4 + dctsad / 2 + dctsad / 4;
        else
            sad = sad / 2 + dctsad / 4 + sad / 4;
    }
    return sad;
}

static int64_t chroma_sad(const Search *S, int vx, int vy) {
    int64_t s = 0;
    for (int p = 1; p <= 2; p++)
        s += mvo_sad(S->pSrc[p], S->src[p].pitch, ref_chroma(S, p, vx, vy), S->ref[p].pitch, S->cw, S->ch, S->bps);
    return s;
}

#define F_PENALTYNEW 2
#define F_UPDATEDIR 4
#define F_UPDATEBEST 8

/* pobCheckMV_Template, PlaneOfBlocks.cpp:219-261.  The reference's early returns
 * are short-circuits of "full cost < nMinCost" (every term is >= 0), restated here
 * in the same order so the number of SAD evaluations matches too. */
static void check_mv(Search *S, int vx, int vy, int flags, int *dir, int val) {
    if (!(vx >= S->dxMin && vy >= S->dyMin && vx < S->dxMax && vy < S->dyMax))
        return;
    int dist = (int)((unsigned)((S->predictor.x - vx) * (S->predictor.x - vx) + (S->predictor.y - vy) * (S->predictor.y - vy)));
    int64_t cost = (int)((S->nLambda * dist) >> 8); /* pobMotionDistorsion :111-114 */
    if (cost >= S->nMinCost)
        return;
    int64_t sad = luma_sad(S, ref_luma(S, vx, vy));
    cost += sad + ((flags & F_PENALTYNEW) ? ((S->penaltyNew * sad) >> 8) : 0);
    if (cost >= S->nMinCost)
        return;
    int64_t saduv = 0;
    if (S->chroma) {
        saduv = chroma_sad(S, vx, vy);
        cost += saduv + ((flags & F_PENALTYNEW) ? ((S->penaltyNew * saduv) >> 8) : 0);
        if (cost >= S->nMinCost)
            return;
    }
    if (flags & F_UPDATEBEST) {
        S->bestMV.x = vx;
        S->bestMV.y = vy;
    }
    S->nMinCost = cost;
    S->bestMV.sad = sad + saduv;
    if (flags & F_UPDATEDIR)
        *dir = val;
}

static void CheckMV0(Search *S, int vx, int vy) { check_mv(S, vx, vy, F_UPDATEBEST, 0, 0); }
static void CheckMV(Search *S, int vx, int vy) { check_mv(S, vx, vy, F_PENALTYNEW | F_UPDATEBEST, 0, 0); }
static void CheckMV2(Search *S, int vx, int vy, int *dir, int val) { check_mv(S, vx, vy, F_PENALTYNEW | F_UPDATEDIR | F_UPDATEBEST, dir, val); }
static void CheckMVdir(Search *S, int vx, int vy, int *dir, int val) { check_mv(S, vx, vy, F_PENALTYNEW | F_UPDATEDIR, dir, val); }

/* pobFetchPredictors, PlaneOfBlocks.cpp:419-463 */
static void fetch_predictors(Search *S) {
    int hasNext = (S->scanDir == 1 && S->blkx < S->nBlkX - 1) || (S->scanDir == -1 && S->blkx > 0);
    if ((S->scanDir == 1 && S->blkx > 0) || (S->scanDir == -1 && S->blkx < S->nBlkX - 1))
        S->predictors[1] = clipmv(S, S->vectors[S->blkIdx - S->scanDir]);
    else
        S->predictors[1] = clipmv(S, S->zeroMVfs);
    if (S->blky > 0)
        S->predictors[2] = clipmv(S, S->vectors[S->blkIdx - S->nBlkX]);
    else
        S->predictors[2] = clipmv(S, S->zeroMVfs);
    if (S->blky < S->nBlkY - 1 && hasNext)
        S->predictors[3] = clipmv(S, S->vectors[S->blkIdx + S->nBlkX + S->scanDir]);
    else if (S->blky > 0 && hasNext)
        S->predictors[3] = clipmv(S, S->vectors[S->blkIdx - S->nBlkX + S->scanDir]);
    else
        S->predictors[3] = clipmv(S, S->zeroMVfs);
    if (S->blky > 0) {
        S->predictors[0].x = median3(S->predictors[1].x, S->predictors[2].x, S->predictors[3].x);
        S->predictors[0].y = median3(S->predictors[1].y, S->predictors[2].y, S->predictors[3].y);
        int64_t m = S->predictors[2].sad > S->predictors[3].sad ? S->predictors[2].sad : S->predictors[3].sad;
        S->predictors[0].sad = S->predictors[1].sad > m ? S->predictors[1].sad : m;
    } else {
        S->predictors[0] = S->predictors[1];
    }
    if (S->smallest)
        S->predictor = S->predictors[0];
    double scale = S->LSAD / (double)(S->LSAD + (S->predictor.sad >> 1));
    S->nLambda = (int64_t)(S->nLambda * scale * scale);
}

/* pobNStepSearch :466-485 */
static void nstep_search(Search *S, int stp) {
    for (int length = stp; length > 0; length--) {
        int dx = S->bestMV.x, dy = S->bestMV.y;
        CheckMV(S, dx + length, dy + length);
        CheckMV(S, dx + length, dy);
        CheckMV(S, dx + length, dy - length);
        CheckMV(S, dx, dy - length);
        CheckMV(S, dx, dy + length);
        CheckMV(S, dx - length, dy + length);
        CheckMV(S, dx - length, dy);
        CheckMV(S, dx - length, dy - length);
    }
}

/* pobOneTimeSearch :488-527 */
static void onetime_search(Search *S, int length) {
    int direction = 0;
    int dx = S->bestMV.x, dy = S->bestMV.y;
    CheckMV2(S, dx - length, dy, &direction, 2);
    CheckMV2(S, dx + length, dy, &direction, 1);
    if (direction == 1) {
        while (direction) {
            direction = 0;
            dx += length;
            CheckMV2(S, dx + length, dy, &direction, 1);
        }
    } else if (direction == 2) {
        while (direction) {
            direction = 0;
            dx -= length;
            CheckMV2(S, dx - length, dy, &direction, 1);
        }
    }
    CheckMV2(S, dx, dy - length, &direction, 2);
    CheckMV2(S, dx, dy + length, &direction, 1);
    if (direction == 1) {
        while (direction) {
            direction = 0;
            dy += length;
            CheckMV2(S, dx, dy + length, &direction, 1);
        }
    } else if (direction == 2) {
        while (direction) {
            direction = 0;
            dy -= length;
            CheckMV2(S, dx, dy - length, &direction, 1);
        }
    }
}

/* pobDiamondSearch :530-632 */
static void diamond_search(Search *S, int length) {
    enum { R = 1, L = 2, D = 4, U = 8 };
    int direction = 15, last;
    while (direction > 0) {
        int dx = S->bestMV.x, dy = S->bestMV.y;
        last = direction;
        direction = 0;
        if (last & R) CheckMV2(S, dx + length, dy, &direction, R);
        if (last & L) CheckMV2(S, dx - length, dy, &direction, L);
        if (last & D) CheckMV2(S, dx, dy + length, &direction, D);
        if (last & U) CheckMV2(S, dx, dy - length, &direction, U);
        if (direction) {
            last = direction;
            dx = S->bestMV.x;
            dy = S->bestMV.y;
            if (last & (R + L)) {
                CheckMV2(S, dx, dy + length, &direction, D);
                CheckMV2(S, dx, dy - length, &direction, U);
            } else {
                CheckMV2(S, dx + length, dy, &direction, R);
                CheckMV2(S, dx - length, dy, &direction, L);
            }
        } else {
            switch (last) {
            case R:
                CheckMV2(S, dx + length, dy + length, &direction, R + D);
                CheckMV2(S, dx + length, dy - length, &direction, R + U);
                break;
            case L:
                CheckMV2(S, dx - length, dy + length, &direction, L + D);
                CheckMV2(S, dx - length, dy - length, &direction, L + U);
                break;
            case D:
                CheckMV2(S, dx + length, dy + length, &direction, R + D);
                CheckMV2(S, dx - length, dy + length, &direction, L + D);
                break;
            case U:
                CheckMV2(S, dx + length, dy - length, &direction, R + U);
                CheckMV2(S, dx - length, dy - length, &direction, L + U);
                break;
            case R + D:
                CheckMV2(S, dx + length, dy + length, &direction, R + D);
                CheckMV2(S, dx - length, dy + length, &direction, L + D);
                CheckMV2(S, dx + length, dy - length, &direction, R + U);
                break;
            case L + D:
                CheckMV2(S, dx + length, dy + length, &direction, R + D);
                CheckMV2(S, dx - length, dy + length, &direction, L + D);
                CheckMV2(S, dx - length, dy - length, &direction, L + U);
                break;
            case R + U:
                CheckMV2(S, dx + length, dy + length, &direction, R + D);
                CheckMV2(S, dx - length, dy - length, &direction, L + U);
                CheckMV2(S, dx + length, dy - length, &direction, R + U);
                break;
            case L + U:
                CheckMV2(S, dx - length, dy - length, &direction, L + U);
                CheckMV2(S, dx - length, dy + length, &direction, L + D);
                CheckMV2(S, dx + length, dy - length, &direction, R + U);
                break;
            default:
                CheckMV2(S, dx + length, dy + length, &direction, R + D);
                CheckMV2(S, dx - length, dy + length, &direction, L + D);
                CheckMV2(S, dx + length, dy - length, &direction, R + U);
                CheckMV2(S, dx - length, dy - length, &direction, L + U);
                break;
            }
        }
    }
}

/* pobExpandingSearch :635-658 */
static void expanding_search(Search *S, int r, int s, int mvx, int mvy) {
    for (int i = -r + s; i < r; i += s) {
        CheckMV(S, mvx + i, mvy - r);
        CheckMV(S, mvx + i, mvy + r);
    }
    for (int j = -r + s; j < r; j += s) {
        CheckMV(S, mvx - r, mvy + j);
        CheckMV(S, mvx + r, mvy + j);
    }
    CheckMV(S, mvx - r, mvy - r);
    CheckMV(S, mvx - r, mvy + r);
    CheckMV(S, mvx + r, mvy - r);
    CheckMV(S, mvx + r, mvy + r);
}

/* pobHex2Search :661-724 */
static void hex2_search(Search *S, int range) {
    static const int mod6m1[8] = {5, 0, 1, 2, 3, 4, 5, 0};
    static const int hex2[8][2] = {{-1, -2}, {-2, 0}, {-1, 2}, {1, 2}, {2, 0}, {1, -2}, {-1, -2}, {-2, 0}};
    int dir = -2;
    int bmx = S->bestMV.x, bmy = S->bestMV.y;
    if (range > 1) {
        CheckMVdir(S, bmx - 2, bmy, &dir, 0);
        CheckMVdir(S, bmx - 1, bmy + 2, &dir, 1);
        CheckMVdir(S, bmx + 1, bmy + 2, &dir, 2);
        CheckMVdir(S, bmx + 2, bmy, &dir, 3);
        CheckMVdir(S, bmx + 1, bmy - 2, &dir, 4);
        CheckMVdir(S, bmx - 1, bmy - 2, &dir, 5);
        if (dir != -2) {
            bmx += hex2[dir + 1][0];
            bmy += hex2[dir + 1][1];
            for (int i = 1; i < range / 2 && (bmx >= S->dxMin && bmy >= S->dyMin && bmx < S->dxMax && bmy < S->dyMax); i++) {
                const int odir = mod6m1[dir + 1];
                dir = -2;
                CheckMVdir(S, bmx + hex2[odir + 0][0], bmy + hex2[odir + 0][1], &dir, odir - 1);
                CheckMVdir(S, bmx + hex2[odir + 1][0], bmy + hex2[odir + 1][1], &dir, odir);
                CheckMVdir(S, bmx + hex2[odir + 2][0], bmy + hex2[odir + 2][1], &dir, odir + 1);
                if (dir == -2)
                    break;
                bmx += hex2[dir + 1][0];
                bmy += hex2[dir + 1][1];
            }
        }
        S->bestMV.x = bmx;
        S->bestMV.y = bmy;
    }
    expanding_search(S, 1, 1, bmx, bmy);
}

/* pobCrossSearch + pobUMHSearch :727-769 */
static void umh_search(Search *S, int range, int omx, int omy) {
    for (int i = 1; i < range; i += 2) {
        CheckMV(S, omx - i, omy);
        CheckMV(S, omx + i, omy);
    }
    for (int j = 1; j < range; j += 2) {
        CheckMV(S, omx, omy - j);
        CheckMV(S, omx, omy + j);
    }
    static const int hex4[16][2] = {{-4, 2}, {-4, 1}, {-4, 0}, {-4, -1}, {-4, -2}, {4, -2}, {4, -1}, {4, 0},
                                    {4, 1},  {4, 2},  {2, 3},  {0, 4},   {-2, 3},  {-2, -3}, {0, -4}, {2, -3}};
    int i = 1;
    do {
        for (int j = 0; j < 16; j++)
            CheckMV(S, omx + hex4[j][0] * i, omy + hex4[j][1] * i);
    } while (++i <= range / 4);
    hex2_search(S, range);
}

/* pobRefine :772-816 */
static void refine(Search *S) {
    switch (S->searchType) {
    case 0:
        for (int i = S->searchParam; i > 0; i /= 2)
            onetime_search(S, i);
        break;
    case 1:
        nstep_search(S, S->searchParam);
        break;
    case 2:
        for (int i = S->searchParam; i > 0; i /= 2)
            diamond_search(S, i);
        break;
    case 3: {
        int mvx = S->bestMV.x, mvy = S->bestMV.y;
        for (int i = 1; i <= S->searchParam; i++)
            expanding_search(S, i, 1, mvx, mvy);
        break;
    }
    case 4:
        hex2_search(S, S->searchParam);
        break;
    case 5:
        umh_search(S, S->searchParam, S->bestMV.x, S->bestMV.y);
        break;
    case 6: {
        int mvx = S->bestMV.x, mvy = S->bestMV.y;
        for (int i = 1; i <= S->searchParam; i++) {
            CheckMV(S, mvx - i, mvy);
            CheckMV(S, mvx + i, mvy);
        }
        break;
    }
    case 7: {
        int mvx = S->bestMV.x, mvy = S->bestMV.y;
        for (int i = 1; i <= S->searchParam; i++) {
            CheckMV(S, mvx, mvy - i);
            CheckMV(S, mvx, mvy + i);
        }
        break;
    }
    }
}

static int64_t full_sad_at(Search *S, int lx, int ly, int cx, int cy) {
    int64_t sad = luma_sad(S, ref_luma(S, lx, ly));
    if (S->chroma)
        sad += chroma_sad(S, cx, cy);
    return sad;
}

/* pobPseudoEPZSearch, PlaneOfBlocks.cpp:819-968 */
static void epz_search(Search *S) {
    fetch_predictors(S);
    if (S->dctmode >= 1 && S->dctmode <= 4) /* :824-828 */
        dct_bytes2d(S->dct, S->pSrc[0], S->src[0].pitch, S->dct->dctSrc);
    if (S->dctmode >= 3)
        S->srcLuma = (int)luma_sum(S->pSrc[0], S->src[0].pitch, S->blkW, S->blkH, S->bps);

    /* zero vector: luma at (0, fieldShift), chroma at (0,0) (:834-846) */
    S->bestMV.x = S->zeroMVfs.x;
    S->bestMV.y = S->zeroMVfs.y;
    int64_t sad = full_sad_at(S, 0, S->zeroMVfs.y, 0, 0);
    S->bestMV.sad = sad;
    S->nMinCost = sad + ((S->penaltyZero * sad) >> 8);

    VECTOR bestMany[8];
    int64_t costMany[8] = {0};
    if (S->tryMany) {
        refine(S);
        bestMany[0] = S->bestMV;
        costMany[0] = S->nMinCost;
    }

    /* global predictor (:859-883) */
    S->globalMV = clipmv(S, S->globalMV);
    sad = full_sad_at(S, S->globalMV.x, S->globalMV.y, S->globalMV.x, S->globalMV.y);
    int64_t cost = sad + ((S->pglobal * sad) >> 8);
    if (cost < S->nMinCost || S->tryMany) {
        S->bestMV.x = S->globalMV.x;
        S->bestMV.y = S->globalMV.y;
        S->bestMV.sad = sad;
        S->nMinCost = cost;
    }
    if (S->tryMany) {
        refine(S);
        bestMany[1] = S->bestMV;
        costMany[1] = S->nMinCost;
    }

    /* hierarchical predictor (:884-907) */
    sad = full_sad_at(S, S->predictor.x, S->predictor.y, S->predictor.x, S->predictor.y);
    cost = sad;
    if (cost < S->nMinCost || S->tryMany) {
        S->bestMV.x = S->predictor.x;
        S->bestMV.y = S->predictor.y;
        S->bestMV.sad = sad;
        S->nMinCost = cost;
    }
    if (S->tryMany) {
        refine(S);
        bestMany[2] = S->bestMV;
        costMany[2] = S->nMinCost;
    }

    /* spatial predictors (:910-922) */
    for (int i = 0; i < 4; i++) {
        if (S->tryMany)
            S->nMinCost = S->verybigSAD + 1;
        CheckMV0(S, S->predictors[i].x, S->predictors[i].y);
        if (S->tryMany) {
            refine(S);
            bestMany[i + 3] = S->bestMV;
            costMany[i + 3] = S->nMinCost;
        }
    }

    if (S->tryMany) {
        S->nMinCost = S->verybigSAD + 1;
        for (int i = 0; i < 7; i++)
            if (costMany[i] < S->nMinCost) {
                S->bestMV = bestMany[i];
                S->nMinCost = costMany[i];
            }
    } else {
        refine(S);
    }

    /* bad-block wide search (:938-963) */
    int64_t foundSAD = S->bestMV.sad;
    if (S->blkIdx > 1 && foundSAD > (S->badSAD + S->badSAD * S->badcount / 16)) {
        S->badcount++;
        if (S->badrange > 0) {
            umh_search(S, S->badrange * (1 << S->logPel), 0, 0);
        } else if (S->badrange < 0) {
            for (int i = 1; i < -S->badrange * (1 << S->logPel); i += (1 << S->logPel)) {
                expanding_search(S, i, 1 << S->logPel, 0, 0);
                if (S->bestMV.sad < foundSAD / 4)
                    break;
            }
        }
        int mvx = S->bestMV.x, mvy = S->bestMV.y;
        for (int i = 1; i < (1 << S->logPel); i++)
            expanding_search(S, i, 1, mvx, mvy);
    }
    S->vectors[S->blkIdx] = S->bestMV;
}

/* doPobSearchMVs, PlaneOfBlocks.cpp:971-1131 */
static void search_level(Search *S, int st, int stp, int lambda, int lsad, int pnew, int plevel, VECTOR *out,
                         const VECTOR *globalMVec, int fieldShift, int *pmeanLumaChange, int pzero, int pglobal,
                         int64_t badSAD, int badrange, int meander, int tryMany) {
    S->dctweight16 = IMIN(16, abs(*pmeanLumaChange) / (S->blkW * S->blkH));
    S->badSAD = badSAD;
    S->badrange = badrange;
    S->zeroMVfs.x = 0;
    S->zeroMVfs.y = fieldShift;
    S->zeroMVfs.sad = 0;
    S->globalMV.x = (1 << S->logPel) * globalMVec->x;
    S->globalMV.y = (1 << S->logPel) * globalMVec->y + fieldShift;
    S->globalMV.sad = globalMVec->sad;
    S->searchType = st;
    S->searchParam = stp;
    int nScale = 1 << S->logScale;
    int nLambdaLevel = lambda / ((1 << S->logPel) * (1 << S->logPel));
    if (plevel == 1)
        nLambdaLevel = nLambdaLevel * nScale;
    else if (plevel == 2)
        nLambdaLevel = nLambdaLevel * nScale * nScale;
    S->penaltyZero = pzero;
    S->pglobal = pglobal;
    S->badcount = 0;
    S->tryMany = tryMany;
    S->sumLumaChange = 0;
    S->penaltyNew = pnew;
    S->LSAD = lsad;
    int stepX = S->blkW - S->ovX, stepY = S->blkH - S->ovY;
    const Lv *g0 = &S->src[0].g;
    for (int p = 0; p < 3; p++)
        S->y[p] = S->src[p].base ? S->src[p].g.vpad : 0;
    for (S->blky = 0; S->blky < S->nBlkY; S->blky++) {
        S->scanDir = (S->blky % 2 == 0 || meander == 0) ? 1 : -1;
        int blkxStart = (S->scanDir == 1) ? 0 : S->nBlkX - 1;
        for (int iblkx = 0; iblkx < S->nBlkX; iblkx++) {
            S->blkx = blkxStart + iblkx * S->scanDir;
            S->blkIdx = S->blky * S->nBlkX + S->blkx;
            S->x[0] = g0->hpad + stepX * S->blkx;
            if (S->chroma) {
                /* :1050-1051 start value uses '/', :1116 the increment uses '>>': identical for xRatioUV in {1,2} */
                S->x[1] = S->src[1].g.hpad + (stepX >> S->logxR) * S->blkx;
                S->x[2] = S->src[2].g.hpad + (stepX >> S->logxR) * S->blkx;
            }
            for (int p = 0; p < (S->chroma ? 3 : 1); p++)
                S->pSrc[p] = S->src[p].base + (ptrdiff_t)S->y[p] * S->src[p].pitch + (ptrdiff_t)S->x[p] * S->bps;
            S->nLambda = S->blky == 0 ? 0 : nLambdaLevel;
            int hps = g0->hpad >> S->logScale, vps = g0->vpad >> S->logScale;
            S->dxMax = (g0->pw - S->x[0] - S->blkW - g0->hpad + hps) << S->logPel;
            S->dyMax = (g0->ph - S->y[0] - S->blkH - g0->vpad + vps) << S->logPel;
            S->dxMin = -((S->x[0] - g0->hpad + hps) << S->logPel);
            S->dyMin = -((S->y[0] - g0->vpad + vps) << S->logPel);
            S->predictor = clipmv(S, S->vectors[S->blkIdx]);
            epz_search(S);
            out[S->blkIdx] = S->bestMV;
            if (S->smallest)
                S->sumLumaChange += (int)luma_sum(ref_luma(S, 0, 0), S->ref[0].pitch, S->blkW, S->blkH, S->bps) -
                                    (int)luma_sum(S->pSrc[0], S->src[0].pitch, S->blkW, S->blkH, S->bps);
        }
        S->y[0] += stepY;
        if (S->src[1].base) S->y[1] += stepY >> S->logyR;
        if (S->src[2].base) S->y[2] += stepY >> S->logyR;
    }
    if (S->smallest)
        *pmeanLumaChange = (int)((int)S->sumLumaChange / (S->nBlkX * S->nBlkY));
}

/* pobEstimateGlobalMVDoubled, PlaneOfBlocks.cpp:1559-1636 */
static void estimate_global_mv(const VECTOR *v, int n, int pel, VECTOR *g) {
    int freqSize = 8192 * pel * 2;
    int *freq = (int *)malloc(sizeof(int) * (size_t)freqSize);
    int med[2];
    for (int c = 0; c < 2; c++) {
        memset(freq, 0, sizeof(int) * (size_t)freqSize);
        int indmin = freqSize - 1, indmax = 0;
        for (int i = 0; i < n; i++) {
            int ind = (freqSize >> 1) + (c ? v[i].y : v[i].x);
            if (ind >= 0 && ind < freqSize) {
                freq[ind]++;
                if (ind > indmax) indmax = ind;
                if (ind < indmin) indmin = ind;
            }
        }
        int count = freq[indmin], index = indmin;
        for (int i = indmin + 1; i <= indmax; i++)
            if (freq[i] > count) {
                count = freq[i];
                index = i;
            }
        med[c] = index - (freqSize >> 1);
    }
    free(freq);
    int meanvx = 0, meanvy = 0, num = 0;
    for (int i = 0; i < n; i++)
        if (abs(v[i].x - med[0]) < 6 && abs(v[i].y - med[1]) < 6) {
            meanvx += v[i].x;
            meanvy += v[i].y;
            num++;
        }
    if (num > 0) {
        g->x = 2 * meanvx / num;
        g->y = 2 * meanvy / num;
    } else {
        g->x = 2 * med[0];
        g->y = 2 * med[1];
    }
}

/* pobInterpolatePrediction, PlaneOfBlocks.cpp:1447-1514 */
static void interpolate_prediction(VECTOR *dst, int nBlkX, int nBlkY, int logPel, const VECTOR *v2, int nBlkX2, int nBlkY2,
                                   int logPel2, int blkW, int blkH, int ovX, int ovY) {
    int normFactor = 3 - logPel + logPel2;
    int mulFactor = normFactor < 0 ? -normFactor : 0;
    normFactor = normFactor < 0 ? 0 : normFactor;
    int normov = (blkW - ovX) * (blkH - ovY);
    int aoddx = blkW * 3 - ovX * 2, aevenx = blkW * 3 - ovX * 4;
    int aoddy = blkH * 3 - ovY * 2, aeveny = blkH * 3 - ovY * 4;
    double scaleov = 1.0 / normov;
    for (int l = 0, index = 0; l < nBlkY; l++)
        for (int k = 0; k < nBlkX; k++, index++) {
            VECTOR v1, w2, v3, v4;
            int i = k, j = l;
            if (i >= 2 * nBlkX2) i = 2 * nBlkX2 - 1;
            if (j >= 2 * nBlkY2) j = 2 * nBlkY2 - 1;
            int offy = -1 + 2 * (j % 2), offx = -1 + 2 * (i % 2);
            int edgex = (i == 0) || (i >= 2 * nBlkX2 - 1), edgey = (j == 0) || (j >= 2 * nBlkY2 - 1);
            if (edgex) {
                if (edgey) {
                    v1 = w2 = v3 = v4 = v2[i / 2 + (j / 2) * nBlkX2];
                } else {
                    v1 = w2 = v2[i / 2 + (j / 2) * nBlkX2];
                    v3 = v4 = v2[i / 2 + (j / 2 + offy) * nBlkX2];
                }
            } else if (edgey) {
                v1 = w2 = v2[i / 2 + (j / 2) * nBlkX2];
                v3 = v4 = v2[i / 2 + offx + (j / 2) * nBlkX2];
            } else {
                v1 = v2[i / 2 + (j / 2) * nBlkX2];
                w2 = v2[i / 2 + offx + (j / 2) * nBlkX2];
                v3 = v2[i / 2 + (j / 2 + offy) * nBlkX2];
                v4 = v2[i / 2 + offx + (j / 2 + offy) * nBlkX2];
            }
            int64_t temp_sad;
            int vx, vy;
            if (ovX == 0 && ovY == 0) {
                vx = 9 * v1.x + 3 * w2.x + 3 * v3.x + v4.x;
                vy = 9 * v1.y + 3 * w2.y + 3 * v3.y + v4.y;
                temp_sad = 9 * v1.sad + 3 * w2.sad + 3 * v3.sad + v4.sad + 8;
            } else { /* overlap <= blk/2 always (MVAnalyse.c:441-445) */
                int ax1 = offx > 0 ? aoddx : aevenx;
                int ax2 = (blkW - ovX) * 4 - ax1;
                int ay1 = offy > 0 ? aoddy : aeveny;
                int ay2 = (blkH - ovY) * 4 - ay1;
                int64_t a11 = ax1 * ay1, a12 = ax1 * ay2, a21 = ax2 * ay1, a22 = ax2 * ay2;
                vx = (int)((a11 * v1.x + a21 * w2.x + a12 * v3.x + a22 * v4.x) * scaleov);
                vy = (int)((a11 * v1.y + a21 * w2.y + a12 * v3.y + a22 * v4.y) * scaleov);
                temp_sad = (int64_t)((a11 * v1.sad + a21 * w2.sad + a12 * v3.sad + a22 * v4.sad) * scaleov);
            }
            dst[index].x = (vx >> normFactor) * (1 << mulFactor);
            dst[index].y = (vy >> normFactor) * (1 << mulFactor);
            dst[index].sad = temp_sad >> 4;
        }
}

static void bind_planes(PlaneView v[3], const uint8_t *const planes[3], const ptrdiff_t stride[3], const mvb200_super_params *sp, int level, int nplanes_used) {
    for (int p = 0; p < 3; p++) {
        v[p].base = NULL;
        if (p < nplanes_used && planes[p]) {
            level_geom(sp->width, sp->height, sp->xRatioUV, sp->yRatioUV, sp->hpad, sp->vpad, sp->pel, p, level, &v[p].g);
            v[p].pitch = stride[p];
            v[p].base = planes[p] + (ptrdiff_t)v[p].g.row0 * stride[p];
        }
    }
}

/* GetMedian, GroupOfPlanes.c:179-203 */
static int med3(int a, int b, int c) {
    if (((b <= a) && (a <= c)) || ((c <= a) && (a <= b))) return a;
    if (((a <= b) && (b <= c)) || ((c <= b) && (b <= a))) return b;
    return c;
}

static void get_median(int32_t *vx, int32_t *vy, int x1, int y1, int x2, int y2, int x3, int y3) {
    *vx = med3(x1, x2, x3);
    *vy = med3(y1, y2, y3);
    if ((*vx == x1 && *vy == y1) || (*vx == x2 && *vy == y2) || (*vx == x3 && *vy == y3))
        return;
    *vx = x1;
    *vy = y1;
}

/* gopExtraDivide, GroupOfPlanes.c:206-302 */
static void extra_divide(const VECTOR *in, VECTOR *out, int nBlkX, int nBlkY, int divide) {
    for (int by = 0; by < nBlkY; by++)
        for (int bx = 0; bx < nBlkX; bx++) {
            VECTOR b = in[by * nBlkX + bx];
            b.sad >>= 2;
            VECTOR *o = out + (by * 2) * (nBlkX * 2) + bx * 2;
            o[0] = o[1] = o[nBlkX * 2] = o[nBlkX * 2 + 1] = b;
            if (divide > 1 && by >= 1 && by < nBlkY - 1 && bx >= 1 && bx < nBlkX - 1) {
                const VECTOR *c = in + by * nBlkX + bx;
                get_median(&o[0].x, &o[0].y, c[0].x, c[0].y, c[-1].x, c[-1].y, c[-nBlkX].x, c[-nBlkX].y);
                get_median(&o[1].x, &o[1].y, c[0].x, c[0].y, c[1].x, c[1].y, c[-nBlkX].x, c[-nBlkX].y);
                get_median(&o[nBlkX * 2].x, &o[nBlkX * 2].y, c[0].x, c[0].y, c[-1].x, c[-1].y, c[nBlkX].x, c[nBlkX].y);
                get_median(&o[nBlkX * 2 + 1].x, &o[nBlkX * 2 + 1].y, c[0].x, c[0].y, c[1].x, c[1].y, c[nBlkX].x, c[nBlkX].y);
            }
        }
}

/* mvanalyseGetFrame (MVAnalyse.c:76-254) + gopSearchMVs (GroupOfPlanes.c:69-125) */
int mvo_analyse_frame(const mvb200_analyse_params *ap, const mvb200_super_params *sp, const uint8_t *const src_super[3], const ptrdiff_t src_stride[3], const uint8_t *const ref_super[3], const ptrdiff_t ref_stride[3], int field_shift, uint8_t *vectors_out) {
    const mvb200_analysis_data *ad = &ap->ad;
    int L = ad->nLvCount;
    int32_t *hdr = (int32_t *)vectors_out;
    int bps = (ad->bitsPerSample + 7) / 8;
    int64_t verybig = (int64_t)ad->nBlkSizeX * ad->nBlkSizeY * (1 << ad->bitsPerSample);
    hdr[0] = ap->vectors_size;
    uint8_t *out = vectors_out + 8;

    if (!ref_super) { /* gopWriteDefaultToArray, GroupOfPlanes.c:150-164 */
        hdr[1] = 0;
        for (int l = L - 1; l >= 0; l--) {
            int n = ap->nBlkX[l] * ap->nBlkY[l];
            *(int32_t *)out = 4 + n * 16;
            VECTOR *v = (VECTOR *)(out + 4);
            for (int i = 0; i < n; i++) {
                v[i].x = 0;
                v[i].y = 0;
                v[i].sad = verybig;
            }
            out += 4 + n * 16;
            if (l == 0 && ap->divide) {
                *(int32_t *)out = 4 + n * 16 * 4;
                v = (VECTOR *)(out + 4);
                for (int i = 0; i < n * 4; i++) {
                    v[i].x = 0;
                    v[i].y = 0;
                    v[i].sad = verybig;
                }
            }
        }
        return 0;
    }
    Dct dctc;
    const int use_dct = ap->dctmode >= 1 && ap->dctmode <= 4; /* MVAnalyse.c:196-201 */
    if (use_dct)
        dct_init(&dctc, ap->ad.nBlkSizeX, ap->ad.nBlkSizeY, sp->bits);

    hdr[1] = 1;
    VECTOR *lvVec[MVB200_MAX_LEVELS];
    VECTOR *lvOut[MVB200_MAX_LEVELS];
    {
        uint8_t *o = out;
        for (int l = L - 1; l >= 0; l--) {
            int n = ap->nBlkX[l] * ap->nBlkY[l];
            *(int32_t *)o = 4 + n * 16; /* pobWriteHeaderToArray */
            lvOut[l] = (VECTOR *)(o + 4);
            lvVec[l] = (VECTOR *)calloc((size_t)n, sizeof(VECTOR)); /* pobInit zeroes vectors :354-355 */
            o += 4 + n * 16;
        }
    }
    int nplanes_used = ap->chroma ? 3 : 1;
    VECTOR globalMV = {0, 0, -1};
    int pglobal = ap->global ? ap->pglobal : ap->pzero;
    int meanLumaChange = 0;

    for (int l = L - 1; l >= 0; l--) {
        Search S;
        memset(&S, 0, sizeof(S));
        S.nBlkX = ap->nBlkX[l];
        S.nBlkY = ap->nBlkY[l];
        S.blkW = ad->nBlkSizeX;
        S.blkH = ad->nBlkSizeY;
        S.ovX = ad->nOverlapX;
        S.ovY = ad->nOverlapY;
        S.logPel = l == 0 ? (ad->nPel == 4 ? 2 : ad->nPel == 2 ? 1 : 0) : 0;
        S.logScale = l;
        S.xR = ad->xRatioUV;
        S.yR = ad->yRatioUV;
        S.logxR = S.xR == 2;
        S.logyR = S.yR == 2;
        S.bps = bps;
        S.chroma = ap->chroma;
        S.cw = S.blkW / S.xR;
        S.ch = S.blkH / S.yR;
        S.smallest = (l == L - 1);
        S.dctmode = ap->dctmode;
        S.dct = use_dct ? &dctc : NULL;
        S.verybigSAD = verybig;
        S.vectors = lvVec[l];
        bind_planes(S.src, src_super, src_stride, sp, l, nplanes_used);
        bind_planes(S.ref, ref_super, ref_stride, sp, l, nplanes_used);
        int st, stp, fs, tm;
        if (l == L - 1) {
            st = (L == 1 || ap->search == 6 || ap->search == 7) ? ap->search : ap->search_coarse;
            stp = (L == 1) ? ap->nPelSearch : ap->nSearchParam;
            fs = (L - 1 == 0) ? field_shift : 0;
            tm = ap->tryMany && L > 1;
        } else {
            st = (l == 0 || ap->search == 6 || ap->search == 7) ? ap->search : ap->search_coarse;
            stp = (l == 0) ? ap->nPelSearch : ap->nSearchParam;
            if (ap->global)
                estimate_global_mv(lvVec[l + 1], ap->nBlkX[l + 1] * ap->nBlkY[l + 1], 1, &globalMV);
            interpolate_prediction(lvVec[l], S.nBlkX, S.nBlkY, S.logPel, lvVec[l + 1], ap->nBlkX[l + 1], ap->nBlkY[l + 1], 0,
                                   S.blkW, S.blkH, S.ovX, S.ovY);
            fs = (l == 0) ? field_shift : 0;
            tm = ap->tryMany && l > 0;
        }
        search_level(&S, st, stp, ap->nLambda, ap->lsad, ap->pnew, ap->plevel, lvOut[l], &globalMV, fs, &meanLumaChange,
                     ap->pzero, pglobal, ap->badSAD, ap->badrange, ap->meander, tm);
    }
    if (ap->divide) {
        int n = ap->nBlkX[0] * ap->nBlkY[0];
        uint8_t *o = (uint8_t *)(lvOut[0] + n);
        *(int32_t *)o = 4 + n * 16 * 4; /* the reference leaves these 4 bytes uninitialised */
        extra_divide(lvOut[0], (VECTOR *)(o + 4), ap->nBlkX[0], ap->nBlkY[0], ap->divide);
    }
    for (int l = 0; l < L; l++)
        free(lvVec[l]);
    if (use_dct)
        dct_free(&dctc);
    return 0;
}

/* ======================================================================== */
/* Degrain                                                                   */

/* overInit, Overlap.cpp:40-125 */
void mvo_overlap_windows(int nx, int ny, int ox, int oy, int16_t *out9) {
    float *wx = (float *)malloc(sizeof(float) * 3 * (size_t)nx);
    float *wy = (float *)malloc(sizeof(float) * 3 * (size_t)ny);
    float *wxm = wx, *wxf = wx + nx, *wxl = wx + 2 * nx;
    float *wym = wy, *wyf = wy + ny, *wyl = wy + 2 * ny;
    for (int i = 0; i < ox; i++) {
        wxm[i] = cosf(M_PI * (i - ox + 0.5f) / (ox * 2));
        wxm[i] = wxm[i] * wxm[i];
        wxf[i] = 1;
        wxl[i] = wxm[i];
    }
    for (int i = ox; i < nx - ox; i++)
        wxm[i] = wxf[i] = wxl[i] = 1;
    for (int i = nx - ox; i < nx; i++) {
        wxm[i] = cosf(M_PI * (i - nx + ox + 0.5f) / (ox * 2));
        wxm[i] = wxm[i] * wxm[i];
        wxf[i] = wxm[i];
        wxl[i] = 1;
    }
    for (int i = 0; i < oy; i++) {
        wym[i] = cosf(M_PI * (i - oy + 0.5f) / (oy * 2));
        wym[i] = wym[i] * wym[i];
        wyf[i] = 1;
        wyl[i] = wym[i];
    }
    for (int i = oy; i < ny - oy; i++)
        wym[i] = wyf[i] = wyl[i] = 1;
    for (int i = ny - oy; i < ny; i++) {
        wym[i] = cosf(M_PI * (i - ny + oy + 0.5f) / (oy * 2));
        wym[i] = wym[i] * wym[i];
        wyf[i] = wym[i];
        wyl[i] = 1;
    }
    const float *ys[3] = {wyf, wym, wyl};
    const float *xs[3] = {wxf, wxm, wxl};
    for (int wy3 = 0; wy3 < 3; wy3++)
        for (int wx3 = 0; wx3 < 3; wx3++) {
            int16_t *w = out9 + (size_t)(wy3 * 3 + wx3) * nx * ny;
            for (int j = 0; j < ny; j++)
                for (int i = 0; i < nx; i++)
                    w[j * nx + i] = (int16_t)(int)(ys[wy3][j] * xs[wx3][i] * 2048 + 0.5f);
        }
    free(wx);
    free(wy);
}

/* DegrainWeight, MVDegrains.h:184-189 */
static int degrain_weight(int64_t thSAD, int64_t blockSAD) {
    if (blockSAD >= thSAD)
        return 0;
    return (int)((thSAD - blockSAD) * (thSAD + blockSAD) * 256 / (double)(thSAD * thSAD + blockSAD * blockSAD));
}

/* level-0 block vectors of a blob (fgopUpdate, Fakery.c:110-121: level 0 is the last plain level) */
static const VECTOR *blob_level0(const uint8_t *blob, const mvb200_analysis_data *vd, int *validity) {
    const int32_t *h = (const int32_t *)blob;
    *validity = (h[1] == 1);
    const uint8_t *p = blob + 8;
    for (int i = vd->nLvCount - 1; i >= 1; i--)
        p += *(const int32_t *)p;
    return (const VECTOR *)(p + 4);
}

/* mvdegrainGetFrame<radius>, MVDegrains.cpp:85-330 */
int mvo_degrain_frame(const mvb200_degrain_params *dp, const mvb200_super_params *sp, const uint8_t *const src[3], const ptrdiff_t src_stride[3], const uint8_t *const *ref_super, const ptrdiff_t *ref_stride, const uint8_t *const *vectors, uint8_t *const dst[3], const ptrdiff_t dst_stride[3]) {
    int R2 = dp->radius * 2;
    int bps = (dp->bits + 7) / 8;
    int nBlkX = dp->vd[0].nBlkX, nBlkY = dp->vd[0].nBlkY;
    int nLogPel = dp->vd[0].nPel == 4 ? 2 : dp->vd[0].nPel == 2 ? 1 : 0;
    int pixelMax = (1 << dp->bits) - 1;
    const VECTOR *vec[12];
    int usable[12];
    for (int r = 0; r < R2; r++) {
        int valid;
        vec[r] = blob_level0(vectors[r], &dp->vd[r], &valid);
        /* fpobIsSceneChange, Fakery.c:52-58 */
        int sum = 0;
        for (int i = 0; i < nBlkX * nBlkY; i++)
            sum += vec[r][i].sad > dp->nSCD1 ? 1 : 0;
        usable[r] = valid && !(sum > dp->nSCD2);
        if (usable[r] && !ref_super[r * 3])
            return fail("oracle degrain: usable vectors but no reference super frame");
    }
    int stepX0 = dp->nBlkSizeX[0] - dp->nOverlapX[0], stepY0 = dp->nBlkSizeY[0] - dp->nOverlapY[0];
    int overlap = dp->nOverlapX[0] > 0 || dp->nOverlapY[0] > 0;

    for (int plane = 0; plane < dp->nplanes; plane++) {
        int W = dp->nWidth[plane], H = dp->nHeight[plane];
        if (!dp->process[plane]) {
            for (int y = 0; y < H; y++)
                memcpy(dst[plane] + (ptrdiff_t)y * dst_stride[plane], src[plane] + (ptrdiff_t)y * src_stride[plane], (size_t)W * bps);
            continue;
        }
        int bw = dp->nBlkSizeX[plane], bh = dp->nBlkSizeY[plane];
        int ox = dp->nOverlapX[plane], oy = dp->nOverlapY[plane];
        int xSub = plane ? dp->ssw : 0, ySub = plane ? dp->ssh : 0;
        PlaneView rv[12];
        for (int r = 0; r < R2; r++)
            if (usable[r]) {
                level_geom(sp->width, sp->height, sp->xRatioUV, sp->yRatioUV, sp->hpad, sp->vpad, sp->pel, plane, 0, &rv[r].g);
                rv[r].pitch = ref_stride[r * 3 + plane];
                rv[r].base = ref_super[r * 3 + plane];
            }
        int16_t *win = NULL;
        uint32_t *acc = NULL;
        int WB = dp->nWidth_B[plane], HB = dp->nHeight_B[plane];
        if (overlap) {
            win = (int16_t *)malloc(sizeof(int16_t) * 9 * (size_t)bw * bh);
            mvo_overlap_windows(bw, bh, ox, oy, win);
            acc = (uint32_t *)calloc((size_t)W * H + 1, sizeof(uint32_t));
        }
        int *tmp = (int *)malloc(sizeof(int) * (size_t)bw * bh);
        for (int by = 0; by < nBlkY; by++)
            for (int bx = 0; bx < nBlkX; bx++) {
                int i = by * nBlkX + bx;
                int px = overlap ? bx * (bw - ox) : bx * bw; /* block origin in this plane */
                int py = overlap ? by * (bh - oy) : by * bh;
                const uint8_t *ptr[12];
                ptrdiff_t pit[12];
                int WRef[12], WSrc;
                for (int r = 0; r < R2; r++) { /* useBlock, MVDegrains.h:192-206 */
                    if (usable[r]) {
                        int blx = ((bx * stepX0) << nLogPel) + vec[r][i].x;
                        int bly = ((by * stepY0) << nLogPel) + vec[r][i].y;
                        int ax = (plane ? blx >> xSub : blx) + rv[r].g.hpad * (1 << nLogPel);
                        int ay = (plane ? bly >> ySub : bly) + rv[r].g.vpad * (1 << nLogPel);
                        ptr[r] = pel_ptr(&rv[r], bps, nLogPel, ax, ay);
                        pit[r] = rv[r].pitch;
                        WRef[r] = degrain_weight(dp->thSAD[plane], vec[r][i].sad);
                    } else {
                        ptr[r] = src[plane] + (ptrdiff_t)py * src_stride[plane] + (ptrdiff_t)px * bps;
                        pit[r] = src_stride[plane];
                        WRef[r] = 0;
                    }
                }
                /* normaliseWeights, MVDegrains.h:209-223 */
                WSrc = 256;
                int WSum = WSrc + 1;
                for (int r = 0; r < R2; r++)
                    WSum += WRef[r];
                double scale = 256.0 / WSum;
                for (int r = 0; r < R2; r++) {
                    WRef[r] = (int)(WRef[r] * scale);
                    WSrc -= WRef[r];
                }
                /* Degrain_C, MVDegrains.h:31-53 */
                const uint8_t *ps = src[plane] + (ptrdiff_t)py * src_stride[plane] + (ptrdiff_t)px * bps;
                for (int y = 0; y < bh; y++)
                    for (int x = 0; x < bw; x++) {
                        int sum = 128 + LD(ps + (ptrdiff_t)y * src_stride[plane], bps, x) * WSrc;
                        for (int r = 0; r < R2; r++)
                            sum += LD(ptr[r] + (ptrdiff_t)y * pit[r], bps, x) * WRef[r];
                        tmp[y * bw + x] = bps == 1 ? (uint8_t)(sum >> 8) : (uint16_t)(sum >> 8);
                    }
                if (!overlap) {
                    for (int y = 0; y < bh; y++)
                        for (int x = 0; x < bw; x++)
                            ST(dst[plane] + (ptrdiff_t)(py + y) * dst_stride[plane], bps, px + x, tmp[y * bw + x]);
                } else {
                    /* window choice MVDegrains.cpp:256-262, overlaps_c Overlap.cpp:144-158 */
                    int wby = ((by + nBlkY - 3) / (nBlkY - 2)) * 3;
                    int wbx = bx == 0 ? 0 : (bx == nBlkX - 1 ? 2 : 1);
                    if (nBlkX == 1) wbx = 2;
                    const int16_t *w = win + (size_t)(wby + wbx) * bw * bh;
                    for (int y = 0; y < bh; y++)
                        for (int x = 0; x < bw; x++) {
                            uint32_t add = (uint32_t)((tmp[y * bw + x] * w[y * bw + x]) >> 6);
                            uint32_t *a = &acc[(size_t)(py + y) * W + px + x];
                            if (bps == 1)
                                *a = (uint16_t)(*a + add); /* uint16 accumulator wraps */
                            else
                                *a += add;
                        }
                }
            }
        if (overlap) { /* ToPixels, Overlap.cpp:335-356 */
            for (int y = 0; y < HB; y++)
                for (int x = 0; x < WB; x++) {
                    int a = (int)((acc[(size_t)y * W + x] + 16) >> 5);
                    ST(dst[plane] + (ptrdiff_t)y * dst_stride[plane], bps, x, IMIN(pixelMax, a));
                }
        }
        /* uncovered right / bottom strips copied from the source (:238-249, :290-298) */
        if (dp->nWidth_B[0] < dp->nWidth[0])
            for (int y = 0; y < HB; y++)
                memcpy(dst[plane] + (ptrdiff_t)y * dst_stride[plane] + (ptrdiff_t)WB * bps, src[plane] + (ptrdiff_t)y * src_stride[plane] + (ptrdiff_t)WB * bps, (size_t)(W - WB) * bps);
        if (dp->nHeight_B[0] < dp->nHeight[0])
            for (int y = HB; y < H; y++)
                memcpy(dst[plane] + (ptrdiff_t)y * dst_stride[plane], src[plane] + (ptrdiff_t)y * src_stride[plane], (size_t)W * bps);
        /* LimitChanges_C, MVDegrains.h:168-181 */
        if (dp->nLimit[plane] < pixelMax)
            for (int y = 0; y < H; y++)
                for (int x = 0; x < W; x++) {
                    int s = LD(src[plane] + (ptrdiff_t)y * src_stride[plane], bps, x);
                    int d = LD(dst[plane] + (ptrdiff_t)y * dst_stride[plane], bps, x);
                    ST(dst[plane] + (ptrdiff_t)y * dst_stride[plane], bps, x, IMIN(IMAX(d, s - dp->nLimit[plane]), s + dp->nLimit[plane]));
                }
        free(tmp);
        free(win);
        free(acc);
    }
    return 0;
}

/* ======================================================================== */
/* Recalculate                                                               */

/* doPobRecalculateMVs (PlaneOfBlocks.cpp:1158-1424) + gopRecalculateMVs (GroupOfPlanes.c:128-147) +
 * mvrecalculateGetFrame (MVRecalculate.c:68-254).  ref_super == NULL or an invalid old blob -> defaults. */
int mvo_recalculate_frame(const mvb200_recalculate_params *rp, const mvb200_super_params *sp, const uint8_t *const src_super[3], const ptrdiff_t src_stride[3], const uint8_t *const ref_super[3], const ptrdiff_t ref_stride[3], const uint8_t *old_vectors, int field_shift, uint8_t *vectors_out) {
    const mvb200_analysis_data *ad = &rp->ad;
    const mvb200_analysis_data *od = &rp->old;
    int32_t *hdr = (int32_t *)vectors_out;
    int bps = (ad->bitsPerSample + 7) / 8;
    int n = ad->nBlkX * ad->nBlkY;
    int64_t verybig = (int64_t)ad->nBlkSizeX * ad->nBlkSizeY * (1 << ad->bitsPerSample);
    hdr[0] = rp->vectors_size;
    uint8_t *out = vectors_out + 8;
    *(int32_t *)out = 4 + n * 16;
    VECTOR *res = (VECTOR *)(out + 4);
    int old_valid;
    const VECTOR *oldv = blob_level0(old_vectors, od, &old_valid);
    if (!ref_super || !old_valid) { /* gopWriteDefaultToArray */
        hdr[1] = 0;
        for (int i = 0; i < n; i++) {
            res[i].x = 0;
            res[i].y = 0;
            res[i].sad = verybig;
        }
        if (rp->divide) {
            uint8_t *o = (uint8_t *)(res + n);
            *(int32_t *)o = 4 + n * 16 * 4;
            VECTOR *v = (VECTOR *)(o + 4);
            for (int i = 0; i < n * 4; i++) {
                v[i].x = 0;
                v[i].y = 0;
                v[i].sad = verybig;
            }
        }
        return 0;
    }
    Dct dctc;
    const int use_dct = rp->dctmode >= 1 && rp->dctmode <= 4; /* MVRecalculate.c:193-198 */
    if (use_dct)
        dct_init(&dctc, rp->ad.nBlkSizeX, rp->ad.nBlkSizeY, sp->bits);
    hdr[1] = 1;
    Search S;
    memset(&S, 0, sizeof(S));
    S.dct = use_dct ? &dctc : NULL;
    S.nBlkX = ad->nBlkX;
    S.nBlkY = ad->nBlkY;
    S.blkW = ad->nBlkSizeX;
    S.blkH = ad->nBlkSizeY;
    S.ovX = ad->nOverlapX;
    S.ovY = ad->nOverlapY;
    S.logPel = ad->nPel == 4 ? 2 : ad->nPel == 2 ? 1 : 0;
    S.xR = ad->xRatioUV;
    S.yR = ad->yRatioUV;
    S.logxR = S.xR == 2;
    S.logyR = S.yR == 2;
    S.bps = bps;
    S.chroma = rp->chroma;
    S.cw = S.blkW / S.xR;
    S.ch = S.blkH / S.yR;
    S.dctmode = rp->dctmode;
    S.dctweight16 = 8; /* :1166 */
    S.verybigSAD = verybig;
    S.vectors = res;
    S.searchType = rp->search;
    S.searchParam = rp->nSearchParam;
    S.penaltyNew = rp->pnew;
    bind_planes(S.src, src_super, src_stride, sp, 0, rp->chroma ? 3 : 1);
    bind_planes(S.ref, ref_super, ref_stride, sp, 0, rp->chroma ? 3 : 1);
    int nLambdaLevel = rp->nLambda / ((1 << S.logPel) * (1 << S.logPel));
    int nBlkXold = od->nBlkX, nBlkYold = od->nBlkY;
    int nStepXold = od->nBlkSizeX - od->nOverlapX, nStepYold = od->nBlkSizeY - od->nOverlapY;
    int nLogPelold = od->nPel == 4 ? 2 : od->nPel == 2 ? 1 : 0;
    int stepX = S.blkW - S.ovX, stepY = S.blkH - S.ovY;
    const Lv *g0 = &S.src[0].g;
    (void)field_shift; /* only zeroMVfieldShifted / globalMVPredictor use it, neither is read by the refinement */
    for (S.blky = 0; S.blky < S.nBlkY; S.blky++)
        for (S.blkx = 0; S.blkx < S.nBlkX; S.blkx++) { /* blocks are independent: scan order does not matter */
            S.blkIdx = S.blky * S.nBlkX + S.blkx;
            S.x[0] = g0->hpad + stepX * S.blkx;
            S.y[0] = g0->vpad + stepY * S.blky;
            if (S.chroma) {
                S.x[1] = S.x[2] = S.src[1].g.hpad + (stepX >> S.logxR) * S.blkx;
                S.y[1] = S.y[2] = S.src[1].g.vpad + (stepY >> S.logyR) * S.blky;
            }
            for (int p = 0; p < (S.chroma ? 3 : 1); p++)
                S.pSrc[p] = S.src[p].base + (ptrdiff_t)S.y[p] * S.src[p].pitch + (ptrdiff_t)S.x[p] * S.bps;
            S.nLambda = S.blky == 0 ? 0 : nLambdaLevel;
            /* :1274-1277 -- the whole padded plane */
            S.dxMax = (g0->pw - S.x[0] - S.blkW) << S.logPel;
            S.dyMax = (g0->ph - S.y[0] - S.blkH) << S.logPel;
            S.dxMin = -(S.x[0] << S.logPel);
            S.dyMin = -(S.y[0] << S.logPel);
            /* :1279-1330 -- predictor from the old vector field */
            int centerX = S.blkW / 2 + stepX * S.blkx;
            int blkxold = (centerX - od->nBlkSizeX / 2) / nStepXold;
            int centerY = S.blkH / 2 + stepY * S.blky;
            int blkyold = (centerY - od->nBlkSizeY / 2) / nStepYold;
            int deltaX = IMAX(0, centerX - (od->nBlkSizeX / 2 + nStepXold * blkxold));
            int deltaY = IMAX(0, centerY - (od->nBlkSizeY / 2 + nStepYold * blkyold));
            int bx1 = IMIN(nBlkXold - 1, IMAX(0, blkxold)), bx2 = IMIN(nBlkXold - 1, IMAX(0, blkxold + 1));
            int by1 = IMIN(nBlkYold - 1, IMAX(0, blkyold)), by2 = IMIN(nBlkYold - 1, IMAX(0, blkyold + 1));
            VECTOR vo;
            if (rp->smooth == 1) {
                VECTOR v1 = oldv[bx1 + by1 * nBlkXold], v2 = oldv[bx2 + by1 * nBlkXold], v3 = oldv[bx1 + by2 * nBlkXold], v4 = oldv[bx2 + by2 * nBlkXold];
                int a_x = v1.x * nStepXold + deltaX * (v2.x - v1.x);
                int a_y = v1.y * nStepXold + deltaX * (v2.y - v1.y);
                int64_t a_s = v1.sad * nStepXold + deltaX * (v2.sad - v1.sad);
                int b_x = v3.x * nStepXold + deltaX * (v4.x - v3.x);
                int b_y = v3.y * nStepXold + deltaX * (v4.y - v3.y);
                int64_t b_s = v3.sad * nStepXold + deltaX * (v4.sad - v3.sad);
                vo.x = (a_x + deltaY * (b_x - a_x) / nStepYold) / nStepXold;
                vo.y = (a_y + deltaY * (b_y - a_y) / nStepYold) / nStepXold;
                vo.sad = (a_s + deltaY * (b_s - a_s) / nStepYold) / nStepXold;
            } else {
                int rx = deltaX * 2 >= nStepXold, ry = deltaY * 2 >= nStepYold;
                vo = oldv[(rx ? bx2 : bx1) + (ry ? by2 : by1) * nBlkXold];
            }
            vo.x = (vo.x << S.logPel) >> nLogPelold;
            vo.y = (vo.y << S.logPel) >> nLogPelold;
            S.predictor = clipmv(&S, vo);
            S.predictor.sad = vo.sad * (S.blkW * S.blkH) / (od->nBlkSizeX * od->nBlkSizeY);
            S.bestMV = S.predictor;
            if (use_dct) /* :1336-1338 */
                dct_bytes2d(&dctc, S.pSrc[0], S.src[0].pitch, dctc.dctSrc);
            if (S.dctmode >= 3)
                S.srcLuma = (int)luma_sum(S.pSrc[0], S.src[0].pitch, S.blkW, S.blkH, S.bps);
            int64_t sad = full_sad_at(&S, S.predictor.x, S.predictor.y, S.predictor.x, S.predictor.y);
            S.bestMV.sad = sad;
            S.nMinCost = sad;
            if (S.bestMV.sad > rp->thSAD)
                refine(&S);
            res[S.blkIdx] = S.bestMV;
        }
    if (rp->divide) {
        uint8_t *o = (uint8_t *)(res + n);
        *(int32_t *)o = 4 + n * 16 * 4;
        extra_divide(res, (VECTOR *)(o + 4), ad->nBlkX, ad->nBlkY, rp->divide);
    }
    if (use_dct)
        dct_free(&dctc);
    return 0;
}

/* ======================================================================== */
/* Compensate                                                                */

/* mvcompensateGetFrame, MVCompensate.c:73-374.  src_super = super frame n, ref_super = super frame nref
 * (NULL when out of range).  Only the planes of the super clip's modeyuv are written (the reference leaves
 * the others uninitialised). */
int mvo_compensate_frame(const mvb200_compensate_params *cp, const mvb200_super_params *sp, const uint8_t *const src_super[3], const ptrdiff_t src_stride[3], const uint8_t *const ref_super[3], const ptrdiff_t ref_stride[3], const uint8_t *vectors, int field_shift, uint8_t *const dst[3], const ptrdiff_t dst_stride[3]) {
    const mvb200_analysis_data *vd = &cp->vd;
    int bps = (cp->bits + 7) / 8;
    int nBlkX = vd->nBlkX, nBlkY = vd->nBlkY;
    int nPel = vd->nPel, nLogPel = nPel == 4 ? 2 : nPel == 2 ? 1 : 0;
    int pixelMax = (1 << cp->bits) - 1;
    int num_planes = (cp->super_modeyuv & (MVB200_UPLANE | MVB200_VPLANE)) ? 3 : 1;
    int valid;
    const VECTOR *vec = blob_level0(vectors, vd, &valid);
    int sum = 0;
    for (int i = 0; i < nBlkX * nBlkY; i++)
        sum += vec[i].sad > cp->nSCD1 ? 1 : 0;
    int usable = valid && !(sum > cp->nSCD2);
    if (usable && !ref_super)
        return fail("oracle compensate: usable vectors but no reference super frame");
    if (!usable) { /* :348-364 */
        const uint8_t *const *s = (!cp->scBehavior && ref_super) ? ref_super : src_super;
        const ptrdiff_t *ss = (!cp->scBehavior && ref_super) ? ref_stride : src_stride;
        for (int plane = 0; plane < num_planes; plane++) {
            Lv g;
            level_geom(sp->width, sp->height, sp->xRatioUV, sp->yRatioUV, sp->hpad, sp->vpad, sp->pel, plane, 0, &g);
            for (int y = 0; y < cp->nHeight[plane]; y++)
                memcpy(dst[plane] + (ptrdiff_t)y * dst_stride[plane], s[plane] + (ptrdiff_t)(y + g.vpad) * ss[plane] + (ptrdiff_t)g.hpad * bps, (size_t)cp->nWidth[plane] * bps);
        }
        return 0;
    }
    int overlap = cp->nOverlapX[0] > 0 || cp->nOverlapY[0] > 0;
    int stepX0 = cp->nBlkSizeX[0] - cp->nOverlapX[0], stepY0 = cp->nBlkSizeY[0] - cp->nOverlapY[0];
    for (int plane = 0; plane < num_planes; plane++) {
        int W = cp->nWidth[plane], H = cp->nHeight[plane], WB = cp->nWidth_B[plane], HB = cp->nHeight_B[plane];
        int bw = cp->nBlkSizeX[plane], bh = cp->nBlkSizeY[plane], ox = cp->nOverlapX[plane], oy = cp->nOverlapY[plane];
        int xSub = plane ? cp->ssw : 0, ySub = plane ? cp->ssh : 0;
        PlaneView rv, sv;
        level_geom(sp->width, sp->height, sp->xRatioUV, sp->yRatioUV, sp->hpad, sp->vpad, sp->pel, plane, 0, &rv.g);
        sv.g = rv.g;
        rv.pitch = ref_stride[plane];
        rv.base = ref_super[plane];
        sv.pitch = src_stride[plane];
        sv.base = src_super[plane];
        int16_t *win = NULL;
        uint32_t *acc = NULL;
        if (overlap) {
            win = (int16_t *)malloc(sizeof(int16_t) * 9 * (size_t)bw * bh);
            mvo_overlap_windows(bw, bh, ox, oy, win);
            acc = (uint32_t *)calloc((size_t)W * H + 1, sizeof(uint32_t));
        }
        for (int by = 0; by < nBlkY; by++)
            for (int bx = 0; bx < nBlkX; bx++) {
                int i = by * nBlkX + bx;
                int blx, bly;
                const PlaneView *pv;
                if (vec[i].sad < cp->thSAD) { /* :238-242 / :286-290 */
                    blx = (bx * stepX0) * nPel + vec[i].x * cp->time256 / 256;
                    bly = (by * stepY0) * nPel + vec[i].y * cp->time256 / 256 + field_shift;
                    pv = &rv;
                } else {
                    blx = (bx * stepX0) * nPel;
                    bly = (by * stepY0) * nPel + field_shift;
                    pv = &sv;
                }
                int ax = (plane ? blx >> xSub : blx) + pv->g.hpad * nPel;
                int ay = (plane ? bly >> ySub : bly) + pv->g.vpad * nPel;
                const uint8_t *p = pel_ptr(pv, bps, nLogPel, ax, ay);
                int px = overlap ? bx * (bw - ox) : bx * bw, py = overlap ? by * (bh - oy) : by * bh;
                if (!overlap) {
                    for (int y = 0; y < bh; y++)
                        memcpy(dst[plane] + (ptrdiff_t)(py + y) * dst_stride[plane] + (ptrdiff_t)px * bps, p + (ptrdiff_t)y * pv->pitch, (size_t)bw * bps);
                } else {
                    int wby = ((by + nBlkY - 3) / (nBlkY - 2)) * 3;
                    int wbx = bx == 0 ? 0 : (bx == nBlkX - 1 ? 2 : 1);
                    if (nBlkX == 1) wbx = 2;
                    const int16_t *w = win + (size_t)(wby + wbx) * bw * bh;
                    for (int y = 0; y < bh; y++)
                        for (int x = 0; x < bw; x++) {
                            uint32_t add = (uint32_t)((LD(p + (ptrdiff_t)y * pv->pitch, bps, x) * w[y * bw + x]) >> 6);
                            uint32_t *a = &acc[(size_t)(py + y) * W + px + x];
                            if (bps == 1)
                                *a = (uint16_t)(*a + add);
                            else
                                *a += add;
                        }
                }
            }
        if (overlap) {
            for (int y = 0; y < HB; y++)
                for (int x = 0; x < WB; x++) {
                    int a = (int)((acc[(size_t)y * W + x] + 16) >> 5);
                    ST(dst[plane] + (ptrdiff_t)y * dst_stride[plane], bps, x, IMIN(pixelMax, a));
                }
        }
        /* uncovered strips from the source (scbehavior=1) or reference super frame's interior (:330-342) */
        const uint8_t *sc = cp->scBehavior ? src_super[plane] : ref_super[plane];
        ptrdiff_t scs = cp->scBehavior ? src_stride[plane] : ref_stride[plane];
        if (cp->nWidth_B[0] < cp->nWidth[0])
            for (int y = 0; y < HB; y++)
                memcpy(dst[plane] + (ptrdiff_t)y * dst_stride[plane] + (ptrdiff_t)WB * bps, sc + (ptrdiff_t)(y + rv.g.vpad) * scs + (ptrdiff_t)(WB + rv.g.hpad) * bps, (size_t)(W - WB) * bps);
        if (cp->nHeight_B[0] < cp->nHeight[0])
            for (int y = HB; y < H; y++)
                memcpy(dst[plane] + (ptrdiff_t)y * dst_stride[plane], sc + (ptrdiff_t)(y + rv.g.vpad) * scs + (ptrdiff_t)rv.g.hpad * bps, (size_t)W * bps);
        free(win);
        free(acc);
    }
    return 0;
}
