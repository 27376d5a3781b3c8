/* TEST INFRASTRUCTURE -- CPU oracle, second part: the consumers of the vector clips that SURVEY.md
 * section 8f ranks next (mv.SCDetection, mv.Mask, mv.Finest, mv.Flow).  A plain-C restatement of the
 * reference's algorithms, every function citing the reference lines it follows
 * (dubhater/vapoursynth-mvtools @ e516e90, paths under src/).  PINNED: tests/test_oracle_flow.py checks it
 * against goldens produced by the compiled reference (oracle/_ref/libmvtools_ref.so) and, where that
 * library is present, against the reference live.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this; the product
 * (libmvtools_b200.so, the plugin) never does. */
#include "mvtools_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct VEC {
    int32_t x, y;
    int64_t sad;
} VEC;

static VEC get_vec(const uint8_t *level0, int i) { /* VECTOR arrays are only 4-byte aligned inside the blob */
    VEC v;
    memcpy(&v, level0 + (size_t)i * 16, 16);
    return v;
}

/* fgopUpdate, Fakery.c:110-121: levels are stored coarsest first, each with a 4-byte size header */
static const uint8_t *blob_level0(const uint8_t *blob, const mvb200_analysis_data *vd, int *validity) {
    int32_t v;
    memcpy(&v, blob + 4, 4);
    *validity = (v == 1); /* fgopGetValidity, Fakery.c:103-107 */
    const uint8_t *p = blob + 8;
    for (int i = vd->nLvCount - 1; i >= 1; i--) {
        int32_t size;
        memcpy(&size, p, 4);
        p += size;
    }
    return p + 4;
}

/* fgopIsUsable, Fakery.c:144-146 with fpobIsSceneChange, Fakery.c:52-58 */
static int is_usable(const uint8_t *blob, const mvb200_analysis_data *vd, int64_t th1, int th2) {
    int validity;
    const uint8_t *l0 = blob_level0(blob, vd, &validity);
    int sum = 0;
    for (int i = 0; i < vd->nBlkX * vd->nBlkY; i++)
        sum += get_vec(l0, i).sad > th1 ? 1 : 0;
    return !(sum > th2) && validity;
}

/* mvscdetectionGetFrame, MVSCDetection.c:43-74: the value of _SceneChangePrev / _SceneChangeNext */
int mvo_scdetection_frame(const mvb200_scdetection_params *p, const uint8_t *vectors, int32_t *scene_change) {
    *scene_change = !is_usable(vectors, &p->vd, p->nSCD1, p->nSCD2);
    return 0;
}

/* ------------------------------------------------------------ SimpleResize */

enum { WSHIFT = 14, WMAX = 1 << WSHIFT, WHALF = WMAX / 2 }; /* SimpleResize.h:31-35 */

/* InitTables, SimpleResize.cpp:26-56 */
void mvo_simple_resize_tables(int out, int in, int32_t *offsets, int32_t *weights) {
    float leftmost = 0.5f, rightmost = in - 0.5f;
    int leftmost_idx = (int)leftmost > 0 ? (int)leftmost : 0;
    int rightmost_idx = (int)rightmost < in - 1 ? (int)rightmost : in - 1;
    for (int i = 0; i < out; i++) {
        float position = (i + 0.5f) * (float)in / (float)out;
        float weight;
        int offset;
        if (position <= leftmost) {
            offset = leftmost_idx;
            weight = 0.0f;
        } else if (position >= rightmost) {
            offset = rightmost_idx - 1;
            weight = 1.0f;
        } else {
            offset = (int)(position - leftmost);
            weight = position - leftmost - offset;
        }
        offsets[i] = offset;
        weights[i] = (int)(weight * WMAX);
    }
}

typedef struct Resizer { /* simpleInit, SimpleResize.cpp:119-157 */
    int dst_w, dst_h, src_w, src_h, limit_w, limit_h, pel;
    int32_t *voff, *vwt, *hoff, *hwt;
} Resizer;

static void resizer_init(Resizer *r, int dst_w, int dst_h, int src_w, int src_h, int limit_w, int limit_h, int pel) {
    r->dst_w = dst_w;
    r->dst_h = dst_h;
    r->src_w = src_w;
    r->src_h = src_h;
    r->limit_w = limit_w;
    r->limit_h = limit_h;
    r->pel = pel;
    r->voff = malloc(sizeof(int32_t) * (size_t)dst_h);
    r->vwt = malloc(sizeof(int32_t) * (size_t)dst_h);
    r->hoff = malloc(sizeof(int32_t) * (size_t)dst_w);
    r->hwt = malloc(sizeof(int32_t) * (size_t)dst_w);
    mvo_simple_resize_tables(dst_w, src_w, r->hoff, r->hwt);
    mvo_simple_resize_tables(dst_h, src_h, r->voff, r->vwt);
}

static void resizer_free(Resizer *r) {
    free(r->voff);
    free(r->vwt);
    free(r->hoff);
    free(r->hwt);
}

/* simpleResize<uint8_t>, SimpleResize.cpp:60-116 (no limiting for 8-bit data) */
static void resize_u8(const Resizer *r, uint8_t *dst, ptrdiff_t dst_stride, const uint8_t *src, int src_stride) {
    uint8_t *work = malloc((size_t)r->src_w);
    for (int y = 0; y < r->dst_h; y++) {
        int wb = r->vwt[y], wt = WMAX - wb;
        const uint8_t *s1 = src + r->voff[y] * src_stride, *s2 = s1 + src_stride;
        for (int x = 0; x < r->src_w; x++)
            work[x] = (uint8_t)((s1[x] * wt + s2[x] * wb + WHALF) >> WSHIFT);
        for (int x = 0; x < r->dst_w; x++) {
            int wr = r->hwt[x], wl = WMAX - wr, o = r->hoff[x];
            dst[x] = (uint8_t)((work[o] * wl + work[o + 1] * wr + WHALF) >> WSHIFT);
        }
        dst += dst_stride;
    }
    free(work);
}

/* simpleResize<int16_t>, SimpleResize.cpp:60-116: vectors are limited so that the position they point at stays
 * inside the frame (the running minimum / maximum of :66-70,83-86,97-102,108-111) */
static void resize_i16(const Resizer *r, int16_t *dst, int dst_stride, const int16_t *src, int src_stride, int horizontal_vectors) {
    int pel = r->pel;
    int minimum = 0, maximum = r->limit_h * pel - 1;
    int hstep = horizontal_vectors ? pel : 0, vstep = horizontal_vectors ? 0 : pel;
    int16_t *work = malloc(sizeof(int16_t) * (size_t)r->src_w);
    for (int y = 0; y < r->dst_h; y++) {
        int wb = r->vwt[y], wt = WMAX - wb;
        const int16_t *s1 = src + r->voff[y] * src_stride, *s2 = s1 + src_stride;
        for (int x = 0; x < r->src_w; x++)
            work[x] = (int16_t)((s1[x] * wt + s2[x] * wb + WHALF) >> WSHIFT);
        if (horizontal_vectors) {
            minimum = 0;
            maximum = r->limit_w * pel - 1;
        }
        for (int x = 0; x < r->dst_w; x++) {
            int wr = r->hwt[x], wl = WMAX - wr, o = r->hoff[x];
            int result = (work[o] * wl + work[o + 1] * wr + WHALF) >> WSHIFT;
            result = result < maximum ? result : maximum; /* std::max(minimum, std::min(result, maximum)) */
            result = result > minimum ? result : minimum;
            minimum -= hstep;
            maximum -= hstep;
            dst[x] = (int16_t)result;
        }
        dst += dst_stride;
        minimum -= vstep;
        maximum -= vstep;
    }
    free(work);
}

/* ------------------------------------------------------------------- Mask */

/* mvmaskLength, MVMask.c:65-71 */
static uint8_t mask_length(VEC v, int pel, float fMaskNormFactor2, float fHalfGamma) {
    double norme = (double)(v.x * v.x + v.y * v.y) / (pel * pel);
    double l = 255 * pow(norme * fMaskNormFactor2, fHalfGamma);
    return (uint8_t)((l > 255) ? 255 : l);
}

/* ByteNorm, MaskFun.cpp:129-133 */
static uint8_t byte_norm(int64_t sad, double dSADNormFactor, double fGamma) {
    double l = 255 * pow(sad * dSADNormFactor, fGamma);
    return (uint8_t)((l > 255) ? 255 : l);
}

/* MakeSADMaskTime, MaskFun.cpp:136-160 */
static void sad_mask_time(const uint8_t *l0, int nBlkX, int nBlkY, double dSADNormFactor, double fGamma, int nPel, uint8_t *mask, int pitch, int time256, int stepX, int stepY, int bits) {
    memset(mask, 0, (size_t)nBlkY * pitch);
    int time4096X = (256 - time256) * 16 / (stepX * nPel);
    int time4096Y = (256 - time256) * 16 / (stepY * nPel);
    for (int by = 0; by < nBlkY; by++)
        for (int bx = 0; bx < nBlkX; bx++) {
            VEC v = get_vec(l0, bx + by * nBlkX);
            int bxi = bx - v.x * time4096X / 4096;
            int byi = by - v.y * time4096Y / 4096;
            if (bxi < 0 || bxi >= nBlkX || byi < 0 || byi >= nBlkY) {
                bxi = bx;
                byi = by;
            }
            int64_t sad = get_vec(l0, bxi + byi * nBlkX).sad >> (bits - 8);
            mask[bx + by * pitch] = byte_norm(sad, dSADNormFactor, fGamma);
        }
}

/* ByteOccMask, MaskFun.cpp:85-90 */
static void byte_occ_mask(uint8_t *m, int occlusion, double occnorm, double fGamma) {
    int v;
    if (fGamma == 1.0)
        v = (int)(255 * occlusion * occnorm);
    else
        v = (int)(255 * pow(occlusion * occnorm, fGamma));
    v = v < 255 ? v : 255;
    *m = *m > v ? *m : (uint8_t)v;
}

/* MakeVectorOcclusionMaskTime, MaskFun.cpp:92-126 */
static void occlusion_mask_time(const uint8_t *l0, int isBackward, int nBlkX, int nBlkY, double dMaskNormDivider, double fGamma, int nPel, uint8_t *occ, int pitch, int time256, int stepX, int stepY) {
    memset(occ, 0, (size_t)pitch * nBlkY);
    int time4096X = time256 * 16 / (stepX * nPel);
    int time4096Y = time256 * 16 / (stepY * nPel);
    double occnormX = 80.0 / (dMaskNormDivider * stepX * nPel);
    double occnormY = 80.0 / (dMaskNormDivider * stepY * nPel);
    for (int by = 0; by < nBlkY; by++)
        for (int bx = 0; bx < nBlkX; bx++) {
            int i = bx + by * nBlkX;
            VEC v = get_vec(l0, i);
            if (bx < nBlkX - 1) {
                int vx1 = get_vec(l0, i + 1).x;
                if (vx1 < v.x) {
                    int occlusion = v.x - vx1;
                    int lo = bx + 1 - occlusion * time4096X / 4096;
                    int minb = isBackward ? (lo > 0 ? lo : 0) : bx;
                    int maxb = isBackward ? bx + 1 : (lo < nBlkX - 1 ? lo : nBlkX - 1);
                    for (int bxi = minb; bxi <= maxb; bxi++)
                        byte_occ_mask(&occ[bxi + by * pitch], occlusion, occnormX, fGamma);
                }
            }
            if (by < nBlkY - 1) {
                int vy1 = get_vec(l0, i + nBlkX).y;
                if (vy1 < v.y) {
                    int occlusion = v.y - vy1;
                    int lo = by + 1 - occlusion * time4096Y / 4096;
                    int minb = isBackward ? (lo > 0 ? lo : 0) : by;
                    int maxb = isBackward ? by + 1 : (lo < nBlkY - 1 ? lo : nBlkY - 1);
                    for (int byi = minb; byi <= maxb; byi++)
                        byte_occ_mask(&occ[bx + byi * pitch], occlusion, occnormY, fGamma);
                }
            }
        }
}

static int clamp255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

/* mvmaskGetFrame, MVMask.c:74-203 */
int mvo_mask_frame(const mvb200_mask_params *p, const uint8_t *src_luma, ptrdiff_t src_stride, const uint8_t *vectors, uint8_t *const dst[3], const ptrdiff_t dst_stride[3]) {
    const mvb200_analysis_data *vd = &p->vd;
    const int kind = p->kind, nWidth = p->nWidth, nHeight = p->nHeight, nWidthUV = p->nWidthUV, nHeightUV = p->nHeightUV;
    if (is_usable(vectors, vd, p->nSCD1, p->nSCD2)) {
        int validity;
        const uint8_t *l0 = blob_level0(vectors, vd, &validity);
        const int nBlkX = vd->nBlkX, nBlkY = vd->nBlkY, n = nBlkX * nBlkY;
        uint8_t *small = malloc((size_t)n), *smallV = malloc((size_t)n);
        const int stepX = vd->nBlkSizeX - vd->nOverlapX, stepY = vd->nBlkSizeY - vd->nOverlapY;
        if (kind == 0) {
            for (int j = 0; j < n; j++)
                small[j] = mask_length(get_vec(l0, j), vd->nPel, p->fMaskNormFactor2, p->fHalfGamma);
        } else if (kind == 1) {
            sad_mask_time(l0, nBlkX, nBlkY, 4.0 * p->fMaskNormFactor / (vd->nBlkSizeX * vd->nBlkSizeY), p->fGamma, vd->nPel, small, nBlkX, p->time256, stepX, stepY, 8 /* the clip is 8-bit, MVMask.c:318 */);
        } else if (kind == 2) {
            occlusion_mask_time(l0, vd->isBackward, nBlkX, nBlkY, 1.0 / p->fMaskNormFactor, p->fGamma, vd->nPel, small, nBlkX, p->time256, stepX, stepY);
        } else if (kind == 3) {
            for (int j = 0; j < n; j++)
                small[j] = (uint8_t)clamp255((int)(get_vec(l0, j).x * p->fMaskNormFactor * 100 + 128));
        } else if (kind == 4) {
            for (int j = 0; j < n; j++)
                small[j] = (uint8_t)clamp255((int)(get_vec(l0, j).y * p->fMaskNormFactor * 100 + 128));
        } else {
            for (int j = 0; j < n; j++) {
                VEC v = get_vec(l0, j);
                small[j] = (uint8_t)clamp255((int)(v.x * p->fMaskNormFactor * 100 + 128));
                smallV[j] = (uint8_t)clamp255((int)(v.y * p->fMaskNormFactor * 100 + 128));
            }
        }
        Resizer up, upUV;
        resizer_init(&up, p->nWidthB, p->nHeightB, nBlkX, nBlkY, nWidth, nHeight, vd->nPel);
        resizer_init(&upUV, p->nWidthBUV, p->nHeightBUV, nBlkX, nBlkY, nWidthUV, nHeightUV, vd->nPel);
        if (kind == 5) {
            for (int h = 0; h < nHeight; h++)
                memcpy(dst[0] + h * dst_stride[0], src_luma + h * src_stride, (size_t)nWidth);
        } else {
            resize_u8(&up, dst[0], dst_stride[0], small, nBlkX);
            for (int h = 0; h < nHeight && nWidth > p->nWidthB; h++) /* :158-161 (rows below nHeightB are rewritten next) */
                for (int w = p->nWidthB; w < nWidth; w++)
                    dst[0][h * dst_stride[0] + w] = dst[0][h * dst_stride[0] + p->nWidthB - 1];
            for (int h = p->nHeightB; h < nHeight; h++) /* :162-163 */
                memcpy(dst[0] + h * dst_stride[0], dst[0] + (p->nHeightB - 1) * dst_stride[0], (size_t)nWidth);
        }
        resize_u8(&upUV, dst[1], dst_stride[1], small, nBlkX);
        if (kind == 5)
            resize_u8(&upUV, dst[2], dst_stride[2], smallV, nBlkX);
        else
            for (int h = 0; h < p->nHeightBUV; h++)
                memcpy(dst[2] + h * dst_stride[2], dst[1] + h * dst_stride[1], (size_t)p->nWidthBUV);
        for (int pl = 1; pl < 3; pl++) { /* :175-185 */
            for (int h = 0; h < nHeightUV && nWidthUV > p->nWidthBUV; h++)
                for (int w = p->nWidthBUV; w < nWidthUV; w++)
                    dst[pl][h * dst_stride[pl] + w] = dst[pl][h * dst_stride[pl] + p->nWidthBUV - 1];
            for (int h = p->nHeightBUV; h < nHeightUV; h++)
                memcpy(dst[pl] + h * dst_stride[pl], dst[pl] + (p->nHeightBUV - 1) * dst_stride[pl], (size_t)nWidthUV);
        }
        resizer_free(&up);
        resizer_free(&upUV);
        free(small);
        free(smallV);
    } else { /* :189-196 */
        for (int h = 0; h < nHeight; h++) {
            if (kind == 5)
                memcpy(dst[0] + h * dst_stride[0], src_luma + h * src_stride, (size_t)nWidth);
            else
                memset(dst[0] + h * dst_stride[0], p->nSceneChangeValue, (size_t)nWidth);
        }
        for (int pl = 1; pl < 3; pl++)
            for (int h = 0; h < nHeightUV; h++)
                memset(dst[pl] + h * dst_stride[pl], p->nSceneChangeValue, (size_t)nWidthUV);
    }
    return 0;
}

/* ----------------------------------------------------------------- Finest */

typedef struct FinestPlane {
    int pw, ph; /* padded size of level 0 of this plane */
} FinestPlane;

static void finest_plane_dims(const mvb200_super_params *sp, int plane, FinestPlane *f) {
    int xr = plane ? sp->xRatioUV : 1, yr = plane ? sp->yRatioUV : 1;
    f->pw = sp->width / xr + 2 * (sp->hpad / xr); /* mvpInit via mvfInit, MVFrame.cpp:1775-1779: chroma padding = pad / ratio */
    f->ph = sp->height / yr + 2 * (sp->vpad / yr);
}

/* mvfinestGetFrame, MVFinest.c:47-125 with Merge4PlanesToBig / Merge16PlanesToBig, MaskFun.cpp:196-336:
 * dst(y * pel + sy, x * pel + sx) = sub-plane (sx, sy) at (x, y); sub-plane (sx, sy) starts
 * (sx | sy << log2 pel) * padded_height rows into the plane (mvpGetAbsolutePointer, MVFrame.h / SURVEY.md a1).
 * pel 1: the top (W + 2 hpad) x (H + 2 vpad) rows, i.e. level 0 (:75-77; the reference copies the LUMA size for
 * every plane there, which overruns subsampled chroma planes -- the oracle copies each plane's own size). */
int mvo_finest_frame(const mvb200_super_params *sp, const uint8_t *const super[3], const ptrdiff_t super_stride[3], uint8_t *const dst[3], const ptrdiff_t dst_stride[3]) {
    const int bps = (sp->bits + 7) / 8, pel = sp->pel, lp = pel == 4 ? 2 : (pel == 2 ? 1 : 0);
    for (int pl = 0; pl < sp->nplanes; pl++) {
        if (pl && !(sp->modeyuv & (MVB200_UPLANE | MVB200_VPLANE)))
            continue;
        FinestPlane f;
        finest_plane_dims(sp, pl, &f);
        for (int y = 0; y < f.ph * pel; y++)
            for (int x = 0; x < f.pw * pel; x++) {
                int idx = (x & (pel - 1)) | ((y & (pel - 1)) << lp);
                const uint8_t *s = super[pl] + ((ptrdiff_t)idx * f.ph + (y >> lp)) * super_stride[pl] + (ptrdiff_t)(x >> lp) * bps;
                memcpy(dst[pl] + (ptrdiff_t)y * dst_stride[pl] + (ptrdiff_t)x * bps, s, (size_t)bps);
            }
    }
    return 0;
}

/* ------------------------------------------------------------------- Flow */

/* MakeVectorSmallMasks (MaskFun.cpp:163-176) + CheckAndPadSmallY (:38-62) + the field shift of MVFlow.cpp:301-305 */
static void small_vectors(const uint8_t *l0, int nBlkX, int nBlkY, int nBlkXP, int nBlkYP, int fieldShift, int16_t *vx, int16_t *vy) {
    for (int by = 0; by < nBlkY; by++)
        for (int bx = 0; bx < nBlkX; bx++) {
            VEC v = get_vec(l0, bx + by * nBlkX);
            vx[bx + by * nBlkXP] = (int16_t)v.x;
            vy[bx + by * nBlkXP] = (int16_t)v.y;
        }
    if (nBlkXP > nBlkX)
        for (int j = 0; j < nBlkY; j++) {
            int16_t r = vx[j * nBlkXP + nBlkX - 1];
            int16_t vxr = r < 0 ? r : 0, vyr = vy[j * nBlkXP + nBlkX - 1];
            for (int dx = nBlkX; dx < nBlkXP; dx++) {
                vx[j * nBlkXP + dx] = vxr;
                vy[j * nBlkXP + dx] = vyr;
            }
        }
    if (nBlkYP > nBlkY)
        for (int i = 0; i < nBlkXP; i++) {
            int16_t vxb = vx[nBlkXP * (nBlkY - 1) + i];
            int16_t b = vy[nBlkXP * (nBlkY - 1) + i];
            int16_t vyb = b < 0 ? b : 0;
            for (int dy = nBlkY; dy < nBlkYP; dy++) {
                vx[nBlkXP * dy + i] = vxb;
                vy[nBlkXP * dy + i] = vyb;
            }
        }
    for (int i = 0; i < nBlkXP * nBlkYP; i++)
        vy[i] = (int16_t)(vy[i] + fieldShift);
}

/* VectorSmallMaskYToHalfUV, MaskFun.cpp:178-199 */
static void small_to_uv(const int16_t *y, int n, int16_t *uv, int ratio) {
    for (int i = 0; i < n; i++)
        uv[i] = ratio == 2 ? (int16_t)(y[i] >> 1) : y[i];
}

static int ld_px(const uint8_t *p, int bps) { return bps == 1 ? p[0] : (p[0] | (p[1] << 8)); }

/* flowFetch / flowShift, MVFlow.cpp:87-141, on a finest plane (`ref` points at the interior's first sample) */
static void flow_plane(int mode, uint8_t *dst, ptrdiff_t dst_stride, const uint8_t *ref, ptrdiff_t ref_stride, const int16_t *VX, const int16_t *VY, int vpitch, int width, int height, int time256, int pel, int bps) {
    int lp = pel == 4 ? 2 : (pel == 2 ? 1 : 0);
    if (mode == 0) {
        for (int h = 0; h < height; h++)
            for (int w = 0; w < width; w++) {
                int vx = (VX[h * vpitch + w] * time256 + 128) >> 8;
                int vy = (VY[h * vpitch + w] * time256 + 128) >> 8;
                const uint8_t *s = ref + ((ptrdiff_t)(h << lp) + vy) * ref_stride + ((ptrdiff_t)(w << lp) + vx) * bps;
                memcpy(dst + h * dst_stride + (ptrdiff_t)w * bps, s, (size_t)bps);
            }
    } else {
        int rounding = 128 << lp, shift = 8 + lp;
        for (int h = 0; h < height; h++)
            for (int w = 0; w < width; w++) {
                int vx = (-VX[h * vpitch + w] * time256 + rounding) >> shift;
                int vy = (-VY[h * vpitch + w] * time256 + rounding) >> shift;
                int href = h + vy, wref = w + vx;
                if (href >= 0 && href < height && wref >= 0 && wref < width)
                    memcpy(dst + href * dst_stride + (ptrdiff_t)wref * bps, ref + (ptrdiff_t)(h << lp) * ref_stride + (ptrdiff_t)(w << lp) * bps, (size_t)bps);
            }
    }
    (void)ld_px;
}

/* mvflowGetFrame, MVFlow.cpp:157-368.  ref_super: the super frame of the reference frame (the reference reads the
 * mv.Finest clip made from it; built here the same way).  *usable = 0: the caller returns the source frame. */
int mvo_flow_frame(const mvb200_flow_params *fp, const mvb200_super_params *sp, const uint8_t *const ref_super[3], const ptrdiff_t ref_stride[3], const uint8_t *vectors, int field_shift, uint8_t *const dst[3], const ptrdiff_t dst_stride[3], int32_t *usable) {
    const mvb200_analysis_data *vd = &fp->vd;
    *usable = ref_super != NULL && is_usable(vectors, vd, fp->nSCD1, fp->nSCD2);
    if (!*usable)
        return 0;
    const int bps = (fp->bits + 7) / 8, pel = vd->nPel, pixel_max = (1 << fp->bits) - 1;
    int validity;
    const uint8_t *l0 = blob_level0(vectors, vd, &validity);
    /* the finest planes */
    uint8_t *fin[3] = {NULL, NULL, NULL};
    ptrdiff_t fstride[3] = {0, 0, 0};
    for (int pl = 0; pl < fp->nplanes; pl++) {
        FinestPlane f;
        finest_plane_dims(sp, pl, &f);
        fstride[pl] = (ptrdiff_t)f.pw * pel * bps;
        fin[pl] = malloc((size_t)fstride[pl] * f.ph * pel);
    }
    if (pel == 1) {
        for (int pl = 0; pl < fp->nplanes; pl++) {
            FinestPlane f;
            finest_plane_dims(sp, pl, &f);
            for (int y = 0; y < f.ph; y++)
                memcpy(fin[pl] + y * fstride[pl], ref_super[pl] + y * ref_stride[pl], (size_t)f.pw * bps);
        }
    } else {
        mvb200_super_params all = *sp;
        all.modeyuv = MVB200_YUVPLANES;
        mvo_finest_frame(&all, ref_super, ref_stride, fin, fstride);
    }
    const int nBlkXP = fp->nBlkXP, nBlkYP = fp->nBlkYP, ns = nBlkXP * nBlkYP;
    int16_t *vxs = malloc(sizeof(int16_t) * (size_t)ns), *vys = malloc(sizeof(int16_t) * (size_t)ns);
    small_vectors(l0, vd->nBlkX, vd->nBlkY, nBlkXP, nBlkYP, field_shift, vxs, vys);
    Resizer up;
    resizer_init(&up, fp->nWidthP, fp->nHeightP, nBlkXP, nBlkYP, vd->nWidth, vd->nHeight, pel);
    int16_t *VX = malloc(sizeof(int16_t) * (size_t)fp->nHeightP * fp->VPitchY), *VY = malloc(sizeof(int16_t) * (size_t)fp->nHeightP * fp->VPitchY);
    resize_i16(&up, VX, fp->VPitchY, vxs, nBlkXP, 1);
    resize_i16(&up, VY, fp->VPitchY, vys, nBlkXP, 0);
    if (fp->mode == 1)
        for (int h = 0; h < vd->nHeight; h++)
            for (int w = 0; w < vd->nWidth; w++)
                memcpy(dst[0] + h * dst_stride[0] + (ptrdiff_t)w * bps, &pixel_max, (size_t)bps);
    flow_plane(fp->mode, dst[0], dst_stride[0], fin[0] + (ptrdiff_t)vd->nVPadding * pel * fstride[0] + (ptrdiff_t)vd->nHPadding * pel * bps, fstride[0], VX, VY, fp->VPitchY, vd->nWidth, vd->nHeight, fp->time256, pel, bps);
    if (fp->nplanes > 1) {
        int16_t *vxuv = malloc(sizeof(int16_t) * (size_t)ns), *vyuv = malloc(sizeof(int16_t) * (size_t)ns);
        small_to_uv(vxs, ns, vxuv, vd->xRatioUV);
        small_to_uv(vys, ns, vyuv, vd->yRatioUV);
        Resizer upUV;
        resizer_init(&upUV, fp->nWidthPUV, fp->nHeightPUV, nBlkXP, nBlkYP, fp->nWidthUV, fp->nHeightUV, pel);
        int16_t *VXUV = malloc(sizeof(int16_t) * (size_t)fp->nHeightPUV * fp->VPitchUV), *VYUV = malloc(sizeof(int16_t) * (size_t)fp->nHeightPUV * fp->VPitchUV);
        resize_i16(&upUV, VXUV, fp->VPitchUV, vxuv, nBlkXP, 1);
        resize_i16(&upUV, VYUV, fp->VPitchUV, vyuv, nBlkXP, 0);
        for (int pl = 1; pl < 3; pl++) {
            if (fp->mode == 1)
                for (int h = 0; h < fp->nHeightUV; h++)
                    for (int w = 0; w < fp->nWidthUV; w++)
                        memcpy(dst[pl] + h * dst_stride[pl] + (ptrdiff_t)w * bps, &pixel_max, (size_t)bps);
            flow_plane(fp->mode, dst[pl], dst_stride[pl], fin[pl] + (ptrdiff_t)fp->nVPaddingUV * pel * fstride[pl] + (ptrdiff_t)fp->nHPaddingUV * pel * bps, fstride[pl], VXUV, VYUV, fp->VPitchUV, fp->nWidthUV, fp->nHeightUV,
                       fp->time256, pel, bps);
        }
        resizer_free(&upUV);
        free(vxuv);
        free(vyuv);
        free(VXUV);
        free(VYUV);
    }
    resizer_free(&up);
    free(vxs);
    free(vys);
    free(VX);
    free(VY);
    for (int pl = 0; pl < 3; pl++)
        free(fin[pl]);
    return 0;
}

/* -------------------------------------------------------------- FlowInter */

/* CheckAndPadMaskSmall, MaskFun.cpp:65-82: the nBlkX x nBlkY mask (pitch nBlkXP) grown to nBlkXP x nBlkYP */
static void pad_mask_small(uint8_t *m, int nBlkXP, int nBlkYP, int nBlkX, int nBlkY) {
    if (nBlkXP > nBlkX)
        for (int j = 0; j < nBlkY; j++)
            for (int dx = nBlkX; dx < nBlkXP; dx++)
                m[j * nBlkXP + dx] = m[j * nBlkXP + nBlkX - 1];
    if (nBlkYP > nBlkY)
        for (int i = 0; i < nBlkXP; i++)
            for (int dy = nBlkY; dy < nBlkYP; dy++)
                m[nBlkXP * dy + i] = m[nBlkXP * (nBlkY - 1) + i];
}

typedef struct InterFields { /* full-size planes of one plane type (luma or chroma) */
    int16_t *VXB, *VYB, *VXF, *VYF, *VXBB, *VYBB, *VXFF, *VYFF;
    uint8_t *MB, *MF;
    int pitch;
} InterFields;

static int px(const uint8_t *p, int bps, ptrdiff_t i) { return bps == 1 ? p[i] : ((const uint16_t *)p)[i]; }
static void put_px(uint8_t *p, int bps, ptrdiff_t i, int v) {
    if (bps == 1)
        p[i] = (uint8_t)v;
    else
        ((uint16_t *)p)[i] = (uint16_t)v;
}
static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* FlowInterSimple / FlowInter / FlowInterExtra, MaskFun.cpp:376-546 (the C versions).  prefB / prefF point at the
 * interior's first sample of the finest planes of the right / left frame; ref_pitch in samples. */
static void flow_inter_plane(int method, uint8_t *dst, ptrdiff_t dst_stride, const uint8_t *prefB, const uint8_t *prefF, ptrdiff_t ref_pitch, const InterFields *f, int width, int height, int time256, int pel, int bps) {
    const int lp = pel == 4 ? 2 : (pel == 2 ? 1 : 0);
    for (int h = 0; h < height; h++) {
        const int16_t *VXB = f->VXB + h * f->pitch, *VYB = f->VYB + h * f->pitch, *VXF = f->VXF + h * f->pitch, *VYF = f->VYF + h * f->pitch;
        const uint8_t *MB = f->MB + h * f->pitch, *MF = f->MF + h * f->pitch;
        const ptrdiff_t row = (ptrdiff_t)(h << lp) * ref_pitch;
        for (int w = 0; w < width; w++) {
            const ptrdiff_t at = row + (w << lp);
            int out;
            if (method == 0 && time256 == 128) { /* :464-485 */
                int vxF = VXF[w] >> 1, vyF = VYF[w] >> 1, vxB = VXB[w] >> 1, vyB = VYB[w] >> 1;
                int dstF = px(prefF, bps, at + vyF * ref_pitch + vxF), dstB = px(prefB, bps, at + vyB * ref_pitch + vxB);
                out = (((dstF + dstB) << 8) + (dstB - dstF) * (MF[w] - MB[w])) >> 9;
            } else {
                int vxF = (VXF[w] * time256) >> 8, vyF = (VYF[w] * time256) >> 8;
                int vxB = (VXB[w] * (256 - time256)) >> 8, vyB = (VYB[w] * (256 - time256)) >> 8;
                int dstF = px(prefF, bps, at + vyF * ref_pitch + vxF), dstB = px(prefB, bps, at + vyB * ref_pitch + vxB);
                if (method == 0) { /* :486-510 */
                    out = (((dstF * (255 - MF[w]) + dstB * MF[w] + 255) >> 8) * (256 - time256) + ((dstB * (255 - MB[w]) + dstF * MB[w] + 255) >> 8) * time256) >> 8;
                } else if (method == 1) { /* :376-418, 64-bit intermediates */
                    int64_t dF = dstF, dB = dstB;
                    int dstF0 = px(prefF, bps, at), dstB0 = px(prefB, bps, at);
                    out = (int)((((dF * (255 - MF[w]) + ((MF[w] * (dB * (255 - MB[w]) + MB[w] * dstF0) + 255) >> 8) + 255) >> 8) * (256 - time256) +
                                 ((dB * (255 - MB[w]) + ((MB[w] * (dF * (255 - MF[w]) + MF[w] * dstB0) + 255) >> 8) + 255) >> 8) * time256) >>
                                8);
                } else { /* :421-461 */
                    const int16_t *VXBB = f->VXBB + h * f->pitch, *VYBB = f->VYBB + h * f->pitch, *VXFF = f->VXFF + h * f->pitch, *VYFF = f->VYFF + h * f->pitch;
                    int vxFF = (VXFF[w] * time256) >> 8, vyFF = (VYFF[w] * time256) >> 8;
                    int vxBB = (VXBB[w] * (256 - time256)) >> 8, vyBB = (VYBB[w] * (256 - time256)) >> 8;
                    int dstFF = px(prefF, bps, at + vyFF * ref_pitch + vxFF), dstBB = px(prefB, bps, at + vyBB * ref_pitch + vxBB);
                    int minfb = imin(dstB, dstF), maxfb = imax(dstB, dstF);
                    int medianBB = imax(minfb, imin(dstBB, maxfb)), medianFF = imax(minfb, imin(dstFF, maxfb));
                    out = (((medianBB * MF[w] + dstF * (255 - MF[w]) + 255) >> 8) * (256 - time256) + ((medianFF * MB[w] + dstB * (255 - MB[w]) + 255) >> 8) * time256) >> 8;
                }
            }
            put_px(dst + h * dst_stride, bps, w, out);
        }
    }
}

static void finest_planes(const mvb200_super_params *sp, int nplanes, const uint8_t *const super[3], const ptrdiff_t stride[3], uint8_t *fin[3], ptrdiff_t fstride[3]) {
    const int bps = (sp->bits + 7) / 8, pel = sp->pel;
    mvb200_super_params all = *sp;
    all.modeyuv = MVB200_YUVPLANES;
    for (int pl = 0; pl < 3; pl++) {
        fin[pl] = NULL;
        fstride[pl] = 0;
    }
    for (int pl = 0; pl < nplanes; pl++) {
        FinestPlane f;
        finest_plane_dims(sp, pl, &f);
        fstride[pl] = (ptrdiff_t)f.pw * pel * bps;
        fin[pl] = malloc((size_t)fstride[pl] * f.ph * pel);
        if (pel == 1)
            for (int y = 0; y < f.ph; y++)
                memcpy(fin[pl] + y * fstride[pl], super[pl] + y * stride[pl], (size_t)f.pw * bps);
    }
    if (pel > 1)
        mvo_finest_frame(&all, super, stride, fin, fstride);
}

/* the full-size vector planes of one vector frame for luma (and chroma): MakeVectorSmallMasks + CheckAndPadSmallY +
 * VectorSmallMaskYToHalfUV + simpleResize_int16_t, as in MVFlowInter.c:196-250 */
static void full_vectors(const mvb200_flowinter_params *p, const uint8_t *blob, const Resizer *up, const Resizer *upUV, int16_t *VXY, int16_t *VYY, int16_t *VXUV, int16_t *VYUV) {
    const mvb200_analysis_data *vd = &p->vd;
    const int ns = p->nBlkXP * p->nBlkYP;
    int validity;
    const uint8_t *l0 = blob_level0(blob, vd, &validity);
    int16_t *vxs = malloc(sizeof(int16_t) * (size_t)ns), *vys = malloc(sizeof(int16_t) * (size_t)ns);
    small_vectors(l0, vd->nBlkX, vd->nBlkY, p->nBlkXP, p->nBlkYP, 0, vxs, vys);
    resize_i16(up, VXY, p->VPitchY, vxs, p->nBlkXP, 1);
    resize_i16(up, VYY, p->VPitchY, vys, p->nBlkXP, 0);
    if (p->nplanes > 1) {
        int16_t *vxuv = malloc(sizeof(int16_t) * (size_t)ns), *vyuv = malloc(sizeof(int16_t) * (size_t)ns);
        small_to_uv(vxs, ns, vxuv, vd->xRatioUV);
        small_to_uv(vys, ns, vyuv, vd->yRatioUV);
        resize_i16(upUV, VXUV, p->VPitchUV, vxuv, p->nBlkXP, 1);
        resize_i16(upUV, VYUV, p->VPitchUV, vyuv, p->nBlkXP, 0);
        free(vxuv);
        free(vyuv);
    }
    free(vxs);
    free(vys);
}

/* mvflowinterGetFrame, MVFlowInter.c:83-448 (usable case), also the core of mvflowfpsGetFrame, MVFlowFPS.c:118-470 */
int mvo_flowinter_frame(const mvb200_flowinter_params *p, const mvb200_super_params *sp, int time256, const uint8_t *const left_super[3], const ptrdiff_t left_stride[3], const uint8_t *const right_super[3],
                        const ptrdiff_t right_stride[3], const uint8_t *vec_b, const uint8_t *vec_f, const uint8_t *vec_bb, const uint8_t *vec_ff, uint8_t *const dst[3], const ptrdiff_t dst_stride[3], int32_t *usable) {
    const mvb200_analysis_data *vd = &p->vd;
    *usable = left_super && right_super && vec_b && vec_f && is_usable(vec_f, vd, p->nSCD1, p->nSCD2) && is_usable(vec_b, vd, p->nSCD1, p->nSCD2);
    if (!*usable)
        return 0;
    const int extra = vec_bb && vec_ff && is_usable(vec_ff, vd, p->nSCD1, p->nSCD2) && is_usable(vec_bb, vd, p->nSCD1, p->nSCD2);
    const int method = extra ? p->mode_primary : p->mode_fallback;
    const int bps = (p->bits + 7) / 8, pel = vd->nPel, nplanes = p->nplanes;
    const int stepX = vd->nBlkSizeX - vd->nOverlapX, stepY = vd->nBlkSizeY - vd->nOverlapY;
    uint8_t *finL[3], *finR[3];
    ptrdiff_t fsL[3], fsR[3];
    finest_planes(sp, nplanes, left_super, left_stride, finL, fsL);
    finest_planes(sp, nplanes, right_super, right_stride, finR, fsR);
    Resizer up, upUV;
    resizer_init(&up, p->nWidthP, p->nHeightP, p->nBlkXP, p->nBlkYP, vd->nWidth, vd->nHeight, pel);
    if (nplanes > 1)
        resizer_init(&upUV, p->nWidthPUV, p->nHeightPUV, p->nBlkXP, p->nBlkYP, p->nWidthUV, p->nHeightUV, pel);
    const size_t nY = (size_t)p->nHeightP * p->VPitchY, nUV = nplanes > 1 ? (size_t)p->nHeightPUV * p->VPitchUV : 1;
    InterFields Y, UV;
    memset(&Y, 0, sizeof(Y));
    memset(&UV, 0, sizeof(UV));
    Y.pitch = p->VPitchY;
    UV.pitch = p->VPitchUV;
    int16_t **ys[8] = {&Y.VXB, &Y.VYB, &Y.VXF, &Y.VYF, &Y.VXBB, &Y.VYBB, &Y.VXFF, &Y.VYFF}, **uvs[8] = {&UV.VXB, &UV.VYB, &UV.VXF, &UV.VYF, &UV.VXBB, &UV.VYBB, &UV.VXFF, &UV.VYFF};
    for (int i = 0; i < 8; i++) {
        *ys[i] = calloc(nY, sizeof(int16_t));
        *uvs[i] = calloc(nUV, sizeof(int16_t));
    }
    Y.MB = calloc(nY, 1);
    Y.MF = calloc(nY, 1);
    UV.MB = calloc(nUV, 1);
    UV.MF = calloc(nUV, 1);
    full_vectors(p, vec_b, &up, &upUV, Y.VXB, Y.VYB, UV.VXB, UV.VYB);
    full_vectors(p, vec_f, &up, &upUV, Y.VXF, Y.VYF, UV.VXF, UV.VYF);
    if (method == 2) {
        full_vectors(p, vec_bb, &up, &upUV, Y.VXBB, Y.VYBB, UV.VXBB, UV.VYBB);
        full_vectors(p, vec_ff, &up, &upUV, Y.VXFF, Y.VYFF, UV.VXFF, UV.VYFF);
    }
    /* occlusion masks, :203-208 */
    const int ns = p->nBlkXP * p->nBlkYP;
    uint8_t *mB = calloc((size_t)ns, 1), *mF = calloc((size_t)ns, 1);
    int validity;
    occlusion_mask_time(blob_level0(vec_b, vd, &validity), 1, vd->nBlkX, vd->nBlkY, p->ml, 1.0, pel, mB, p->nBlkXP, 256 - time256, stepX, stepY);
    occlusion_mask_time(blob_level0(vec_f, vd, &validity), 0, vd->nBlkX, vd->nBlkY, p->ml, 1.0, pel, mF, p->nBlkXP, time256, stepX, stepY);
    pad_mask_small(mB, p->nBlkXP, p->nBlkYP, vd->nBlkX, vd->nBlkY);
    pad_mask_small(mF, p->nBlkXP, p->nBlkYP, vd->nBlkX, vd->nBlkY);
    resize_u8(&up, Y.MB, p->VPitchY, mB, p->nBlkXP);
    resize_u8(&up, Y.MF, p->VPitchY, mF, p->nBlkXP);
    if (nplanes > 1) {
        resize_u8(&upUV, UV.MB, p->VPitchUV, mB, p->nBlkXP);
        resize_u8(&upUV, UV.MF, p->VPitchUV, mF, p->nBlkXP);
    }
    for (int pl = 0; pl < nplanes; pl++) {
        const int hp = pl ? p->nHPaddingUV : vd->nHPadding, vp = pl ? p->nVPaddingUV : vd->nVPadding;
        const ptrdiff_t offR = (ptrdiff_t)vp * pel * fsR[pl] + (ptrdiff_t)hp * pel * bps, offL = (ptrdiff_t)vp * pel * fsL[pl] + (ptrdiff_t)hp * pel * bps;
        flow_inter_plane(method, dst[pl], dst_stride[pl], finR[pl] + offR, finL[pl] + offL, fsL[pl] / bps, pl ? &UV : &Y, pl ? p->nWidthUV : vd->nWidth, pl ? p->nHeightUV : vd->nHeight, time256, pel, bps);
    }
    for (int i = 0; i < 8; i++) {
        free(*ys[i]);
        free(*uvs[i]);
    }
    free(Y.MB);
    free(Y.MF);
    free(UV.MB);
    free(UV.MF);
    free(mB);
    free(mF);
    resizer_free(&up);
    if (nplanes > 1)
        resizer_free(&upUV);
    for (int pl = 0; pl < 3; pl++) {
        free(finL[pl]);
        free(finR[pl]);
    }
    return 0;
}

/* Blend, MaskFun.cpp:349-373 */
int mvo_blend_frame(const mvb200_super_params *sp, int time256, const uint8_t *const src[3], const ptrdiff_t src_stride[3], const uint8_t *const ref[3], const ptrdiff_t ref_stride[3], uint8_t *const dst[3],
                    const ptrdiff_t dst_stride[3]) {
    const int bps = (sp->bits + 7) / 8;
    for (int pl = 0; pl < sp->nplanes; pl++) {
        const int w = pl ? sp->width >> sp->ssw : sp->width, h = pl ? sp->height >> sp->ssh : sp->height;
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++)
                put_px(dst[pl] + y * dst_stride[pl], bps, x, (px(src[pl] + y * src_stride[pl], bps, x) * (256 - time256) + px(ref[pl] + y * ref_stride[pl], bps, x) * time256) >> 8);
    }
    return 0;
}

/* --------------------------------------------------------------- FlowBlur */

/* RealFlowBlur, MVFlowBlur.c:72-127 on a finest plane (`ref` points at the interior's first sample) */
static void flow_blur_plane(uint8_t *dst, ptrdiff_t dst_stride, const uint8_t *ref, ptrdiff_t ref_pitch, const int16_t *VXB, const int16_t *VXF, const int16_t *VYB, const int16_t *VYF, int vpitch, int width, int height, int blur256, int prec,
                            int pel, int bps) {
    const int lp = pel == 4 ? 2 : (pel == 2 ? 1 : 0);
    for (int h = 0; h < height; h++) {
        const ptrdiff_t row = (ptrdiff_t)(h << lp) * ref_pitch;
        for (int w = 0; w < width; w++) {
            const ptrdiff_t at = row + (w << lp);
            int bluredsum = px(ref, bps, at);
            int vxF0 = VXF[h * vpitch + w] * blur256, vyF0 = VYF[h * vpitch + w] * blur256;
            int mF = (imax(abs(vxF0), abs(vyF0)) / prec) >> 8;
            if (mF > 0) {
                vxF0 /= mF;
                vyF0 /= mF;
                int vxF = vxF0, vyF = vyF0;
                for (int i = 0; i < mF; i++) {
                    bluredsum += px(ref, bps, at + (vyF >> 8) * ref_pitch + (vxF >> 8));
                    vxF += vxF0;
                    vyF += vyF0;
                }
            }
            int vxB0 = VXB[h * vpitch + w] * blur256, vyB0 = VYB[h * vpitch + w] * blur256;
            int mB = (imax(abs(vxB0), abs(vyB0)) / prec) >> 8;
            if (mB > 0) {
                vxB0 /= mB;
                vyB0 /= mB;
                int vxB = vxB0, vyB = vyB0;
                for (int i = 0; i < mB; i++) {
                    bluredsum += px(ref, bps, at + (vyB >> 8) * ref_pitch + (vxB >> 8));
                    vxB += vxB0;
                    vyB += vyB0;
                }
            }
            put_px(dst + h * dst_stride, bps, w, bluredsum / (mF + mB + 1));
        }
    }
}

/* mvflowblurGetFrame, MVFlowBlur.c:144-337.  vec_b = mvbw of frame n - delta, vec_f = mvfw of frame n + delta. */
int mvo_flowblur_frame(const mvb200_flowblur_params *p, const mvb200_super_params *sp, const uint8_t *const super[3], const ptrdiff_t super_stride[3], const uint8_t *vec_b, const uint8_t *vec_f, uint8_t *const dst[3],
                       const ptrdiff_t dst_stride[3], int32_t *usable) {
    const mvb200_analysis_data *vd = &p->vd;
    *usable = super && vec_b && vec_f && is_usable(vec_f, vd, p->nSCD1, p->nSCD2) && is_usable(vec_b, vd, p->nSCD1, p->nSCD2);
    if (!*usable)
        return 0;
    const int bps = (p->bits + 7) / 8, pel = vd->nPel, nplanes = p->nplanes, nb = vd->nBlkX * vd->nBlkY;
    uint8_t *fin[3];
    ptrdiff_t fs[3];
    finest_planes(sp, nplanes, super, super_stride, fin, fs);
    Resizer up, upUV; /* simpleInit, :531-533: straight from nBlkX x nBlkY to the frame size, no padded field */
    resizer_init(&up, vd->nWidth, vd->nHeight, vd->nBlkX, vd->nBlkY, vd->nWidth, vd->nHeight, pel);
    if (nplanes > 1)
        resizer_init(&upUV, p->nWidthUV, p->nHeightUV, vd->nBlkX, vd->nBlkY, p->nWidthUV, p->nHeightUV, pel);
    const size_t nY = (size_t)vd->nWidth * vd->nHeight, nUV = nplanes > 1 ? (size_t)p->nWidthUV * p->nHeightUV : 1;
    int16_t *full[2][4]; /* [b, f][xY, yY, xUV, yUV] */
    const uint8_t *blobs[2] = {vec_b, vec_f};
    for (int k = 0; k < 2; k++) {
        int validity;
        const uint8_t *l0 = blob_level0(blobs[k], vd, &validity);
        int16_t *vxs = malloc(sizeof(int16_t) * (size_t)nb), *vys = malloc(sizeof(int16_t) * (size_t)nb), *vxuv = malloc(sizeof(int16_t) * (size_t)nb), *vyuv = malloc(sizeof(int16_t) * (size_t)nb);
        small_vectors(l0, vd->nBlkX, vd->nBlkY, vd->nBlkX, vd->nBlkY, 0, vxs, vys);
        small_to_uv(vxs, nb, vxuv, vd->xRatioUV);
        small_to_uv(vys, nb, vyuv, vd->yRatioUV);
        full[k][0] = calloc(nY, sizeof(int16_t));
        full[k][1] = calloc(nY, sizeof(int16_t));
        full[k][2] = calloc(nUV, sizeof(int16_t));
        full[k][3] = calloc(nUV, sizeof(int16_t));
        resize_i16(&up, full[k][0], vd->nWidth, vxs, vd->nBlkX, 1);
        resize_i16(&up, full[k][1], vd->nWidth, vys, vd->nBlkX, 0);
        if (nplanes > 1) {
            resize_i16(&upUV, full[k][2], p->nWidthUV, vxuv, vd->nBlkX, 1);
            resize_i16(&upUV, full[k][3], p->nWidthUV, vyuv, vd->nBlkX, 0);
        }
        free(vxs);
        free(vys);
        free(vxuv);
        free(vyuv);
    }
    for (int pl = 0; pl < nplanes; pl++) {
        const int hp = pl ? p->nHPaddingUV : vd->nHPadding, vp = pl ? p->nVPaddingUV : vd->nVPadding, u = pl ? 2 : 0;
        const int W = pl ? p->nWidthUV : vd->nWidth, H = pl ? p->nHeightUV : vd->nHeight;
        flow_blur_plane(dst[pl], dst_stride[pl], fin[pl] + (ptrdiff_t)vp * pel * fs[pl] + (ptrdiff_t)hp * pel * bps, fs[pl] / bps, full[0][u], full[1][u], full[0][u + 1], full[1][u + 1], W, W, H, p->blur256, p->prec, pel, bps);
    }
    for (int k = 0; k < 2; k++)
        for (int i = 0; i < 4; i++)
            free(full[k][i]);
    resizer_free(&up);
    if (nplanes > 1)
        resizer_free(&upUV);
    for (int pl = 0; pl < 3; pl++)
        free(fin[pl]);
    return 0;
}

/* --------------------------------------------------------------- BlockFPS */

/* MakeSADMaskTime with the arguments BlockFPS passes (MVBlockFPS.c:419-420), gamma = 1 */
/* (sad_mask_time above takes the bit depth: the SAD is scaled down to 8 bits, MaskFun.cpp:157) */

typedef struct SuperPlaneView { /* level 0 of one plane of a super frame (mvpUpdate / mvpGetPointer, MVFrame.cpp:1356-1364,1732-1734) */
    const uint8_t *base;
    ptrdiff_t pitch;
    int hpad, vpad, padded_h, pel, lp, bps;
} SuperPlaneView;

static const uint8_t *mvp_pointer(const SuperPlaneView *v, int nX, int nY) {
    int ax = nX + v->hpad * v->pel, ay = nY + v->vpad * v->pel;
    int idx = (ax & (v->pel - 1)) | ((ay & (v->pel - 1)) << v->lp);
    return v->base + ((ptrdiff_t)idx * v->padded_h + (ay >> v->lp)) * v->pitch + (ptrdiff_t)(ax >> v->lp) * v->bps;
}

static int median3(int a, int b, int c) { /* MEDIAN, MVBlockFPS.c:101-110 */
    int mn = imin(a, b), mx = imax(a, b);
    return imax(mn, imin(mx, c));
}

/* RealResultBlock, MVBlockFPS.c:116-222 */
static void result_block(uint8_t *pDst, ptrdiff_t dst_pitch, const uint8_t *pMCB, ptrdiff_t MCB_pitch, const uint8_t *pMCF, ptrdiff_t MCF_pitch, const uint8_t *pRef, ptrdiff_t ref_pitch, const uint8_t *pSrc, ptrdiff_t src_pitch,
                         const uint8_t *maskB, int mask_pitch, const uint8_t *maskF, const uint8_t *pOcc, int bw, int bh, int time256, int mode, int bits) {
    const int bps = (bits + 7) / 8;
    for (int h = 0; h < bh; h++) {
        for (int w = 0; w < bw; w++) {
            int mcb = px(pMCB + h * MCB_pitch, bps, w), mcf = px(pMCF + h * MCF_pitch, bps, w), out;
            if (mode == 0) {
                out = (mcb * time256 + mcf * (256 - time256)) >> 8;
            } else if (mode == 1) {
                int mca = (mcb * time256 + mcf * (256 - time256)) >> 8;
                out = median3(px(pRef + h * ref_pitch, bps, w), px(pSrc + h * src_pitch, bps, w), mca);
            } else if (mode == 2) {
                int avg = (px(pRef + h * ref_pitch, bps, w) * time256 + px(pSrc + h * src_pitch, bps, w) * (256 - time256)) >> 8;
                out = median3(avg, mcb, mcf);
            } else if (mode == 3 || mode == 6) {
                int mB = maskB[h * mask_pitch + w], mF = maskF[h * mask_pitch + w];
                out = (((mB * mcf + (255 - mB) * mcb + 255) >> 8) * time256 + ((mF * mcb + (255 - mF) * mcf + 255) >> 8) * (256 - time256)) >> 8;
            } else if (mode == 4 || mode == 7) {
                int mB = maskB[h * mask_pitch + w], mF = maskF[h * mask_pitch + w], occ = pOcc[h * mask_pitch + w];
                int f = (mF * mcb + (255 - mF) * mcf + 255) >> 8;
                int b = (mB * mcf + (255 - mB) * mcb + 255) >> 8;
                int avg = (px(pRef + h * ref_pitch, bps, w) * time256 + px(pSrc + h * src_pitch, bps, w) * (256 - time256) + 255) >> 8;
                int m = (b * time256 + f * (256 - time256)) >> 8;
                out = (avg * occ + m * (255 - occ) + 255) >> 8;
            } else {
                out = pOcc[h * mask_pitch + w] << (bits - 8);
            }
            put_px(pDst + h * dst_pitch, bps, w, out);
        }
    }
}

static void blend_rect(uint8_t *dst, const uint8_t *src, const uint8_t *ref, int height, int width, ptrdiff_t dp, ptrdiff_t sp, ptrdiff_t rp, int time256, int bps) { /* Blend, MaskFun.cpp:349-373 */
    for (int h = 0; h < height; h++)
        for (int w = 0; w < width; w++)
            put_px(dst + h * dp, bps, w, (px(src + h * sp, bps, w) * (256 - time256) + px(ref + h * rp, bps, w) * time256) >> 8);
}

/* mvblockfpsGetFrame, MVBlockFPS.c:240-640 (usable case) */
int mvo_blockfps_frame(const mvb200_blockfps_params *p, const mvb200_super_params *sp, int time256, const uint8_t *const left_super[3], const ptrdiff_t left_stride[3], const uint8_t *const right_super[3],
                       const ptrdiff_t right_stride[3], const uint8_t *vec_b, const uint8_t *vec_f, uint8_t *const dst[3], const ptrdiff_t dst_stride[3], int32_t *usable) {
    const mvb200_analysis_data *vd = &p->vd;
    *usable = left_super && right_super && vec_b && vec_f && is_usable(vec_f, vd, p->nSCD1, p->nSCD2) && is_usable(vec_b, vd, p->nSCD1, p->nSCD2);
    if (!*usable)
        return 0;
    const int bits = p->bits, bps = (bits + 7) / 8, pel = vd->nPel, lp = pel == 4 ? 2 : (pel == 2 ? 1 : 0), mode = p->mode;
    const int planes = (p->super_modeyuv & (MVB200_UPLANE | MVB200_VPLANE)) ? 3 : 1;
    const int xR[3] = {1, vd->xRatioUV, vd->xRatioUV}, yR[3] = {1, vd->yRatioUV, vd->yRatioUV};
    const int nBlkX = vd->nBlkX, nBlkY = vd->nBlkY, nBlkXP = p->nBlkXP, nBlkYP = p->nBlkYP;
    int W[3], H[3], bw[3], bh[3], ox[3], oy[3], WB[3], HB[3], pitchM[3] = {p->nPitchY, p->nPitchUV, p->nPitchUV};
    for (int i = 0; i < 3; i++) {
        W[i] = i ? p->nWidthUV : vd->nWidth;
        H[i] = i ? p->nHeightUV : vd->nHeight;
        bw[i] = vd->nBlkSizeX / xR[i];
        bh[i] = vd->nBlkSizeY / yR[i];
        ox[i] = vd->nOverlapX / xR[i];
        oy[i] = vd->nOverlapY / yR[i];
        WB[i] = (nBlkX * (vd->nBlkSizeX - vd->nOverlapX) + vd->nOverlapX) / xR[i];
        HB[i] = (nBlkY * (vd->nBlkSizeY - vd->nOverlapY) + vd->nOverlapY) / yR[i];
    }
    int validity;
    const uint8_t *l0B = blob_level0(vec_b, vd, &validity), *l0F = blob_level0(vec_f, vd, &validity);
    /* masks, :393-440 */
    uint8_t *MB[3] = {NULL, NULL, NULL}, *MF[3] = {NULL, NULL, NULL}, *MO[3] = {NULL, NULL, NULL};
    MB[0] = calloc((size_t)p->nHeightP * pitchM[0], 1);
    MF[0] = calloc((size_t)p->nHeightP * pitchM[0], 1);
    MO[0] = calloc((size_t)p->nHeightP * pitchM[0], 1);
    if (planes > 1) {
        MB[1] = MB[2] = calloc((size_t)p->nHeightPUV * pitchM[1], 1);
        MF[1] = MF[2] = calloc((size_t)p->nHeightPUV * pitchM[1], 1);
        MO[1] = MO[2] = calloc((size_t)p->nHeightPUV * pitchM[1], 1);
    }
    if (mode >= 3) {
        const int ns = nBlkXP * nBlkYP, stepX = vd->nBlkSizeX - vd->nOverlapX, stepY = vd->nBlkSizeY - vd->nOverlapY;
        uint8_t *sB = calloc((size_t)ns, 1), *sF = calloc((size_t)ns, 1), *sO = calloc((size_t)ns, 1);
        if (mode <= 5) {
            occlusion_mask_time(l0F, 0, nBlkX, nBlkY, p->ml, 1.0, pel, sF, nBlkXP, time256, stepX, stepY);
            occlusion_mask_time(l0B, 1, nBlkX, nBlkY, p->ml, 1.0, pel, sB, nBlkXP, 256 - time256, stepX, stepY);
        } else {
            sad_mask_time(l0F, nBlkX, nBlkY, 4.0 / (p->ml * vd->nBlkSizeX * vd->nBlkSizeY), 1.0, pel, sF, nBlkXP, time256, stepX, stepY, bits);
            sad_mask_time(l0B, nBlkX, nBlkY, 4.0 / (p->ml * vd->nBlkSizeX * vd->nBlkSizeY), 1.0, pel, sB, nBlkXP, 256 - time256, stepX, stepY, bits);
        }
        pad_mask_small(sF, nBlkXP, nBlkYP, nBlkX, nBlkY);
        pad_mask_small(sB, nBlkXP, nBlkYP, nBlkX, nBlkY);
        Resizer up, upUV;
        resizer_init(&up, p->nWidthP, p->nHeightP, nBlkXP, nBlkYP, vd->nWidth, vd->nHeight, pel);
        resize_u8(&up, MF[0], pitchM[0], sF, nBlkXP);
        resize_u8(&up, MB[0], pitchM[0], sB, nBlkXP);
        if (planes > 1) {
            resizer_init(&upUV, p->nWidthPUV, p->nHeightPUV, nBlkXP, nBlkYP, p->nWidthUV, p->nHeightUV, pel);
            resize_u8(&upUV, MF[1], pitchM[1], sF, nBlkXP);
            resize_u8(&upUV, MB[1], pitchM[1], sB, nBlkXP);
        }
        if (mode == 4 || mode == 5 || mode == 7 || mode == 8) {
            for (int i = 0; i < ns; i++) /* MultMasks, :90-98 */
                sO[i] = (uint8_t)((sF[i] * sB[i]) / 255);
            resize_u8(&up, MO[0], pitchM[0], sO, nBlkXP);
            if (planes > 1)
                resize_u8(&upUV, MO[1], pitchM[1], sO, nBlkXP);
        }
        resizer_free(&up);
        if (planes > 1)
            resizer_free(&upUV);
        free(sB);
        free(sF);
        free(sO);
    }
    /* plane views and the source / reference pointers at the interior (:443-452; chroma offsets are pad >> 1 whatever the subsampling) */
    SuperPlaneView vB[3], vF[3];
    const uint8_t *pRef[3], *pSrc[3];
    for (int i = 0; i < planes; i++) {
        FinestPlane f;
        finest_plane_dims(sp, i, &f);
        vB[i].base = right_super[i];
        vB[i].pitch = right_stride[i];
        vF[i].base = left_super[i];
        vF[i].pitch = left_stride[i];
        vB[i].hpad = vF[i].hpad = i ? sp->hpad / sp->xRatioUV : sp->hpad;
        vB[i].vpad = vF[i].vpad = i ? sp->vpad / sp->yRatioUV : sp->vpad;
        vB[i].padded_h = vF[i].padded_h = f.ph;
        vB[i].pel = vF[i].pel = pel;
        vB[i].lp = vF[i].lp = lp;
        vB[i].bps = vF[i].bps = bps;
        const int hp = i ? sp->hpad >> 1 : sp->hpad, vp = i ? sp->vpad >> 1 : sp->vpad;
        pSrc[i] = left_super[i] + (ptrdiff_t)hp * bps + left_stride[i] * vp;
        pRef[i] = right_super[i] + (ptrdiff_t)hp * bps + right_stride[i] * vp;
    }
    const int stepX0 = vd->nBlkSizeX - vd->nOverlapX, stepY0 = vd->nBlkSizeY - vd->nOverlapY;
    if (vd->nOverlapX == 0 && vd->nOverlapY == 0) { /* :466-520 */
        for (int i = 0; i < nBlkX * nBlkY; i++) {
            const int bx = i % nBlkX, by = i / nBlkX;
            VEC B = get_vec(l0B, i), F = get_vec(l0F, i);
            for (int pl = 0; pl < planes; pl++) {
                const int x0 = bx * bw[pl], y0 = by * bh[pl];
                const uint8_t *mcb = mvp_pointer(&vB[pl], (bx * stepX0 * pel + ((B.x * (256 - time256)) >> 8)) / xR[pl], (by * stepY0 * pel + ((B.y * (256 - time256)) >> 8)) / yR[pl]);
                const uint8_t *mcf = mvp_pointer(&vF[pl], (bx * stepX0 * pel + ((F.x * time256) >> 8)) / xR[pl], (by * stepY0 * pel + ((F.y * time256) >> 8)) / yR[pl]);
                result_block(dst[pl] + y0 * dst_stride[pl] + (ptrdiff_t)x0 * bps, dst_stride[pl], mcb, vB[pl].pitch, mcf, vF[pl].pitch, pRef[pl] + y0 * right_stride[pl] + (ptrdiff_t)x0 * bps, right_stride[pl],
                             pSrc[pl] + y0 * left_stride[pl] + (ptrdiff_t)x0 * bps, left_stride[pl], MB[pl] + y0 * pitchM[pl] + x0, pitchM[pl], MF[pl] + y0 * pitchM[pl] + x0, MO[pl] + y0 * pitchM[pl] + x0, bw[pl], bh[pl], time256, mode,
                             bits);
            }
        }
        for (int pl = 0; pl < planes; pl++) { /* the rest right of and below the blocks: time-weighted blend */
            const int xr = bw[pl] * nBlkX, yb = bh[pl] * nBlkY;
            blend_rect(dst[pl] + (ptrdiff_t)xr * bps, pSrc[pl] + (ptrdiff_t)xr * bps, pRef[pl] + (ptrdiff_t)xr * bps, yb, W[pl] - xr, dst_stride[pl], left_stride[pl], right_stride[pl], time256, bps);
            blend_rect(dst[pl] + yb * dst_stride[pl], pSrc[pl] + yb * left_stride[pl], pRef[pl] + yb * right_stride[pl], H[pl] - yb, W[pl], dst_stride[pl], left_stride[pl], right_stride[pl], time256, bps);
        }
    } else { /* overlap, :521-606 */
        int16_t *winY = malloc(sizeof(int16_t) * 9 * (size_t)bw[0] * bh[0]), *winUV = NULL;
        mvo_overlap_windows(bw[0], bh[0], ox[0], oy[0], winY);
        if (planes > 1) {
            winUV = malloc(sizeof(int16_t) * 9 * (size_t)bw[1] * bh[1]);
            mvo_overlap_windows(bw[1], bh[1], ox[1], oy[1], winUV);
        }
        uint8_t *tmp = malloc((size_t)bh[0] * bw[0] * bps);
        for (int pl = 0; pl < planes; pl++) {
            blend_rect(dst[pl] + (ptrdiff_t)WB[pl] * bps, pSrc[pl] + (ptrdiff_t)WB[pl] * bps, pRef[pl] + (ptrdiff_t)WB[pl] * bps, HB[pl], W[pl] - WB[pl], dst_stride[pl], left_stride[pl], right_stride[pl], time256, bps);
            blend_rect(dst[pl] + dst_stride[pl] * HB[pl], pSrc[pl] + left_stride[pl] * HB[pl], pRef[pl] + right_stride[pl] * HB[pl], H[pl] - HB[pl], W[pl], dst_stride[pl], left_stride[pl], right_stride[pl], time256, bps);
            uint32_t *acc = calloc((size_t)W[pl] * H[pl], sizeof(uint32_t)); /* DstTemp: uint16 (8-bit clips) or uint32 sums */
            const int16_t *win9 = pl ? winUV : winY;
            for (int by = 0; by < nBlkY; by++) {
                const int wby = ((by + nBlkY - 3) / (nBlkY - 2)) * 3;
                for (int bx = 0; bx < nBlkX; bx++) {
                    const int wbx = bx == nBlkX - 1 ? 2 : (bx == 0 ? 0 : 1);
                    const int16_t *win = win9 + (size_t)(wby + wbx) * bw[pl] * bh[pl];
                    const int i = by * nBlkX + bx, x0 = bx * (bw[pl] - ox[pl]), y0 = by * (bh[pl] - oy[pl]);
                    VEC B = get_vec(l0B, i), F = get_vec(l0F, i);
                    const uint8_t *mcb = mvp_pointer(&vB[pl], (bx * stepX0 * pel + ((B.x * (256 - time256)) >> 8)) / xR[pl], (by * stepY0 * pel + ((B.y * (256 - time256)) >> 8)) / yR[pl]);
                    const uint8_t *mcf = mvp_pointer(&vF[pl], (bx * stepX0 * pel + ((F.x * time256) >> 8)) / xR[pl], (by * stepY0 * pel + ((F.y * time256) >> 8)) / yR[pl]);
                    result_block(tmp, (ptrdiff_t)bw[pl] * bps, mcb, vB[pl].pitch, mcf, vF[pl].pitch, pRef[pl] + y0 * right_stride[pl] + (ptrdiff_t)x0 * bps, right_stride[pl], pSrc[pl] + y0 * left_stride[pl] + (ptrdiff_t)x0 * bps,
                                 left_stride[pl], MB[pl] + y0 * pitchM[pl] + x0, pitchM[pl], MF[pl] + y0 * pitchM[pl] + x0, MO[pl] + y0 * pitchM[pl] + x0, bw[pl], bh[pl], time256, mode, bits);
                    for (int h = 0; h < bh[pl]; h++) /* overlaps_c, Overlap.cpp:144-158 */
                        for (int w = 0; w < bw[pl]; w++) {
                            uint32_t *a = &acc[(size_t)(y0 + h) * W[pl] + x0 + w];
                            *a += (uint32_t)((px(tmp + (ptrdiff_t)h * bw[pl] * bps, bps, w) * win[h * bw[pl] + w]) >> 6);
                            if (bps == 1)
                                *a &= 0xffffu;
                        }
                }
            }
            const int pixelMax = (1 << bits) - 1;
            for (int h = 0; h < HB[pl]; h++) /* ToPixels, Overlap.cpp:335-356 */
                for (int w = 0; w < WB[pl]; w++) {
                    int a = (int)((acc[(size_t)h * W[pl] + w] + 16) >> 5);
                    put_px(dst[pl] + h * dst_stride[pl], bps, w, a > pixelMax ? pixelMax : a);
                }
            free(acc);
        }
        free(tmp);
        free(winY);
        free(winUV);
    }
    free(MB[0]);
    free(MF[0]);
    free(MO[0]);
    if (planes > 1) {
        free(MB[1]);
        free(MF[1]);
        free(MO[1]);
    }
    return 0;
}
