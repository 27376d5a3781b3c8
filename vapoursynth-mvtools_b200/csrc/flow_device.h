// Per-thread arithmetic of the vector-clip consumers (mv.SCDetection / mv.Mask / mv.Finest / mv.Flow,
// SURVEY.md section 8f), written as plain functions of (output coordinate, parameter block) so that the
// kernels in flow.cu are thin index wrappers around them.  The same functions also compile as host C++
// (MVB_HD drops the CUDA qualifiers): tests/emul/flow_emul.cpp runs them thread by thread on the CPU and
// compares with the oracle -- a test of the index arithmetic that needs no GPU.  That harness is test
// infrastructure; the product never executes these functions on the host.
//
// Reference semantics (dubhater/vapoursynth-mvtools @ e516e90):
//   the reference builds full-resolution intermediates (small per-block fields -> SimpleResize to
//   full-size int16 vector planes / uint8 mask planes -> per-pixel pass; mv.Flow additionally reads a
//   materialised mv.Finest clip).  Here every output sample is a pure function of the per-block field
//   and the reference SUPER frame: the bilinear upsizing (4 taps) and the sub-plane selection happen in
//   registers, nothing full-size is staged.
#pragma once

#include <stdint.h>

#ifdef __CUDACC__
#define MVB_HD __host__ __device__ __forceinline__
#else
#define MVB_HD inline
#include <cmath>
#endif

namespace mvb200 {
namespace flowdev {

constexpr int kWShift = 14, kWMax = 1 << kWShift, kWHalf = kWMax / 2; // SimpleResize.h:31-35

// the level-0 VECTOR array of a blob (fgopUpdate, Fakery.c:110-121): the blob ends with it (4-byte aligned only)
MVB_HD const int *blob_level0(const uint8_t *blob, int nBlk) { return (const int *)(blob + *(const int *)blob - 16 * nBlk); }
MVB_HD long long vec_sad(const int *v) { return (long long)(((unsigned long long)(unsigned)v[3] << 32) | (unsigned)v[2]); }

// One output sample of simpleResize<T> (SimpleResize.cpp:60-116): the vertical pass at columns o and o + 1
// (stored as T in the reference's work row), then the horizontal pass.
template <typename T>
MVB_HD int resize_sample(const T *src, int src_stride, int voff, int vwt, int hoff, int hwt) {
    const T *s1 = src + voff * src_stride, *s2 = s1 + src_stride;
    const int wt = kWMax - vwt;
    const T a = (T)(((int)s1[hoff] * wt + (int)s2[hoff] * vwt + kWHalf) >> kWShift);
    const T b = (T)(((int)s1[hoff + 1] * wt + (int)s2[hoff + 1] * vwt + kWHalf) >> kWShift);
    return ((int)a * (kWMax - hwt) + (int)b * hwt + kWHalf) >> kWShift;
}

// the running limits of the 16-bit resizer (SimpleResize.cpp:66-70,83-86,97-102,108-111) in closed form: a
// horizontal vector at column x stays in [-x pel, limit_w pel - 1 - x pel], a vertical one at row y likewise
MVB_HD int limit_vector(int v, int pos, int limit, int pel) {
    const int minimum = -pos * pel, maximum = limit * pel - 1 - pos * pel;
    v = v < maximum ? v : maximum;
    return v > minimum ? v : minimum;
}

// ------------------------------------------------------------------ small (per-block) fields

// MakeVectorSmallMasks (MaskFun.cpp:163-176) + CheckAndPadSmallY (:38-62) + the field shift (MVFlow.cpp:301-305)
// + VectorSmallMaskYToHalfUV (:178-199): entry (i, j) of the padded nBlkXP x nBlkYP field.
struct SmallVec {
    short xY, yY, xUV, yUV;
};
MVB_HD SmallVec small_vector(const int *level0, int nBlkX, int nBlkY, int i, int j, int fieldShift, int xRatioUV, int yRatioUV) {
    const int bx = i < nBlkX ? i : nBlkX - 1, by = j < nBlkY ? j : nBlkY - 1;
    const int *v = level0 + 4 * (by * nBlkX + bx);
    short x = (short)v[0], y = (short)v[1];
    if (i >= nBlkX) // right fill: not positive
        x = x < 0 ? x : (short)0;
    if (j >= nBlkY) // bottom fill (after the right fill): not positive
        y = y < 0 ? y : (short)0;
    y = (short)(y + fieldShift);
    SmallVec s;
    s.xY = x;
    s.yY = y;
    s.xUV = xRatioUV == 2 ? (short)(x >> 1) : x;
    s.yUV = yRatioUV == 2 ? (short)(y >> 1) : y;
    return s;
}

struct MaskSmallParams {
    int kind, nBlkX, nBlkY, nPel, isBackward, time256, stepX, stepY, blkSizeXY;
    float fGamma, fMaskNormFactor, fMaskNormFactor2, fHalfGamma;
};

MVB_HD double pow_like(double x, double y) { // pow() with the exactly-rounded cases taken out (y = 1, y = 0.5)
    if (y == 1.0)
        return x;
    if (y == 0.5)
        return sqrt(x);
    return pow(x, y);
}

MVB_HD int clamp255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

// kinds 0, 1, 3, 4, 5 of mvmaskGetFrame (MVMask.c:141-160): entry j of the small mask(s)
MVB_HD void mask_small_direct(const MaskSmallParams &P, const int *level0, int j, int *small, int *smallV) {
    const int *v = level0 + 4 * j;
    if (P.kind == 0) { // mvmaskLength, MVMask.c:65-71
        const double norme = (double)(v[0] * v[0] + v[1] * v[1]) / (P.nPel * P.nPel);
        const double l = 255 * pow_like(norme * P.fMaskNormFactor2, P.fHalfGamma);
        small[j] = (int)(unsigned char)((l > 255) ? 255 : l);
    } else if (P.kind == 1) { // MakeSADMaskTime, MaskFun.cpp:136-160 (8-bit clip: no shift of the SAD)
        const int time4096X = (256 - P.time256) * 16 / (P.stepX * P.nPel);
        const int time4096Y = (256 - P.time256) * 16 / (P.stepY * P.nPel);
        const int bx = j % P.nBlkX, by = j / P.nBlkX;
        int bxi = bx - v[0] * time4096X / 4096;
        int byi = by - v[1] * time4096Y / 4096;
        if (bxi < 0 || bxi >= P.nBlkX || byi < 0 || byi >= P.nBlkY) {
            bxi = bx;
            byi = by;
        }
        const long long sad = vec_sad(level0 + 4 * (bxi + byi * P.nBlkX));
        const double dSADNormFactor = 4.0 * P.fMaskNormFactor / P.blkSizeXY;
        const double l = 255 * pow_like(sad * dSADNormFactor, P.fGamma); // ByteNorm, :129-133
        small[j] = (int)(unsigned char)((l > 255) ? 255 : l);
    } else if (P.kind == 3) {
        small[j] = clamp255((int)(v[0] * P.fMaskNormFactor * 100 + 128));
    } else if (P.kind == 4) {
        small[j] = clamp255((int)(v[1] * P.fMaskNormFactor * 100 + 128));
    } else if (P.kind == 5) {
        small[j] = clamp255((int)(v[0] * P.fMaskNormFactor * 100 + 128));
        smallV[j] = clamp255((int)(v[1] * P.fMaskNormFactor * 100 + 128));
    }
}

// ByteOccMask, MaskFun.cpp:85-90 (the value; the caller takes the maximum)
MVB_HD int occ_value(int occlusion, double occnorm, double fGamma) {
    int v;
    if (fGamma == 1.0)
        v = (int)(255 * occlusion * occnorm);
    else
        v = (int)(255 * pow_like(occlusion * occnorm, fGamma));
    return v < 255 ? v : 255;
}

// MakeVectorOcclusionMaskTime, MaskFun.cpp:92-126: what block (bx, by) contributes.  `update(index, value)` raises
// entry `index` of the (zeroed) small mask to at least `value`: a maximum, so the order of the blocks does not matter.
template <typename Update>
MVB_HD void occlusion_contrib(const int *level0, int isBackward, int nBlkX, int nBlkY, double dMaskNormDivider, double fGamma, int nPel, int time256, int stepX, int stepY, int bx, int by, Update update) {
    const int time4096X = time256 * 16 / (stepX * nPel);
    const int time4096Y = time256 * 16 / (stepY * nPel);
    const double occnormX = 80.0 / (dMaskNormDivider * stepX * nPel);
    const double occnormY = 80.0 / (dMaskNormDivider * stepY * nPel);
    const int i = bx + by * nBlkX;
    const int vx = level0[4 * i], vy = level0[4 * i + 1];
    if (bx < nBlkX - 1) {
        const int vx1 = level0[4 * (i + 1)];
        if (vx1 < vx) {
            const int occlusion = vx - vx1, lo = bx + 1 - occlusion * time4096X / 4096;
            const int minb = isBackward ? (lo > 0 ? lo : 0) : bx;
            const int maxb = isBackward ? bx + 1 : (lo < nBlkX - 1 ? lo : nBlkX - 1);
            const int val = occ_value(occlusion, occnormX, fGamma);
            for (int bxi = minb; bxi <= maxb; bxi++)
                update(bxi + by * nBlkX, val);
        }
    }
    if (by < nBlkY - 1) {
        const int vy1 = level0[4 * (i + nBlkX) + 1];
        if (vy1 < vy) {
            const int occlusion = vy - vy1, lo = by + 1 - occlusion * time4096Y / 4096;
            const int minb = isBackward ? (lo > 0 ? lo : 0) : by;
            const int maxb = isBackward ? by + 1 : (lo < nBlkY - 1 ? lo : nBlkY - 1);
            const int val = occ_value(occlusion, occnormY, fGamma);
            for (int byi = minb; byi <= maxb; byi++)
                update(bx + byi * nBlkX, val);
        }
    }
}

// ------------------------------------------------------------------ Mask, one output sample

struct ResizeTables { // device pointers: offsets / weights of the horizontal (dst_w) and vertical (dst_h) passes
    const int *hoff, *hwt, *voff, *vwt;
};

struct MaskPlane {
    int plane, W, H, WB, HB; // plane size and the part the blocks cover (MVMask.c:303-311)
    ResizeTables t;
};

// mvmaskGetFrame, MVMask.c:152-196: sample (x, y) of plane B.plane; small masks are int arrays of nBlkX x nBlkY
MVB_HD int mask_sample(const MaskPlane &B, int kind, int usable, int ysc, int nBlkX, const int *small, const int *smallV, const uint8_t *src_luma, long long src_pitch, int x, int y) {
    if (B.plane == 0 && kind == 5)
        return src_luma[(long long)y * src_pitch + x];
    if (!usable)
        return ysc;
    const int xx = x < B.WB ? x : B.WB - 1, yy = y < B.HB ? y : B.HB - 1; // right / bottom strips repeat the last covered column / row
    const int *m = (B.plane == 2 && kind == 5) ? smallV : small;
    return resize_sample<int>(m, nBlkX, B.t.voff[yy], B.t.vwt[yy], B.t.hoff[xx], B.t.hwt[xx]) & 255;
}

// ------------------------------------------------------------------ Finest / Flow

struct SuperPlane { // level 0 of one plane of a super frame
    long long pitch, offset, sub_stride; // sub_stride = pitch * padded_h: distance between pel sub-planes
    int hpad, vpad, padded_w, padded_h;
};

// sample (fx, fy) of the pel-times upsized padded plane = sub-plane ((fx & m) | (fy & m) << logPel) at (fx >> logPel,
// fy >> logPel) (Merge4PlanesToBig / Merge16PlanesToBig, MaskFun.cpp:196-336; mvpGetAbsolutePointer)
template <typename T>
MVB_HD T finest_sample(const uint8_t *super_frame, const SuperPlane &S, int logPel, int fx, int fy) {
    const int m = (1 << logPel) - 1;
    const int idx = (fx & m) | ((fy & m) << logPel);
    return *(const T *)(super_frame + S.offset + idx * S.sub_stride + (long long)(fy >> logPel) * S.pitch + (long long)(fx >> logPel) * (int)sizeof(T));
}

struct FlowPlane {
    int plane, W, H;          // output plane
    int limitW, limitH;       // the resizer's limits: the plane's frame size (simpleInit, MVFlow.cpp:561-563)
    int time256, pel, logPel, mode;
    int nBlkXP;               // pitch of the small fields
    const short *vxs, *vys;   // small fields of this plane (luma, or the chroma-scaled copies)
    ResizeTables t;
    SuperPlane S;
};

// the upsized vector at (x, y): simpleResize<int16_t> with horizontal / vertical limits
MVB_HD void flow_vector(const FlowPlane &F, int x, int y, int &VX, int &VY) {
    const int voff = F.t.voff[y], vwt = F.t.vwt[y], hoff = F.t.hoff[x], hwt = F.t.hwt[x];
    VX = (short)limit_vector(resize_sample<short>(F.vxs, F.nBlkXP, voff, vwt, hoff, hwt), x, F.limitW, F.pel);
    VY = (short)limit_vector(resize_sample<short>(F.vys, F.nBlkXP, voff, vwt, hoff, hwt), y, F.limitH, F.pel);
}

// flowFetch, MVFlow.cpp:87-110
template <typename T>
MVB_HD T flow_fetch_sample(const FlowPlane &F, const uint8_t *ref_super, int x, int y) {
    int VX, VY;
    flow_vector(F, x, y, VX, VY);
    const int vx = (VX * F.time256 + 128) >> 8;
    const int vy = (VY * F.time256 + 128) >> 8;
    return finest_sample<T>(ref_super, F.S, F.logPel, ((x + F.S.hpad) << F.logPel) + vx, ((y + F.S.vpad) << F.logPel) + vy);
}

// flowShift, MVFlow.cpp:113-141: where source sample (x, y) lands, or -1 (the reference's loop lets the LAST source
// sample in raster order win a destination: the kernels keep the largest source index per destination)
MVB_HD int flow_shift_target(const FlowPlane &F, int x, int y) {
    int VX, VY;
    flow_vector(F, x, y, VX, VY);
    const int rounding = 128 << F.logPel, shift = 8 + F.logPel;
    const int vx = (-VX * F.time256 + rounding) >> shift;
    const int vy = (-VY * F.time256 + rounding) >> shift;
    const int href = y + vy, wref = x + vx;
    if (href >= 0 && href < F.H && wref >= 0 && wref < F.W)
        return href * F.W + wref;
    return -1;
}

// ------------------------------------------------------------------ FlowInter / FlowFPS

// CheckAndPadMaskSmall, MaskFun.cpp:65-82: entry (i, j) of the nBlkXP x nBlkYP mask grown from nBlkX x nBlkY (right
// fill first, then the bottom fill from the already filled last row: a clamp of both indices)
MVB_HD int padded_mask_entry(const int *raw, int nBlkX, int nBlkY, int i, int j) {
    return raw[(j < nBlkY ? j : nBlkY - 1) * nBlkX + (i < nBlkX ? i : nBlkX - 1)];
}

struct InterPlane {
    int plane, W, H, limitW, limitH, pel, logPel, nBlkXP, time256;
    const short *vxB, *vyB, *vxF, *vyF, *vxBB, *vyBB, *vxFF, *vyFF; // small fields of this plane
    const int *maskB, *maskF;                                       // padded small occlusion masks
    ResizeTables t;
    SuperPlane S;
};

MVB_HD void inter_vector(const InterPlane &P, const short *vxs, const short *vys, int x, int y, int &VX, int &VY) {
    const int voff = P.t.voff[y], vwt = P.t.vwt[y], hoff = P.t.hoff[x], hwt = P.t.hwt[x];
    VX = (short)limit_vector(resize_sample<short>(vxs, P.nBlkXP, voff, vwt, hoff, hwt), x, P.limitW, P.pel);
    VY = (short)limit_vector(resize_sample<short>(vys, P.nBlkXP, voff, vwt, hoff, hwt), y, P.limitH, P.pel);
}

// One output sample of FlowInterSimple (method 0) / FlowInter (1) / FlowInterExtra (2), MaskFun.cpp:376-546.
// right_super = the frame the backward vectors point at (prefB), left_super = the current frame (prefF).
template <typename T>
MVB_HD int flow_inter_sample(const InterPlane &P, int method, const uint8_t *right_super, const uint8_t *left_super, int x, int y) {
    const int voff = P.t.voff[y], vwt = P.t.vwt[y], hoff = P.t.hoff[x], hwt = P.t.hwt[x];
    const int MB = resize_sample<int>(P.maskB, P.nBlkXP, voff, vwt, hoff, hwt) & 255;
    const int MF = resize_sample<int>(P.maskF, P.nBlkXP, voff, vwt, hoff, hwt) & 255;
    int VXB, VYB, VXF, VYF;
    inter_vector(P, P.vxB, P.vyB, x, y, VXB, VYB);
    inter_vector(P, P.vxF, P.vyF, x, y, VXF, VYF);
    const int fx = (x + P.S.hpad) << P.logPel, fy = (y + P.S.vpad) << P.logPel, time256 = P.time256;
    if (method == 0 && time256 == 128) { // :464-485
        const int dstF = finest_sample<T>(left_super, P.S, P.logPel, fx + (VXF >> 1), fy + (VYF >> 1));
        const int dstB = finest_sample<T>(right_super, P.S, P.logPel, fx + (VXB >> 1), fy + (VYB >> 1));
        return (((dstF + dstB) << 8) + (dstB - dstF) * (MF - MB)) >> 9;
    }
    const int dstF = finest_sample<T>(left_super, P.S, P.logPel, fx + ((VXF * time256) >> 8), fy + ((VYF * time256) >> 8));
    const int dstB = finest_sample<T>(right_super, P.S, P.logPel, fx + ((VXB * (256 - time256)) >> 8), fy + ((VYB * (256 - time256)) >> 8));
    if (method == 0) // :486-510
        return (((dstF * (255 - MF) + dstB * MF + 255) >> 8) * (256 - time256) + ((dstB * (255 - MB) + dstF * MB + 255) >> 8) * time256) >> 8;
    if (method == 1) { // :376-418 (64-bit intermediates)
        const long long dF = dstF, dB = dstB;
        const int dstF0 = finest_sample<T>(left_super, P.S, P.logPel, fx, fy), dstB0 = finest_sample<T>(right_super, P.S, P.logPel, fx, fy);
        return (int)((((dF * (255 - MF) + ((MF * (dB * (255 - MB) + MB * dstF0) + 255) >> 8) + 255) >> 8) * (256 - time256) +
                      ((dB * (255 - MB) + ((MB * (dF * (255 - MF) + MF * dstB0) + 255) >> 8) + 255) >> 8) * time256) >>
                     8);
    }
    int VXBB, VYBB, VXFF, VYFF; // :421-461
    inter_vector(P, P.vxBB, P.vyBB, x, y, VXBB, VYBB);
    inter_vector(P, P.vxFF, P.vyFF, x, y, VXFF, VYFF);
    const int dstFF = finest_sample<T>(left_super, P.S, P.logPel, fx + ((VXFF * time256) >> 8), fy + ((VYFF * time256) >> 8));
    const int dstBB = finest_sample<T>(right_super, P.S, P.logPel, fx + ((VXBB * (256 - time256)) >> 8), fy + ((VYBB * (256 - time256)) >> 8));
    const int minfb = dstB < dstF ? dstB : dstF, maxfb = dstB > dstF ? dstB : dstF;
    const int mBB = dstBB < maxfb ? dstBB : maxfb, mFF = dstFF < maxfb ? dstFF : maxfb;
    const int medianBB = minfb > mBB ? minfb : mBB, medianFF = minfb > mFF ? minfb : mFF;
    return (((medianBB * MF + dstF * (255 - MF) + 255) >> 8) * (256 - time256) + ((medianFF * MB + dstB * (255 - MB) + 255) >> 8) * time256) >> 8;
}

// Blend, MaskFun.cpp:349-373
MVB_HD int blend_sample(int src, int ref, int time256) { return (src * (256 - time256) + ref * time256) >> 8; }

// ------------------------------------------------------------------ FlowBlur

struct BlurPlane {
    int plane, W, H, pel, logPel, nBlkX, blur256, prec;
    const short *vxB, *vyB, *vxF, *vyF; // small fields (nBlkX x nBlkY, not padded: MVFlowBlur.c:243-244)
    ResizeTables t;
    SuperPlane S;
};

// RealFlowBlur, MVFlowBlur.c:72-127: the mean of the samples met along the forward and the backward vector
template <typename T>
MVB_HD int flow_blur_sample(const BlurPlane &P, const uint8_t *super_frame, int x, int y) {
    const int voff = P.t.voff[y], vwt = P.t.vwt[y], hoff = P.t.hoff[x], hwt = P.t.hwt[x];
    const int fx = (x + P.S.hpad) << P.logPel, fy = (y + P.S.vpad) << P.logPel;
    int sum = finest_sample<T>(super_frame, P.S, P.logPel, fx, fy), count = 1;
    for (int dir = 0; dir < 2; dir++) {
        const short *vxs = dir ? P.vxB : P.vxF, *vys = dir ? P.vyB : P.vyF;
        const int VX = (short)limit_vector(resize_sample<short>(vxs, P.nBlkX, voff, vwt, hoff, hwt), x, P.W, P.pel);
        const int VY = (short)limit_vector(resize_sample<short>(vys, P.nBlkX, voff, vwt, hoff, hwt), y, P.H, P.pel);
        int vx0 = VX * P.blur256, vy0 = VY * P.blur256;
        const int ax = vx0 < 0 ? -vx0 : vx0, ay = vy0 < 0 ? -vy0 : vy0;
        const int m = ((ax > ay ? ax : ay) / P.prec) >> 8;
        if (m > 0) {
            vx0 /= m;
            vy0 /= m;
            int vx = vx0, vy = vy0;
            for (int i = 0; i < m; i++) {
                sum += finest_sample<T>(super_frame, P.S, P.logPel, fx + (vx >> 8), fy + (vy >> 8));
                vx += vx0;
                vy += vy0;
            }
        }
        count += m;
    }
    return sum / count;
}

// ------------------------------------------------------------------ BlockFPS

// MakeSADMaskTime (MaskFun.cpp:136-160) with gamma = 1 as BlockFPS calls it (MVBlockFPS.c:419-420): entry j
MVB_HD int sad_mask_entry(const int *level0, int nBlkX, int nBlkY, int j, double dSADNormFactor, int nPel, int time256, int stepX, int stepY, int shift) {
    const int time4096X = (256 - time256) * 16 / (stepX * nPel);
    const int time4096Y = (256 - time256) * 16 / (stepY * nPel);
    const int bx = j % nBlkX, by = j / nBlkX;
    const int *v = level0 + 4 * j;
    int bxi = bx - v[0] * time4096X / 4096;
    int byi = by - v[1] * time4096Y / 4096;
    if (bxi < 0 || bxi >= nBlkX || byi < 0 || byi >= nBlkY) {
        bxi = bx;
        byi = by;
    }
    const long long sad = vec_sad(level0 + 4 * (bxi + byi * nBlkX)) >> shift;
    const double l = 255 * (sad * dSADNormFactor); // pow(x, 1.0) is x
    return (int)(unsigned char)((l > 255) ? 255 : l);
}

struct BfpsPlane {
    int plane, W, H, WB, HB, bw, bh, ox, oy, xR, yR, stepX0, stepY0, nBlkX, nBlkY, nBlkXP, pel, logPel, time256, mode, bits;
    int refHpad, refVpad;             // where the reference puts the interior for its pRef / pSrc reads (MVBlockFPS.c:443-452)
    const int *maskB, *maskF, *maskO; // padded small masks (nBlkXP x nBlkYP)
    const short *win;                 // nine overlap windows of bw * bh, or null without overlap
    ResizeTables t;
    SuperPlane S;
};

MVB_HD int bfps_median(int a, int b, int c) { // MEDIAN, MVBlockFPS.c:101-110
    const int mn = a < b ? a : b, mx = a > b ? a : b, m = mx < c ? mx : c;
    return mn > m ? mn : m;
}

// RealResultBlock (MVBlockFPS.c:116-222): the sample block (bx, by) contributes at plane position (x, y)
template <typename T>
MVB_HD int bfps_result(const BfpsPlane &P, const int *l0B, const int *l0F, const uint8_t *right_super, const uint8_t *left_super, int bx, int by, int x, int y) {
    const int i = by * P.nBlkX + bx, t = P.time256;
    const int ix = x - bx * (P.bw - P.ox), iy = y - by * (P.bh - P.oy);
    // mvpGetPointer at the time-scaled vector (C division toward zero for the chroma planes), then (ix, iy) inside that sub-plane
    const int bxB = (bx * P.stepX0 * P.pel + ((l0B[4 * i] * (256 - t)) >> 8)) / P.xR + P.S.hpad * P.pel;
    const int byB = (by * P.stepY0 * P.pel + ((l0B[4 * i + 1] * (256 - t)) >> 8)) / P.yR + P.S.vpad * P.pel;
    const int bxF = (bx * P.stepX0 * P.pel + ((l0F[4 * i] * t) >> 8)) / P.xR + P.S.hpad * P.pel;
    const int byF = (by * P.stepY0 * P.pel + ((l0F[4 * i + 1] * t) >> 8)) / P.yR + P.S.vpad * P.pel;
    const int mcb = finest_sample<T>(right_super, P.S, P.logPel, bxB + (ix << P.logPel), byB + (iy << P.logPel));
    const int mcf = finest_sample<T>(left_super, P.S, P.logPel, bxF + (ix << P.logPel), byF + (iy << P.logPel));
    const int mode = P.mode;
    if (mode == 0)
        return (mcb * t + mcf * (256 - t)) >> 8;
    const long long at = P.S.offset + (long long)(y + P.refVpad) * P.S.pitch + (long long)(x + P.refHpad) * (int)sizeof(T);
    if (mode == 1 || mode == 2) {
        const int ref = *(const T *)(right_super + at), src = *(const T *)(left_super + at);
        if (mode == 1)
            return bfps_median(ref, src, (mcb * t + mcf * (256 - t)) >> 8);
        return bfps_median((ref * t + src * (256 - t)) >> 8, mcb, mcf);
    }
    const int voff = P.t.voff[y], vwt = P.t.vwt[y], hoff = P.t.hoff[x], hwt = P.t.hwt[x];
    if (mode == 5 || mode == 8)
        return (resize_sample<int>(P.maskO, P.nBlkXP, voff, vwt, hoff, hwt) & 255) << (P.bits - 8);
    const int mB = resize_sample<int>(P.maskB, P.nBlkXP, voff, vwt, hoff, hwt) & 255, mF = resize_sample<int>(P.maskF, P.nBlkXP, voff, vwt, hoff, hwt) & 255;
    if (mode == 3 || mode == 6)
        return (((mB * mcf + (255 - mB) * mcb + 255) >> 8) * t + ((mF * mcb + (255 - mF) * mcf + 255) >> 8) * (256 - t)) >> 8;
    const int occ = resize_sample<int>(P.maskO, P.nBlkXP, voff, vwt, hoff, hwt) & 255; // modes 4, 7
    const int ref = *(const T *)(right_super + at), src = *(const T *)(left_super + at);
    const int f = (mF * mcb + (255 - mF) * mcf + 255) >> 8, b = (mB * mcf + (255 - mB) * mcb + 255) >> 8;
    const int avg = (ref * t + src * (256 - t) + 255) >> 8, m = (b * t + f * (256 - t)) >> 8;
    return (avg * occ + m * (255 - occ) + 255) >> 8;
}

// mvblockfpsGetFrame, MVBlockFPS.c:466-606, as a gather: output sample (x, y) of a plane
template <typename T>
MVB_HD int bfps_sample(const BfpsPlane &P, const int *l0B, const int *l0F, const uint8_t *right_super, const uint8_t *left_super, int x, int y) {
    if (x >= P.WB || y >= P.HB) { // right of / below the blocks: time-weighted blend of the two frames
        const long long at = P.S.offset + (long long)(y + P.refVpad) * P.S.pitch + (long long)(x + P.refHpad) * (int)sizeof(T);
        return blend_sample(*(const T *)(left_super + at), *(const T *)(right_super + at), P.time256);
    }
    if (!P.win)
        return bfps_result<T>(P, l0B, l0F, right_super, left_super, x / P.bw, y / P.bh, x, y);
    const int sx = P.bw - P.ox, sy = P.bh - P.oy;
    const int bx1 = x / sx < P.nBlkX - 1 ? x / sx : P.nBlkX - 1, by1 = y / sy < P.nBlkY - 1 ? y / sy : P.nBlkY - 1;
    const int bx0 = (bx1 >= 1 && x - bx1 * sx < P.ox) ? bx1 - 1 : bx1, by0 = (by1 >= 1 && y - by1 * sy < P.oy) ? by1 - 1 : by1;
    unsigned acc = 0; // the reference's uint16 (8-bit clips) / uint32 accumulator: sums commute, the block order does not matter
    for (int by = by0; by <= by1; by++)
        for (int bx = bx0; bx <= bx1; bx++) {
            const int v = bfps_result<T>(P, l0B, l0F, right_super, left_super, bx, by, x, y);
            const int wby = ((by + P.nBlkY - 3) / (P.nBlkY - 2)) * 3, wbx = bx == P.nBlkX - 1 ? 2 : (bx == 0 ? 0 : 1);
            acc += (unsigned)((v * (int)P.win[(wby + wbx) * P.bw * P.bh + (y - by * sy) * P.bw + (x - bx * sx)]) >> 6);
            if (sizeof(T) == 1)
                acc &= 0xffffu;
        }
    const int a = (int)((acc + 16) >> 5), pixelMax = (1 << P.bits) - 1; // ToPixels, Overlap.cpp:335-356
    return a < pixelMax ? a : pixelMax;
}

} // namespace flowdev
} // namespace mvb200
