// Host-side construction of the parameter blocks the flow.cu kernels take (flow_device.h), from the C ABI's
// parameter structs.  Shared by the launchers in flow.cu and by the CPU emulation test (tests/emul/flow_emul.cpp),
// so that the test exercises the same set-up code as the product.
#pragma once

#include "../../include/mvtools_b200.h"
#include "flow_device.h"

#include <cstring>

namespace mvb200 {
namespace flowdev {

inline int ilog2_pel(int pel) { return pel == 4 ? 2 : (pel == 2 ? 1 : 0); }

// level 0 of one plane of a super frame: `pitch` / `offset` of the plane inside the frame buffer
inline SuperPlane make_super_plane(long long pitch, long long offset, const mvb200_level_geom &g) {
    SuperPlane S;
    S.pitch = pitch;
    S.offset = offset;
    S.sub_stride = pitch * g.padded_h;
    S.hpad = g.hpad;
    S.vpad = g.vpad;
    S.padded_w = g.padded_w;
    S.padded_h = g.padded_h;
    return S;
}

// where the four tables of an upsizer (dst_w x dst_h) live inside one int array: [hoff dw][hwt dw][voff dh][vwt dh]
inline ResizeTables tables_at(const int *base, int dw, int dh) {
    ResizeTables t;
    t.hoff = base;
    t.hwt = base + dw;
    t.voff = base + 2 * dw;
    t.vwt = base + 2 * dw + dh;
    return t;
}
inline size_t tables_ints(int dw, int dh) { return (size_t)dw * 2 + (size_t)dh * 2; }
// fills `host` (tables_ints(dw, dh) ints) for an upsizer from (sw x sh) blocks
inline void fill_tables(int32_t *host, int dw, int dh, int sw, int sh) {
    mvb200_simple_resize_tables(dw, sw, host, host + dw);
    mvb200_simple_resize_tables(dh, sh, host + 2 * dw, host + 2 * dw + dh);
}

inline FlowPlane make_flow_plane(const mvb200_flow_params *fp, int plane, const short *vxs, const short *vys, const ResizeTables &t, const SuperPlane &S) {
    FlowPlane F;
    memset(&F, 0, sizeof(F));
    F.plane = plane;
    F.W = plane ? fp->nWidthUV : fp->vd.nWidth;
    F.H = plane ? fp->nHeightUV : fp->vd.nHeight;
    F.limitW = F.W; // simpleInit(..., nWidth, nHeight, ...) / (..., nWidthUV, nHeightUV, ...), MVFlow.cpp:561-563
    F.limitH = F.H;
    F.time256 = fp->time256;
    F.pel = fp->vd.nPel;
    F.logPel = ilog2_pel(fp->vd.nPel);
    F.mode = fp->mode;
    F.nBlkXP = fp->nBlkXP;
    F.vxs = vxs;
    F.vys = vys;
    F.t = t;
    F.S = S;
    return F;
}

inline MaskSmallParams make_mask_small_params(const mvb200_mask_params *mp) {
    const mvb200_analysis_data &vd = mp->vd;
    MaskSmallParams P;
    memset(&P, 0, sizeof(P));
    P.kind = mp->kind;
    P.nBlkX = vd.nBlkX;
    P.nBlkY = vd.nBlkY;
    P.nPel = vd.nPel;
    P.isBackward = vd.isBackward;
    P.time256 = mp->time256;
    P.stepX = vd.nBlkSizeX - vd.nOverlapX;
    P.stepY = vd.nBlkSizeY - vd.nOverlapY;
    P.blkSizeXY = vd.nBlkSizeX * vd.nBlkSizeY;
    P.fGamma = mp->fGamma;
    P.fMaskNormFactor = mp->fMaskNormFactor;
    P.fMaskNormFactor2 = mp->fMaskNormFactor2;
    P.fHalfGamma = mp->fHalfGamma;
    return P;
}

inline MaskPlane make_mask_plane(const mvb200_mask_params *mp, int plane, const ResizeTables &t) {
    MaskPlane B;
    B.plane = plane;
    B.W = plane ? mp->nWidthUV : mp->nWidth;
    B.H = plane ? mp->nHeightUV : mp->nHeight;
    B.WB = plane ? mp->nWidthBUV : mp->nWidthB;
    B.HB = plane ? mp->nHeightBUV : mp->nHeightB;
    B.t = t;
    return B;
}

struct InterSmall { // the small fields of one plane type
    const short *vxB, *vyB, *vxF, *vyF, *vxBB, *vyBB, *vxFF, *vyFF;
};

inline InterPlane make_inter_plane(const mvb200_flowinter_params *p, int time256, int plane, const InterSmall &v, const int *maskB, const int *maskF, const ResizeTables &t, const SuperPlane &S) {
    InterPlane P;
    memset(&P, 0, sizeof(P));
    P.plane = plane;
    P.W = plane ? p->nWidthUV : p->vd.nWidth;
    P.H = plane ? p->nHeightUV : p->vd.nHeight;
    P.limitW = P.W; // simpleInit, MVFlowInter.c:662-664
    P.limitH = P.H;
    P.pel = p->vd.nPel;
    P.logPel = ilog2_pel(p->vd.nPel);
    P.nBlkXP = p->nBlkXP;
    P.time256 = time256;
    P.vxB = v.vxB;
    P.vyB = v.vyB;
    P.vxF = v.vxF;
    P.vyF = v.vyF;
    P.vxBB = v.vxBB;
    P.vyBB = v.vyBB;
    P.vxFF = v.vxFF;
    P.vyFF = v.vyFF;
    P.maskB = maskB;
    P.maskF = maskF;
    P.t = t;
    P.S = S;
    return P;
}

inline BlurPlane make_blur_plane(const mvb200_flowblur_params *p, int plane, const short *vxB, const short *vyB, const short *vxF, const short *vyF, const ResizeTables &t, const SuperPlane &S) {
    BlurPlane P;
    memset(&P, 0, sizeof(P));
    P.plane = plane;
    P.W = plane ? p->nWidthUV : p->vd.nWidth; // also the resizer's limits (simpleInit, MVFlowBlur.c:531-533)
    P.H = plane ? p->nHeightUV : p->vd.nHeight;
    P.pel = p->vd.nPel;
    P.logPel = ilog2_pel(p->vd.nPel);
    P.nBlkX = p->vd.nBlkX;
    P.blur256 = p->blur256;
    P.prec = p->prec;
    P.vxB = vxB;
    P.vyB = vyB;
    P.vxF = vxF;
    P.vyF = vyF;
    P.t = t;
    P.S = S;
    return P;
}

inline BfpsPlane make_bfps_plane(const mvb200_blockfps_params *p, int time256, int plane, const int *maskB, const int *maskF, const int *maskO, const short *win, const ResizeTables &t, const SuperPlane &S) {
    const mvb200_analysis_data &vd = p->vd;
    BfpsPlane P;
    memset(&P, 0, sizeof(P));
    P.plane = plane;
    P.xR = plane ? vd.xRatioUV : 1;
    P.yR = plane ? vd.yRatioUV : 1;
    P.W = plane ? p->nWidthUV : vd.nWidth;
    P.H = plane ? p->nHeightUV : vd.nHeight;
    P.bw = vd.nBlkSizeX / P.xR;
    P.bh = vd.nBlkSizeY / P.yR;
    P.ox = vd.nOverlapX / P.xR;
    P.oy = vd.nOverlapY / P.yR;
    P.WB = (vd.nBlkX * (vd.nBlkSizeX - vd.nOverlapX) + vd.nOverlapX) / P.xR; // nWidth_B, MVBlockFPS.c:338-339
    P.HB = (vd.nBlkY * (vd.nBlkSizeY - vd.nOverlapY) + vd.nOverlapY) / P.yR;
    P.stepX0 = vd.nBlkSizeX - vd.nOverlapX;
    P.stepY0 = vd.nBlkSizeY - vd.nOverlapY;
    P.nBlkX = vd.nBlkX;
    P.nBlkY = vd.nBlkY;
    P.nBlkXP = p->nBlkXP;
    P.pel = vd.nPel;
    P.logPel = ilog2_pel(vd.nPel);
    P.time256 = time256;
    P.mode = p->mode;
    P.bits = p->bits;
    P.refHpad = plane ? p->super_hpad >> 1 : p->super_hpad; // (sic: >> 1 whatever the subsampling, :447-451)
    P.refVpad = plane ? p->super_vpad >> 1 : p->super_vpad;
    P.maskB = maskB;
    P.maskF = maskF;
    P.maskO = maskO;
    P.win = (vd.nOverlapX || vd.nOverlapY) ? win : nullptr;
    P.t = t;
    P.S = S;
    return P;
}

} // namespace flowdev
} // namespace mvb200
