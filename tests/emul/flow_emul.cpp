// TEST INFRASTRUCTURE -- CPU emulation of the flow.cu kernels.
//
// The kernels of vapoursynth-mvtools_b200/csrc/flow.cu are index wrappers around the per-thread functions of
// csrc/flow_device.h, with parameter blocks built by csrc/flow_setup.h.  This file includes those two headers as
// host C++ and runs the functions "thread by thread" in plain loops (an atomicMax becomes a max), with the same
// frame-level structure as the launchers: scene-change test -> per-block fields -> one pass per plane.  tests/
// compare the result with the oracle bit for bit, which checks the kernels' index arithmetic without a GPU.
// Nothing in the product links, loads or calls this file: the engine has no CPU path.
#include "../../vapoursynth-mvtools_b200/csrc/flow_device.h"
#include "../../vapoursynth-mvtools_b200/csrc/flow_setup.h"

#include <cstring>
#include <vector>

using namespace mvb200::flowdev;

#define EMUL_API extern "C" __attribute__((visibility("default")))

static int usable_of(const uint8_t *blob, int nBlk, long long nSCD1, int nSCD2, int haveRef) { // k_flow_scd
    const int *l0 = blob_level0(blob, nBlk);
    int cnt = 0;
    for (int i = 0; i < nBlk; i++)
        cnt += vec_sad(l0 + 4 * i) > nSCD1 ? 1 : 0;
    return (((const int *)blob)[1] == 1) && !(cnt > nSCD2) && haveRef;
}

EMUL_API int emul_scdetection_frame(const mvb200_scdetection_params *p, const uint8_t *vectors, int32_t *scene_change) {
    *scene_change = !usable_of(vectors, p->vd.nBlkX * p->vd.nBlkY, p->nSCD1, p->nSCD2, 1);
    return 0;
}

EMUL_API int emul_mask_frame(const mvb200_mask_params *mp, const uint8_t *src_luma, ptrdiff_t src_stride, const uint8_t *vectors, uint8_t *const dst[3], const ptrdiff_t dst_stride[3]) {
    const mvb200_analysis_data &vd = mp->vd;
    const int n = vd.nBlkX * vd.nBlkY;
    const int usable = usable_of(vectors, n, mp->nSCD1, mp->nSCD2, 1);
    std::vector<int32_t> tab(tables_ints(mp->nWidthB, mp->nHeightB) + tables_ints(mp->nWidthBUV, mp->nHeightBUV));
    fill_tables(tab.data(), mp->nWidthB, mp->nHeightB, vd.nBlkX, vd.nBlkY);
    fill_tables(tab.data() + tables_ints(mp->nWidthB, mp->nHeightB), mp->nWidthBUV, mp->nHeightBUV, vd.nBlkX, vd.nBlkY);
    const ResizeTables t[2] = {tables_at(tab.data(), mp->nWidthB, mp->nHeightB), tables_at(tab.data() + tables_ints(mp->nWidthB, mp->nHeightB), mp->nWidthBUV, mp->nHeightBUV)};
    std::vector<int> small(n, 0), smallV(n, 0);
    const MaskSmallParams P = make_mask_small_params(mp);
    const int *l0 = blob_level0(vectors, n);
    int *sm = small.data();
    for (int j = 0; j < n; j++) { // k_mask_small
        if (P.kind == 2)
            occlusion_contrib(l0, P.isBackward, P.nBlkX, P.nBlkY, 1.0 / P.fMaskNormFactor, P.fGamma, P.nPel, P.time256, P.stepX, P.stepY, j % P.nBlkX, j / P.nBlkX, [sm](int idx, int val) { sm[idx] = sm[idx] > val ? sm[idx] : val; });
        else
            mask_small_direct(P, l0, j, small.data(), smallV.data());
    }
    for (int plane = 0; plane < 3; plane++) { // k_mask_plane
        const MaskPlane B = make_mask_plane(mp, plane, t[plane ? 1 : 0]);
        for (int y = 0; y < B.H; y++)
            for (int x = 0; x < B.W; x++)
                dst[plane][y * dst_stride[plane] + x] = (uint8_t)mask_sample(B, mp->kind, usable, mp->nSceneChangeValue, vd.nBlkX, small.data(), smallV.data(), src_luma, src_stride, x, y);
    }
    return 0;
}

static void level0_geom(const mvb200_super_params *sp, int plane, mvb200_level_geom *g) {
    mvb200_super_geom geom;
    mvb200_super_geometry(sp, &geom);
    *g = geom.lv[plane][0];
}

EMUL_API int emul_finest_frame(const mvb200_super_params *sp, const uint8_t *const super[3], const ptrdiff_t super_stride[3], uint8_t *const dst[3], const ptrdiff_t dst_stride[3]) {
    const int bps = (sp->bits + 7) / 8, lp = ilog2_pel(sp->pel);
    for (int plane = 0; plane < sp->nplanes; plane++) { // k_finest
        if (plane && !(sp->modeyuv & (MVB200_UPLANE | MVB200_VPLANE)))
            continue;
        mvb200_level_geom g;
        level0_geom(sp, plane, &g);
        const SuperPlane S = make_super_plane(super_stride[plane], 0, g);
        const int W = S.padded_w * sp->pel, H = S.padded_h * sp->pel;
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                if (bps == 1)
                    dst[plane][y * dst_stride[plane] + x] = finest_sample<uint8_t>(super[plane], S, lp, x, y);
                else
                    *((uint16_t *)(dst[plane] + y * dst_stride[plane]) + x) = finest_sample<uint16_t>(super[plane], S, lp, x, y);
            }
    }
    return 0;
}

EMUL_API int emul_flow_frame(const mvb200_flow_params *fp, const mvb200_super_params *sp, const uint8_t *const ref_super[3], const ptrdiff_t ref_stride[3], const uint8_t *vectors, int field_shift, uint8_t *const dst[3],
                             const ptrdiff_t dst_stride[3], int32_t *usable_out) {
    const mvb200_analysis_data &vd = fp->vd;
    const int usable = usable_of(vectors, vd.nBlkX * vd.nBlkY, fp->nSCD1, fp->nSCD2, ref_super != nullptr);
    *usable_out = usable;
    if (!usable)
        return 0;
    const int ns = fp->nBlkXP * fp->nBlkYP, bps = (fp->bits + 7) / 8;
    std::vector<short> xY(ns), yY(ns), xUV(ns), yUV(ns);
    const int *l0 = blob_level0(vectors, vd.nBlkX * vd.nBlkY);
    for (int k = 0; k < ns; k++) { // k_small_vectors
        const SmallVec s = small_vector(l0, vd.nBlkX, vd.nBlkY, k % fp->nBlkXP, k / fp->nBlkXP, field_shift, vd.xRatioUV, vd.yRatioUV);
        xY[k] = s.xY;
        yY[k] = s.yY;
        xUV[k] = s.xUV;
        yUV[k] = s.yUV;
    }
    const size_t n0 = tables_ints(fp->nWidthP, fp->nHeightP), n1 = fp->nplanes > 1 ? tables_ints(fp->nWidthPUV, fp->nHeightPUV) : 0;
    std::vector<int32_t> tab(n0 + n1);
    fill_tables(tab.data(), fp->nWidthP, fp->nHeightP, fp->nBlkXP, fp->nBlkYP);
    if (n1)
        fill_tables(tab.data() + n0, fp->nWidthPUV, fp->nHeightPUV, fp->nBlkXP, fp->nBlkYP);
    const ResizeTables t[2] = {tables_at(tab.data(), fp->nWidthP, fp->nHeightP), tables_at(tab.data() + n0, fp->nWidthPUV, fp->nHeightPUV)};
    for (int plane = 0; plane < fp->nplanes; plane++) {
        mvb200_level_geom g;
        level0_geom(sp, plane, &g);
        const FlowPlane F = make_flow_plane(fp, plane, plane ? xUV.data() : xY.data(), plane ? yUV.data() : yY.data(), t[plane ? 1 : 0], make_super_plane(ref_stride[plane], 0, g));
        auto put = [&](int x, int y, int v) {
            if (bps == 1)
                dst[plane][y * dst_stride[plane] + x] = (uint8_t)v;
            else
                *((uint16_t *)(dst[plane] + y * dst_stride[plane]) + x) = (uint16_t)v;
        };
        if (fp->mode == 0) { // k_flow_fetch
            for (int y = 0; y < F.H; y++)
                for (int x = 0; x < F.W; x++)
                    put(x, y, bps == 1 ? (int)flow_fetch_sample<uint8_t>(F, ref_super[plane], x, y) : (int)flow_fetch_sample<uint16_t>(F, ref_super[plane], x, y));
        } else {
            std::vector<int> owner((size_t)F.W * F.H, 0);
            for (int y = 0; y < F.H; y++) // k_flow_shift_mark (atomicMax)
                for (int x = 0; x < F.W; x++) {
                    const int tgt = flow_shift_target(F, x, y);
                    if (tgt >= 0 && owner[tgt] < y * F.W + x + 1)
                        owner[tgt] = y * F.W + x + 1;
                }
            for (int y = 0; y < F.H; y++) // k_flow_shift_resolve
                for (int x = 0; x < F.W; x++) {
                    const int o = owner[y * F.W + x];
                    if (o == 0) {
                        put(x, y, (1 << fp->bits) - 1);
                        continue;
                    }
                    const int sx = (o - 1) % F.W, sy = (o - 1) / F.W;
                    put(x, y, bps == 1 ? (int)finest_sample<uint8_t>(ref_super[plane], F.S, F.logPel, (sx + F.S.hpad) << F.logPel, (sy + F.S.vpad) << F.logPel)
                                       : (int)finest_sample<uint16_t>(ref_super[plane], F.S, F.logPel, (sx + F.S.hpad) << F.logPel, (sy + F.S.vpad) << F.logPel));
                }
        }
    }
    return 0;
}

EMUL_API int emul_flowinter_frame(const mvb200_flowinter_params *p, const mvb200_super_params *sp, int time256, const uint8_t *const left_super[3], const ptrdiff_t left_stride[3], const uint8_t *const right_super[3],
                                  const ptrdiff_t right_stride[3], const uint8_t *vec_b, const uint8_t *vec_f, const uint8_t *vec_bb, const uint8_t *vec_ff, uint8_t *const dst[3], const ptrdiff_t dst_stride[3], int32_t *usable_out) {
    const mvb200_analysis_data &vd = p->vd;
    const int nb = vd.nBlkX * vd.nBlkY, ns = p->nBlkXP * p->nBlkYP, bps = (p->bits + 7) / 8;
    const uint8_t *blobs[4] = {vec_b, vec_f, vec_bb, vec_ff};
    int flags[4];
    std::vector<short> fld[16];
    for (int i = 0; i < 4; i++) { // k_flow_scd + k_small_vectors per blob
        flags[i] = blobs[i] ? usable_of(blobs[i], nb, p->nSCD1, p->nSCD2, 1) : 0;
        for (int k = 0; k < 4; k++)
            fld[4 * i + k].assign(ns, 0);
        if (!blobs[i])
            continue;
        const int *l0 = blob_level0(blobs[i], nb);
        for (int k = 0; k < ns; k++) {
            const SmallVec s = small_vector(l0, vd.nBlkX, vd.nBlkY, k % p->nBlkXP, k / p->nBlkXP, 0, vd.xRatioUV, vd.yRatioUV);
            fld[4 * i][k] = s.xY;
            fld[4 * i + 1][k] = s.yY;
            fld[4 * i + 2][k] = s.xUV;
            fld[4 * i + 3][k] = s.yUV;
        }
    }
    *usable_out = flags[0] && flags[1];
    if (!*usable_out)
        return 0;
    const int stepX = vd.nBlkSizeX - vd.nOverlapX, stepY = vd.nBlkSizeY - vd.nOverlapY;
    std::vector<int> raw[2] = {std::vector<int>(nb, 0), std::vector<int>(nb, 0)}, mask[2] = {std::vector<int>(ns), std::vector<int>(ns)};
    for (int m = 0; m < 2; m++) { // k_occ_small + k_pad_mask
        int *r = raw[m].data();
        const int *l0 = blob_level0(blobs[m], nb);
        for (int j = 0; j < nb; j++)
            occlusion_contrib(l0, m == 0, vd.nBlkX, vd.nBlkY, p->ml, 1.0, vd.nPel, m == 0 ? 256 - time256 : time256, stepX, stepY, j % vd.nBlkX, j / vd.nBlkX, [r](int idx, int val) { r[idx] = r[idx] > val ? r[idx] : val; });
        for (int k = 0; k < ns; k++)
            mask[m][k] = padded_mask_entry(r, vd.nBlkX, vd.nBlkY, k % p->nBlkXP, k / p->nBlkXP);
    }
    const size_t n0 = tables_ints(p->nWidthP, p->nHeightP), n1 = p->nplanes > 1 ? tables_ints(p->nWidthPUV, p->nHeightPUV) : 0;
    std::vector<int32_t> tab(n0 + n1);
    fill_tables(tab.data(), p->nWidthP, p->nHeightP, p->nBlkXP, p->nBlkYP);
    if (n1)
        fill_tables(tab.data() + n0, p->nWidthPUV, p->nHeightPUV, p->nBlkXP, p->nBlkYP);
    const ResizeTables t[2] = {tables_at(tab.data(), p->nWidthP, p->nHeightP), tables_at(tab.data() + n0, p->nWidthPUV, p->nHeightPUV)};
    const int method = (flags[2] && flags[3]) ? p->mode_primary : p->mode_fallback;
    for (int plane = 0; plane < p->nplanes; plane++) { // k_flow_inter
        const int u = plane ? 2 : 0;
        const InterSmall v = {fld[0 + u].data(), fld[1 + u].data(), fld[4 + u].data(), fld[5 + u].data(), fld[8 + u].data(), fld[9 + u].data(), fld[12 + u].data(), fld[13 + u].data()};
        mvb200_level_geom g;
        level0_geom(sp, plane, &g);
        // (both frames share one layout in the engine; the emulation asserts equal strides instead)
        if (left_stride[plane] != right_stride[plane])
            return 1;
        const InterPlane P = make_inter_plane(p, time256, plane, v, mask[0].data(), mask[1].data(), t[plane ? 1 : 0], make_super_plane(left_stride[plane], 0, g));
        for (int y = 0; y < P.H; y++)
            for (int x = 0; x < P.W; x++) {
                if (bps == 1)
                    dst[plane][y * dst_stride[plane] + x] = (uint8_t)flow_inter_sample<uint8_t>(P, method, right_super[plane], left_super[plane], x, y);
                else
                    *((uint16_t *)(dst[plane] + y * dst_stride[plane]) + x) = (uint16_t)flow_inter_sample<uint16_t>(P, method, right_super[plane], left_super[plane], x, y);
            }
    }
    return 0;
}

EMUL_API int emul_blend_frame(const mvb200_super_params *sp, int time256, const uint8_t *const src[3], const ptrdiff_t src_stride[3], const uint8_t *const ref[3], const ptrdiff_t ref_stride[3], uint8_t *const dst[3],
                              const ptrdiff_t dst_stride[3]) {
    const int bps = (sp->bits + 7) / 8;
    for (int pl = 0; pl < sp->nplanes; pl++) { // k_blend
        const int w = pl ? sp->width >> sp->ssw : sp->width, h = pl ? sp->height >> sp->ssh : sp->height;
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                if (bps == 1)
                    dst[pl][y * dst_stride[pl] + x] = (uint8_t)blend_sample(src[pl][y * src_stride[pl] + x], ref[pl][y * ref_stride[pl] + x], time256);
                else
                    *((uint16_t *)(dst[pl] + y * dst_stride[pl]) + x) = (uint16_t)blend_sample(*((const uint16_t *)(src[pl] + y * src_stride[pl]) + x), *((const uint16_t *)(ref[pl] + y * ref_stride[pl]) + x), time256);
            }
    }
    return 0;
}

EMUL_API int emul_flowblur_frame(const mvb200_flowblur_params *p, const mvb200_super_params *sp, const uint8_t *const super[3], const ptrdiff_t super_stride[3], const uint8_t *vec_b, const uint8_t *vec_f, uint8_t *const dst[3],
                                 const ptrdiff_t dst_stride[3], int32_t *usable_out) {
    const mvb200_analysis_data &vd = p->vd;
    const int nb = vd.nBlkX * vd.nBlkY, bps = (p->bits + 7) / 8;
    const uint8_t *blobs[2] = {vec_b, vec_f};
    *usable_out = vec_b && vec_f && usable_of(vec_b, nb, p->nSCD1, p->nSCD2, 1) && usable_of(vec_f, nb, p->nSCD1, p->nSCD2, 1);
    if (!*usable_out)
        return 0;
    std::vector<short> fld[8];
    for (int i = 0; i < 2; i++) { // k_small_vectors (no padded field here)
        const int *l0 = blob_level0(blobs[i], nb);
        for (int k = 0; k < 4; k++)
            fld[4 * i + k].assign(nb, 0);
        for (int k = 0; k < nb; k++) {
            const SmallVec s = small_vector(l0, vd.nBlkX, vd.nBlkY, k % vd.nBlkX, k / vd.nBlkX, 0, vd.xRatioUV, vd.yRatioUV);
            fld[4 * i][k] = s.xY;
            fld[4 * i + 1][k] = s.yY;
            fld[4 * i + 2][k] = s.xUV;
            fld[4 * i + 3][k] = s.yUV;
        }
    }
    const size_t n0 = tables_ints(vd.nWidth, vd.nHeight), n1 = p->nplanes > 1 ? tables_ints(p->nWidthUV, p->nHeightUV) : 0;
    std::vector<int32_t> tab(n0 + n1);
    fill_tables(tab.data(), vd.nWidth, vd.nHeight, vd.nBlkX, vd.nBlkY);
    if (n1)
        fill_tables(tab.data() + n0, p->nWidthUV, p->nHeightUV, vd.nBlkX, vd.nBlkY);
    const ResizeTables t[2] = {tables_at(tab.data(), vd.nWidth, vd.nHeight), tables_at(tab.data() + n0, p->nWidthUV, p->nHeightUV)};
    for (int plane = 0; plane < p->nplanes; plane++) { // k_flow_blur
        const int u = plane ? 2 : 0;
        mvb200_level_geom g;
        level0_geom(sp, plane, &g);
        const BlurPlane P = make_blur_plane(p, plane, fld[0 + u].data(), fld[1 + u].data(), fld[4 + u].data(), fld[5 + u].data(), t[plane ? 1 : 0], make_super_plane(super_stride[plane], 0, g));
        for (int y = 0; y < P.H; y++)
            for (int x = 0; x < P.W; x++) {
                if (bps == 1)
                    dst[plane][y * dst_stride[plane] + x] = (uint8_t)flow_blur_sample<uint8_t>(P, super[plane], x, y);
                else
                    *((uint16_t *)(dst[plane] + y * dst_stride[plane]) + x) = (uint16_t)flow_blur_sample<uint16_t>(P, super[plane], x, y);
            }
    }
    return 0;
}

EMUL_API int emul_blockfps_frame(const mvb200_blockfps_params *p, const mvb200_super_params *sp, int time256, const uint8_t *const left_super[3], const ptrdiff_t left_stride[3], const uint8_t *const right_super[3],
                                 const ptrdiff_t right_stride[3], const uint8_t *vec_b, const uint8_t *vec_f, uint8_t *const dst[3], const ptrdiff_t dst_stride[3], int32_t *usable_out) {
    const mvb200_analysis_data &vd = p->vd;
    const int nb = vd.nBlkX * vd.nBlkY, ns = p->nBlkXP * p->nBlkYP, bps = (p->bits + 7) / 8;
    const int planes = (p->super_modeyuv & (MVB200_UPLANE | MVB200_VPLANE)) ? 3 : 1;
    *usable_out = vec_b && vec_f && usable_of(vec_b, nb, p->nSCD1, p->nSCD2, 1) && usable_of(vec_f, nb, p->nSCD1, p->nSCD2, 1);
    if (!*usable_out)
        return 0;
    const int *l0B = blob_level0(vec_b, nb), *l0F = blob_level0(vec_f, nb);
    std::vector<int> raw[2] = {std::vector<int>(nb, 0), std::vector<int>(nb, 0)}, mask[3] = {std::vector<int>(ns, 0), std::vector<int>(ns, 0), std::vector<int>(ns, 0)};
    if (p->mode >= 3) {
        const int stepX = vd.nBlkSizeX - vd.nOverlapX, stepY = vd.nBlkSizeY - vd.nOverlapY;
        for (int m = 0; m < 2; m++) { // 0 = backward, 1 = forward
            int *r = raw[m].data();
            const int *l0 = m == 0 ? l0B : l0F;
            const int t = m == 0 ? 256 - time256 : time256;
            for (int j = 0; j < nb; j++) {
                if (p->mode <= 5) // k_occ_small
                    occlusion_contrib(l0, m == 0, vd.nBlkX, vd.nBlkY, p->ml, 1.0, vd.nPel, t, stepX, stepY, j % vd.nBlkX, j / vd.nBlkX, [r](int idx, int val) { r[idx] = r[idx] > val ? r[idx] : val; });
                else // k_sad_small
                    r[j] = sad_mask_entry(l0, vd.nBlkX, vd.nBlkY, j, 4.0 / (p->ml * vd.nBlkSizeX * vd.nBlkSizeY), vd.nPel, t, stepX, stepY, p->bits - 8);
            }
            for (int k = 0; k < ns; k++) // k_pad_mask
                mask[m][k] = padded_mask_entry(r, vd.nBlkX, vd.nBlkY, k % p->nBlkXP, k / p->nBlkXP);
        }
        for (int k = 0; k < ns; k++) // k_mult_masks
            mask[2][k] = (mask[1][k] * mask[0][k]) / 255;
    }
    const size_t n0 = tables_ints(p->nWidthP, p->nHeightP), n1 = planes > 1 ? tables_ints(p->nWidthPUV, p->nHeightPUV) : 0;
    std::vector<int32_t> tab(n0 + n1);
    fill_tables(tab.data(), p->nWidthP, p->nHeightP, p->nBlkXP, p->nBlkYP);
    if (n1)
        fill_tables(tab.data() + n0, p->nWidthPUV, p->nHeightPUV, p->nBlkXP, p->nBlkYP);
    const ResizeTables t[2] = {tables_at(tab.data(), p->nWidthP, p->nHeightP), tables_at(tab.data() + n0, p->nWidthPUV, p->nHeightPUV)};
    std::vector<int16_t> win[2];
    win[0].resize((size_t)9 * vd.nBlkSizeX * vd.nBlkSizeY);
    mvb200_overlap_windows(vd.nBlkSizeX, vd.nBlkSizeY, vd.nOverlapX, vd.nOverlapY, win[0].data());
    if (planes > 1) {
        win[1].resize((size_t)9 * (vd.nBlkSizeX / vd.xRatioUV) * (vd.nBlkSizeY / vd.yRatioUV));
        mvb200_overlap_windows(vd.nBlkSizeX / vd.xRatioUV, vd.nBlkSizeY / vd.yRatioUV, vd.nOverlapX / vd.xRatioUV, vd.nOverlapY / vd.yRatioUV, win[1].data());
    }
    for (int plane = 0; plane < planes; plane++) { // k_blockfps
        mvb200_level_geom g;
        level0_geom(sp, plane, &g);
        if (left_stride[plane] != right_stride[plane])
            return 1;
        const BfpsPlane P = make_bfps_plane(p, time256, plane, mask[0].data(), mask[1].data(), mask[2].data(), win[plane ? 1 : 0].data(), t[plane ? 1 : 0], make_super_plane(left_stride[plane], 0, g));
        for (int y = 0; y < P.H; y++)
            for (int x = 0; x < P.W; x++) {
                if (bps == 1)
                    dst[plane][y * dst_stride[plane] + x] = (uint8_t)bfps_sample<uint8_t>(P, l0B, l0F, right_super[plane], left_super[plane], x, y);
                else
                    *((uint16_t *)(dst[plane] + y * dst_stride[plane]) + x) = (uint16_t)bfps_sample<uint16_t>(P, l0B, l0F, right_super[plane], left_super[plane], x, y);
            }
    }
    return 0;
}
