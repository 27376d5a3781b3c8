// Host-side parameter derivation for the mv.* hot path: what the reference's
// *Create functions compute before any frame is processed.  Pure C++ (no CUDA):
// usable -- and unit-tested -- without a GPU.  Error texts are the reference's.
//
//   mvb200_super_init     <- mvsuperCreate     MVSuper.c:140-264
//   mvb200_super_geometry <- mvgofInit/Update  MVFrame.cpp:1852-1903, PlaneSuperOffset :1229-1247
//   mvb200_analyse_init   <- mvanalyseCreate   MVAnalyse.c:267-624, gopInit GroupOfPlanes.c:25-56
//   mvb200_degrain_init   <- mvdegrainCreate   MVDegrains.cpp:512-782, scaleThSCD MVAnalysisData.c:7-31
//   mvb200_scdetection_init <- mvscdetectionCreate MVSCDetection.c:89-131
//   mvb200_mask_init      <- mvmaskCreate      MVMask.c:221-344
//   mvb200_flow_init      <- mvflowCreate      MVFlow.cpp:395-575
//   mvb200_simple_resize_tables <- InitTables  SimpleResize.cpp:26-56
#include "engine.h"

#include <climits>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>

namespace mvb200 {

static thread_local std::string t_error;

#define GIVEN(x) ((x) != MVB200_UNSET)
#define GIVEN64(x) ((x) != MVB200_UNSET64)

int set_error(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    t_error = buf;
    return 1;
}

// PlaneHeightLuma / PlaneWidthLuma, MVFrame.cpp:1209-1226
int plane_dim_luma(int src, int level, int ratio, int pad) {
    int d = src;
    for (int i = 1; i <= level; i++)
        d = pad >= ratio ? ((d / ratio + 1) / 2) * ratio : ((d / ratio) / 2) * ratio;
    return d;
}

// PlaneSuperOffset, MVFrame.cpp:1229-1247, in rows instead of bytes
static int plane_super_row(int chroma, int src_height, int level, int pel, int vpad, int yRatioUV) {
    if (level == 0)
        return 0;
    int rows = pel * pel * (src_height + vpad * 2);
    for (int i = 1; i < level; i++) {
        int height = chroma ? plane_dim_luma(src_height * yRatioUV, i, yRatioUV, vpad * yRatioUV) / yRatioUV
                            : plane_dim_luma(src_height, i, yRatioUV, vpad);
        rows += height + vpad * 2;
    }
    return rows;
}

// scaleThSCD, MVAnalysisData.c:7-31
static int scale_thscd(int64_t *thscd1, int *thscd2, const mvb200_analysis_data *ad, const char *filter) {
    const int maxSAD = 8 * 8 * 255;
    if (*thscd1 > maxSAD)
        return set_error("%s: thscd1 can be at most %d.", filter, maxSAD);
    *thscd1 = *thscd1 * (ad->nBlkSizeX * ad->nBlkSizeY) / 64;
    if (ad->nMotionFlags & 0x08 /* MOTION_USE_CHROMA_MOTION */)
        *thscd1 += *thscd1 / (ad->xRatioUV * ad->yRatioUV) * 2;
    int pixelMax = (1 << ad->bitsPerSample) - 1;
    *thscd1 = (int64_t)((double)*thscd1 * pixelMax / 255.0 + 0.5);
    *thscd2 = *thscd2 * ad->nBlkX * ad->nBlkY / 256;
    return 0;
}

// adataCheckSimilarity, MVAnalysisData.c:68-98 (the LAST failing check wins, as in the reference)
static int check_similarity(const mvb200_analysis_data *a, const mvb200_analysis_data *b, const char *filter, const char *n1, const char *n2) {
    const char *what = nullptr;
    if (a->nWidth != b->nWidth) what = "widths";
    if (a->nHeight != b->nHeight) what = "heights";
    if (a->nBlkSizeX != b->nBlkSizeX || a->nBlkSizeY != b->nBlkSizeY) what = "block sizes";
    if (a->nPel != b->nPel) what = "pel precision";
    if (a->nOverlapX != b->nOverlapX || a->nOverlapY != b->nOverlapY) what = "overlap";
    if (a->xRatioUV != b->xRatioUV) what = "horizontal subsampling";
    if (a->yRatioUV != b->yRatioUV) what = "vertical subsampling";
    if (a->bitsPerSample != b->bitsPerSample) what = "bit depths";
    if (!what)
        return 0;
    return set_error("%s: %s and %s have different %s.", filter, n1, n2, what);
}

} // namespace mvb200

using namespace mvb200;

extern "C" {

const char *mvb200_last_error(void) { return t_error.c_str(); }

const char *mvb200_version(void) { return "mvtools_b200 0.1 (mv.Super/Analyse/DegrainN on sm_100a; MVTools v24 semantics)"; }

int mvb200_super_init(mvb200_super_params *p) {
    // defaults, MVSuper.c:149-177
    if (p->hpad == MVB200_UNSET) p->hpad = 16;
    if (p->vpad == MVB200_UNSET) p->vpad = 16;
    if (p->pel == MVB200_UNSET) p->pel = 2;
    if (p->levels == MVB200_UNSET) p->levels = 0;
    if (p->chroma == MVB200_UNSET) p->chroma = 1;
    p->chroma = !!p->chroma;
    if (p->sharp == MVB200_UNSET) p->sharp = 2;
    if (p->rfilter == MVB200_UNSET) p->rfilter = 2;
    if (p->pel != 1 && p->pel != 2 && p->pel != 4)
        return set_error("Super: pel must be 1, 2, or 4.");
    if (p->sharp < 0 || p->sharp > 2)
        return set_error("Super: sharp must be between 0 and 2 (inclusive).");
    if (p->rfilter < 0 || p->rfilter > 4)
        return set_error("Super: rfilter must be between 0 and 4 (inclusive).");
    if (p->width <= 0 || p->height <= 0 || p->bits > 16 || p->bits < 8 || p->ssw > 1 || p->ssh > 1 || p->ssw < 0 || p->ssh < 0)
        return set_error("Super: input clip must be GRAY, 420, 422, 440, or 444, up to 16 bits, with constant dimensions.");
    if (p->gray) {
        p->chroma = 0;
        p->ssw = p->ssh = 0;
    }
    p->nplanes = p->gray ? 1 : 3;
    p->modeyuv = p->chroma ? MVB200_YUVPLANES : MVB200_YPLANE;
    p->xRatioUV = 1 << p->ssw;
    p->yRatioUV = 1 << p->ssh;
    int nLevelsMax = 0; // MVSuper.c:220-227
    while (plane_dim_luma(p->height, nLevelsMax, p->yRatioUV, p->vpad) >= p->yRatioUV * 2 &&
           plane_dim_luma(p->width, nLevelsMax, p->xRatioUV, p->hpad) >= p->xRatioUV * 2)
        nLevelsMax++;
    if (p->levels <= 0 || p->levels > nLevelsMax)
        p->levels = nLevelsMax;
    if (p->levels > MVB200_MAX_LEVELS)
        return set_error("Super: too many levels.");
    // MVSuper.c:254-262
    p->super_width = p->width + 2 * p->hpad;
    p->super_height = plane_super_row(0, p->height, p->levels, p->pel, p->vpad, p->yRatioUV);
    if (p->yRatioUV == 2 && (p->super_height & 1))
        p->super_height++;
    if (p->xRatioUV == 2 && (p->super_width & 1))
        p->super_width++;
    return 0;
}

// the pelclip part of mvsuperCreate, MVSuper.c:229-253
int mvb200_super_pelclip_check(const mvb200_super_params *p, int same_format, int pel_width, int pel_height, int *use_pelclip, int *pel_padded) {
    *use_pelclip = 0;
    *pel_padded = 0;
    if (!same_format)
        return set_error("Super: pelclip must have the same format as the input clip, and it must have constant dimensions.");
    if (p->pel >= 2) {
        if (pel_width == p->width * p->pel && pel_height == p->height * p->pel) {
            *use_pelclip = 1;
        } else if (pel_width == (p->width + p->hpad * 2) * p->pel && pel_height == (p->height + p->vpad * 2) * p->pel) {
            *use_pelclip = 1;
            *pel_padded = 1;
        } else {
            return set_error("Super: pelclip's dimensions must be multiples of the input clip's dimensions.");
        }
    }
    return 0;
}

void mvb200_super_geometry(const mvb200_super_params *p, mvb200_super_geom *g) {
    memset(g, 0, sizeof(*g));
    g->nplanes = p->nplanes;
    g->levels = p->levels;
    for (int plane = 0; plane < p->nplanes; plane++) {
        g->plane_w[plane] = plane ? p->super_width >> p->ssw : p->super_width;
        g->plane_h[plane] = plane ? p->super_height >> p->ssh : p->super_height;
        int w0 = plane ? p->width / p->xRatioUV : p->width;
        int h0 = plane ? p->height / p->yRatioUV : p->height;
        int hp = plane ? p->hpad / p->xRatioUV : p->hpad;
        int vp = plane ? p->vpad / p->yRatioUV : p->vpad;
        for (int l = 0; l < p->levels; l++) {
            mvb200_level_geom *lg = &g->lv[plane][l];
            if (l == 0) {
                lg->width = w0;
                lg->height = h0;
            } else {
                int wl = plane_dim_luma(p->width, l, p->xRatioUV, p->hpad);
                int hl = plane_dim_luma(p->height, l, p->yRatioUV, p->vpad);
                lg->width = plane ? wl / p->xRatioUV : wl;
                lg->height = plane ? hl / p->yRatioUV : hl;
            }
            lg->hpad = hp;
            lg->vpad = vp;
            lg->padded_w = lg->width + 2 * hp;
            lg->padded_h = lg->height + 2 * vp;
            lg->pel = l == 0 ? p->pel : 1;
            lg->row0 = plane_super_row(plane != 0, h0, l, p->pel, vp, p->yRatioUV);
        }
    }
}

int mvb200_analyse_init(const mvb200_analyse_args *a, const mvb200_super_params *sp, int num_frames, mvb200_analyse_params *d) {
    memset(d, 0, sizeof(*d));
    mvb200_analysis_data &ad = d->ad;
    // MVAnalyse.c:275-371
    ad.nBlkSizeX = GIVEN(a->blksize) ? a->blksize : 8;
    ad.nBlkSizeY = GIVEN(a->blksizev) ? a->blksizev : ad.nBlkSizeX;
    int levels = GIVEN(a->levels) ? a->levels : 0;
    d->search = GIVEN(a->search) ? a->search : 4;
    d->search_coarse = GIVEN(a->search_coarse) ? a->search_coarse : 3;
    int searchparam = GIVEN(a->searchparam) ? a->searchparam : 2;
    d->nPelSearch = GIVEN(a->pelsearch) ? a->pelsearch : 0;
    ad.isBackward = GIVEN(a->isb) && a->isb != 0;
    d->chroma = GIVEN(a->chroma) ? !!a->chroma : 1;
    ad.nDeltaFrame = GIVEN(a->delta) ? a->delta : 1;
    int truemotion = GIVEN(a->truemotion) ? !!a->truemotion : 1;
    d->nLambda = GIVEN(a->lambda) ? a->lambda : (truemotion ? (1000 * ad.nBlkSizeX * ad.nBlkSizeY / 64) : 0);
    d->lsad = GIVEN(a->lsad) ? a->lsad : (truemotion ? 1200 : 400);
    d->plevel = GIVEN(a->plevel) ? a->plevel : (truemotion ? 1 : 0);
    d->global = GIVEN(a->global) ? !!a->global : (truemotion ? 1 : 0);
    d->pnew = GIVEN(a->pnew) ? a->pnew : (truemotion ? 50 : 0);
    d->pzero = GIVEN(a->pzero) ? a->pzero : d->pnew;
    d->pglobal = GIVEN(a->pglobal) ? a->pglobal : 0;
    ad.nOverlapX = GIVEN(a->overlap) ? a->overlap : 0;
    ad.nOverlapY = GIVEN(a->overlapv) ? a->overlapv : ad.nOverlapX;
    d->dctmode = GIVEN(a->dct) ? a->dct : 0;
    d->divide = GIVEN(a->divide) ? a->divide : 0;
    d->badSAD = GIVEN(a->badsad) ? a->badsad : 10000;
    d->badrange = GIVEN(a->badrange) ? a->badrange : 24;
    int opt = GIVEN(a->opt) ? !!a->opt : 1;
    d->meander = GIVEN(a->meander) ? !!a->meander : 1;
    d->tryMany = GIVEN(a->trymany) && a->trymany != 0;
    d->fields = GIVEN(a->fields) && a->fields != 0;
    d->tff = GIVEN(a->tff) && a->tff != 0;
    d->tff_exists = GIVEN(a->tff);

    // MVAnalyse.c:373-450
    if (d->search < 0 || d->search > 7) return set_error("Analyse: search must be between 0 and 7 (inclusive).");
    if (d->search_coarse < 0 || d->search_coarse > 7) return set_error("Analyse: search_coarse must be between 0 and 7 (inclusive).");
    if (d->dctmode < 0 || d->dctmode > 10) return set_error("Analyse: dct must be between 0 and 10 (inclusive).");
    if (d->dctmode >= 5 && ad.nBlkSizeX == 16 && ad.nBlkSizeY == 2) return set_error("Analyse: dct 5..10 cannot work with 16x2 blocks.");
    if (d->divide < 0 || d->divide > 2) return set_error("Analyse: divide must be between 0 and 2 (inclusive).");
    {
        static const int ok[12][2] = {{4, 4}, {8, 4}, {8, 8}, {16, 2}, {16, 8}, {16, 16}, {32, 16}, {32, 32}, {64, 32}, {64, 64}, {128, 64}, {128, 128}};
        bool found = false;
        for (auto &s : ok)
            found |= (s[0] == ad.nBlkSizeX && s[1] == ad.nBlkSizeY);
        if (!found)
            return set_error("Analyse: the block size must be 4x4, 8x4, 8x8, 16x2, 16x8, 16x16, 32x16, 32x32, 64x32, 64x64, 128x64, or 128x128.");
    }
    if (d->plevel < 0 || d->plevel > 2) return set_error("Analyse: plevel must be between 0 and 2 (inclusive).");
    if (d->pnew < 0 || d->pnew > 256) return set_error("Analyse: pnew must be between 0 and 256 (inclusive).");
    if (d->pzero < 0 || d->pzero > 256) return set_error("Analyse: pzero must be between 0 and 256 (inclusive).");
    if (d->pglobal < 0 || d->pglobal > 256) return set_error("Analyse: pglobal must be between 0 and 256 (inclusive).");
    if (ad.nOverlapX < 0 || ad.nOverlapX > ad.nBlkSizeX / 2 || ad.nOverlapY < 0 || ad.nOverlapY > ad.nBlkSizeY / 2)
        return set_error("Analyse: overlap must be at most half of blksize, overlapv must be at most half of blksizev, and they both need to be at least 0.");
    if (d->divide && (ad.nBlkSizeX < 8 || ad.nBlkSizeY < 8))
        return set_error("Analyse: blksize and blksizev must be at least 8 when divide=True.");
    if (d->search == 1)
        d->nSearchParam = searchparam < 0 ? 0 : searchparam;
    else
        d->nSearchParam = searchparam < 1 ? 1 : searchparam;

    // MVAnalyse.c:459-483
    if (sp->bits > 16 || sp->ssw > 1 || sp->ssh > 1)
        return set_error("Analyse: Input clip must be GRAY, 420, 422, 440, or 444, up to 16 bits, with constant format and dimensions.");
    if (sp->gray)
        d->chroma = 0;
    int nModeYUV = d->chroma ? MVB200_YUVPLANES : MVB200_YPLANE;
    ad.bitsPerSample = sp->bits;
    int pixelMax = (1 << sp->bits) - 1;
    d->lsad = (int)((double)d->lsad * pixelMax / 255.0 + 0.5);
    d->badSAD = (int)((double)d->badSAD * pixelMax / 255.0 + 0.5);
    d->nLambda = (int)((double)d->nLambda * pixelMax / 255.0 + 0.5);
    d->lsad = (int)((int64_t)d->lsad * (ad.nBlkSizeX * ad.nBlkSizeY) / 64);
    d->badSAD = d->badSAD * (ad.nBlkSizeX * ad.nBlkSizeY) / 64;

    ad.nMotionFlags = 0;
    ad.nMotionFlags |= opt ? 0x01 : 0;
    ad.nMotionFlags |= ad.isBackward ? 0x02 : 0;
    ad.nMotionFlags |= d->chroma ? 0x08 : 0;
    ad.nCPUFlags = 0; // MVAnalyse.c:492-494 stores the host's x264 cpu flags; meaningless here

    if (ad.nOverlapX % (1 << sp->ssw) || ad.nOverlapY % (1 << sp->ssh))
        return set_error("Analyse: The requested overlap is incompatible with the super clip's subsampling.");
    if (d->divide && (ad.nOverlapX % (2 << sp->ssw) || ad.nOverlapY % (2 << sp->ssh)))
        return set_error("Analyse: overlap and overlapv must be multiples of 2 or 4 when divide=True, depending on the super clip's subsampling.");
    if (ad.nDeltaFrame <= 0 && (-ad.nDeltaFrame) >= num_frames)
        return set_error("Analyse: delta points to frame past the input clip's end.");
    ad.yRatioUV = 1 << sp->ssh;
    ad.xRatioUV = 1 << sp->ssw;

    // Super_* props, MVAnalyse.c:523-568
    d->super_hpad = sp->hpad;
    d->super_vpad = sp->vpad;
    d->super_pel = sp->pel;
    d->super_modeyuv = sp->modeyuv;
    d->super_levels = sp->levels;
    d->ssw = sp->ssw;
    d->ssh = sp->ssh;
    d->gray = sp->gray;
    d->num_frames = num_frames;
    if (sp->height <= 0 || sp->hpad < 0 || sp->hpad >= sp->super_width / 2 || sp->vpad < 0 || sp->pel < 1 || sp->pel > 4 || sp->levels < 1)
        return set_error("Analyse: parameters from super clip appear to be wrong.");
    if ((nModeYUV & sp->modeyuv) != nModeYUV)
        return set_error("Analyse: super clip does not contain needed colour data.");

    ad.nWidth = sp->super_width - sp->hpad * 2;
    ad.nHeight = sp->height;
    ad.nPel = sp->pel;
    ad.nHPadding = sp->hpad;
    ad.nVPadding = sp->vpad;
    int nBlkX = (ad.nWidth - ad.nOverlapX) / (ad.nBlkSizeX - ad.nOverlapX);
    int nBlkY = (ad.nHeight - ad.nOverlapY) / (ad.nBlkSizeY - ad.nOverlapY);
    ad.nBlkX = nBlkX;
    ad.nBlkY = nBlkY;
    int nWidth_B = (ad.nBlkSizeX - ad.nOverlapX) * nBlkX + ad.nOverlapX;
    int nHeight_B = (ad.nBlkSizeY - ad.nOverlapY) * nBlkY + ad.nOverlapY;
    int nLevelsMax = 0;
    while (((nWidth_B >> nLevelsMax) - ad.nOverlapX) / (ad.nBlkSizeX - ad.nOverlapX) > 0 &&
           ((nHeight_B >> nLevelsMax) - ad.nOverlapY) / (ad.nBlkSizeY - ad.nOverlapY) > 0)
        nLevelsMax++;
    ad.nLvCount = levels > 0 ? levels : nLevelsMax + levels;
    if (ad.nLvCount < 1 || ad.nLvCount > nLevelsMax)
        return set_error("Analyse: invalid number of levels.");
    if (ad.nLvCount > sp->levels)
        return set_error("Analyse: super clip has %d levels. Analyse needs %d levels.", sp->levels, ad.nLvCount);
    if (ad.nLvCount > MVB200_MAX_LEVELS)
        return set_error("Analyse: invalid number of levels.");
    if (d->nPelSearch <= 0)
        d->nPelSearch = ad.nPel;
    if (d->divide) { // MVAnalyse.c:615-624
        d->ad_divided = ad;
        d->ad_divided.nBlkX = ad.nBlkX * 2;
        d->ad_divided.nBlkY = ad.nBlkY * 2;
        d->ad_divided.nBlkSizeX = ad.nBlkSizeX / 2;
        d->ad_divided.nBlkSizeY = ad.nBlkSizeY / 2;
        d->ad_divided.nOverlapX = ad.nOverlapX / 2;
        d->ad_divided.nOverlapY = ad.nOverlapY / 2;
        d->ad_divided.nLvCount = ad.nLvCount + 1;
    }
    // gopInit, GroupOfPlanes.c:43-55 + gopGetArraySize :167-174
    int size = 8;
    for (int i = 0; i < ad.nLvCount; i++) {
        d->nBlkX[i] = ((nWidth_B >> i) - ad.nOverlapX) / (ad.nBlkSizeX - ad.nOverlapX);
        d->nBlkY[i] = ((nHeight_B >> i) - ad.nOverlapY) / (ad.nBlkSizeY - ad.nOverlapY);
        size += 4 + d->nBlkX[i] * d->nBlkY[i] * 16;
        if (i == 0 && d->divide)
            size += 4 + d->nBlkX[i] * d->nBlkY[i] * 16 * 4;
    }
    d->vectors_size = size;
    return 0;
}

int mvb200_degrain_init(const mvb200_degrain_args *a, const mvb200_super_params *sp, const mvb200_analysis_data *vd, int clip_w, int clip_h, mvb200_degrain_params *d) {
    memset(d, 0, sizeof(*d));
    int radius = a->radius;
    if (radius < 1 || radius > MVB200_MAX_RADIUS)
        return set_error("Degrain: radius must be between 1 and 6.");
    char filter[16];
    snprintf(filter, sizeof(filter), "Degrain%d", radius);
    d->radius = radius;
    d->thSAD[0] = GIVEN64(a->thsad) ? a->thsad : 400;
    d->thSAD[1] = d->thSAD[2] = GIVEN64(a->thsadc) ? a->thsadc : d->thSAD[0];
    int plane = GIVEN(a->plane) ? a->plane : 4;
    d->nSCD1 = GIVEN64(a->thscd1) ? a->thscd1 : 400;
    d->nSCD2 = GIVEN(a->thscd2) ? a->thscd2 : 130;
    if (plane < 0 || plane > 4)
        return set_error("%s: plane must be between 0 and 4 (inclusive).", filter);
    static const int planes[5] = {MVB200_YPLANE, MVB200_UPLANE, MVB200_VPLANE, MVB200_UPLANE | MVB200_VPLANE, MVB200_YUVPLANES};
    d->YUVplanes = planes[plane];
    for (int r = 0; r < radius * 2; r++)
        d->vd[r] = vd[r];
    int64_t nSCD1_old = d->nSCD1;
    if (scale_thscd(&d->nSCD1, &d->nSCD2, &d->vd[0], filter))
        return 1;
    static const char *names[12] = {"mvbw", "mvfw", "mvbw2", "mvfw2", "mvbw3", "mvfw3", "mvbw4", "mvfw4", "mvbw5", "mvfw5", "mvbw6", "mvfw6"};
    for (int r = 1; r < radius * 2; r++)
        if (check_similarity(&d->vd[0], &d->vd[r], filter, names[0], names[r]))
            return 1;
    // vector-clip ordering checks, MVDegrains.cpp:616-655 (the LAST failing check wins)
    const char *err = nullptr;
    for (int r = 0; r < radius * 2; r++)
        if (d->vd[r].nDeltaFrame <= 0)
            err = "cannot use motion vectors with absolute frame references.";
    if (!d->vd[0].isBackward) err = "mvbw must be generated with isb=True.";
    if (d->vd[1].isBackward) err = "mvfw must be generated with isb=False.";
    for (int k = 1; k < radius; k++) {
        if (!d->vd[2 * k].isBackward) err = "mvbw must be generated with isb=True.";
        if (d->vd[2 * k + 1].isBackward) err = "mvfw must be generated with isb=False.";
        if (d->vd[2 * k].nDeltaFrame <= d->vd[2 * k - 2].nDeltaFrame) err = "mvbwN must have greater delta than mvbwP.";
        if (d->vd[2 * k + 1].nDeltaFrame <= d->vd[2 * k - 1].nDeltaFrame) err = "mvfwN must have greater delta than mvfwP.";
    }
    if (err)
        return set_error("%s: %s", filter, err);
    d->thSAD[0] = d->thSAD[0] * d->nSCD1 / nSCD1_old;
    d->thSAD[1] = d->thSAD[2] = d->thSAD[1] * d->nSCD1 / nSCD1_old;
    if (d->thSAD[0] >= INT_MAX || d->thSAD[1] >= INT_MAX) {
        int64_t maximum = (int64_t)INT_MAX * nSCD1_old / d->nSCD1;
        bool c = d->thSAD[0] < INT_MAX;
        return set_error("%s: with this block size and video format, thsad%s must not exceed %lld or some calculations would overflow.", filter, c ? "c" : "", (long long)maximum);
    }
    if (d->vd[0].nHeight != sp->height || d->vd[0].nHeight != clip_h || d->vd[0].nWidth != sp->super_width - sp->hpad * 2 || d->vd[0].nWidth != clip_w || d->vd[0].nPel != sp->pel)
        return set_error("%s: wrong source or super clip frame size.", filter);
    if (sp->bits > 16 || sp->ssw > 1 || sp->ssh > 1)
        return set_error("%s: input clip must be GRAY, 420, 422, 440, or 444, up to 16 bits, with constant dimensions.", filter);
    int pixelMax = (1 << sp->bits) - 1;
    d->nLimit[0] = GIVEN(a->limit) ? a->limit : pixelMax;
    d->nLimit[1] = d->nLimit[2] = GIVEN(a->limitc) ? a->limitc : d->nLimit[0];
    if (d->nLimit[0] < 0 || d->nLimit[0] > pixelMax)
        return set_error("%s: limit must be between 0 and %d (inclusive).", filter, pixelMax);
    if (d->nLimit[1] < 0 || d->nLimit[1] > pixelMax)
        return set_error("%s: limitc must be between 0 and %d (inclusive).", filter, pixelMax);
    d->bits = sp->bits;
    d->ssw = sp->ssw;
    d->ssh = sp->ssh;
    d->gray = sp->gray;
    d->nplanes = sp->nplanes;
    d->super_hpad = sp->hpad;
    d->super_vpad = sp->vpad;
    d->super_pel = sp->pel;
    d->super_modeyuv = sp->modeyuv;
    d->super_levels = sp->levels;
    d->process[0] = !!(d->YUVplanes & MVB200_YPLANE);
    d->process[1] = !!(d->YUVplanes & MVB200_UPLANE & sp->modeyuv);
    d->process[2] = !!(d->YUVplanes & MVB200_VPLANE & sp->modeyuv);
    const mvb200_analysis_data &v0 = d->vd[0];
    for (int p = 0; p < 3; p++) {
        int xs = p ? sp->ssw : 0, ys = p ? sp->ssh : 0;
        d->nWidth[p] = v0.nWidth >> xs;
        d->nHeight[p] = v0.nHeight >> ys;
        d->nOverlapX[p] = v0.nOverlapX >> xs;
        d->nOverlapY[p] = v0.nOverlapY >> ys;
        d->nBlkSizeX[p] = v0.nBlkSizeX >> xs;
        d->nBlkSizeY[p] = v0.nBlkSizeY >> ys;
        d->nWidth_B[p] = (v0.nBlkX * (v0.nBlkSizeX - v0.nOverlapX) + v0.nOverlapX) >> xs;
        d->nHeight_B[p] = (v0.nBlkY * (v0.nBlkSizeY - v0.nOverlapY) + v0.nOverlapY) >> ys;
    }
    return 0;
}

// mvcompensateCreate, MVCompensate.c:419-575
int mvb200_compensate_init(const mvb200_compensate_args *a, const mvb200_super_params *sp, const mvb200_analysis_data *vd, int clip_w, int clip_h, int num_frames, mvb200_compensate_params *d) {
    memset(d, 0, sizeof(*d));
    d->scBehavior = GIVEN(a->scbehavior) ? !!a->scbehavior : 1;
    d->thSAD = GIVEN64(a->thsad) ? a->thsad : 10000;
    d->fields = GIVEN(a->fields) && a->fields != 0;
    double time = (a->time == a->time) ? a->time : 100.0; // NaN = not given
    d->nSCD1 = GIVEN64(a->thscd1) ? a->thscd1 : 400;
    d->nSCD2 = GIVEN(a->thscd2) ? a->thscd2 : 130;
    d->tff = GIVEN(a->tff) && a->tff != 0;
    d->tff_exists = GIVEN(a->tff);
    if (time < 0.0 || time > 100.0)
        return set_error("Compensate: time must be between 0.0 and 100.0 (inclusive).");
    d->vd = *vd;
    int64_t nSCD1_old = d->nSCD1;
    if (scale_thscd(&d->nSCD1, &d->nSCD2, &d->vd, "Compensate"))
        return 1;
    if (d->fields && vd->nPel < 2)
        return set_error("Compensate: fields option requires pel > 1.");
    d->thSAD = d->thSAD * d->nSCD1 / nSCD1_old;
    if (vd->nHeight != sp->height || vd->nHeight != clip_h || vd->nWidth != sp->super_width - sp->hpad * 2 || vd->nWidth != clip_w || vd->nPel != sp->pel)
        return set_error("Compensate: wrong source or super clip frame size.");
    if (sp->bits > 16 || sp->ssw > 1 || sp->ssh > 1)
        return set_error("Compensate: input clip must be GRAY, 420, 422, 440, or 444, up to 16 bits, with constant dimensions.");
    d->time256 = (int)(time * 256 / 100);
    d->bits = sp->bits;
    d->ssw = sp->ssw;
    d->ssh = sp->ssh;
    d->gray = sp->gray;
    d->nplanes = sp->nplanes;
    d->super_hpad = sp->hpad;
    d->super_vpad = sp->vpad;
    d->super_pel = sp->pel;
    d->super_modeyuv = sp->modeyuv;
    d->super_levels = sp->levels;
    d->num_frames = num_frames;
    const int xs = vd->xRatioUV == 2 ? 1 : 0, ys = vd->yRatioUV == 2 ? 1 : 0;
    for (int p = 0; p < 3; p++) {
        int x = p ? xs : 0, y = p ? ys : 0;
        d->nWidth[p] = vd->nWidth >> x;
        d->nHeight[p] = vd->nHeight >> y;
        d->nOverlapX[p] = vd->nOverlapX >> x;
        d->nOverlapY[p] = vd->nOverlapY >> y;
        d->nBlkSizeX[p] = vd->nBlkSizeX >> x;
        d->nBlkSizeY[p] = vd->nBlkSizeY >> y;
        d->nWidth_B[p] = (vd->nBlkX * (vd->nBlkSizeX - vd->nOverlapX) + vd->nOverlapX) >> x;
        d->nHeight_B[p] = (vd->nBlkY * (vd->nBlkSizeY - vd->nOverlapY) + vd->nOverlapY) >> y;
    }
    return 0;
}

// mvrecalculateCreate, MVRecalculate.c:265-553
int mvb200_recalculate_init(const mvb200_recalculate_args *a, const mvb200_super_params *sp, const mvb200_analysis_data *old, int num_frames, mvb200_recalculate_params *d) {
    memset(d, 0, sizeof(*d));
    mvb200_analysis_data &ad = d->ad;
    d->thSAD = GIVEN64(a->thsad) ? a->thsad : 200;
    d->smooth = GIVEN(a->smooth) ? a->smooth : 1;
    ad.nBlkSizeX = GIVEN(a->blksize) ? a->blksize : 8;
    ad.nBlkSizeY = GIVEN(a->blksizev) ? a->blksizev : ad.nBlkSizeX;
    d->search = GIVEN(a->search) ? a->search : 4;
    int searchparam = GIVEN(a->searchparam) ? a->searchparam : 2;
    d->chroma = GIVEN(a->chroma) ? !!a->chroma : 1;
    int truemotion = GIVEN(a->truemotion) ? !!a->truemotion : 1;
    d->nLambda = GIVEN(a->lambda) ? a->lambda : (truemotion ? (1000 * ad.nBlkSizeX * ad.nBlkSizeY / 64) : 0);
    d->pnew = GIVEN(a->pnew) ? a->pnew : (truemotion ? 50 : 0);
    ad.nOverlapX = GIVEN(a->overlap) ? a->overlap : 0;
    ad.nOverlapY = GIVEN(a->overlapv) ? a->overlapv : ad.nOverlapX;
    d->dctmode = GIVEN(a->dct) ? a->dct : 0;
    d->divide = GIVEN(a->divide) ? a->divide : 0;
    int opt = GIVEN(a->opt) ? !!a->opt : 1;
    d->meander = GIVEN(a->meander) ? !!a->meander : 1;
    d->fields = GIVEN(a->fields) && a->fields != 0;
    d->tff = GIVEN(a->tff) && a->tff != 0;
    d->tff_exists = GIVEN(a->tff);
    if (d->search < 0 || d->search > 7) return set_error("Recalculate: search must be between 0 and 7 (inclusive).");
    if (d->dctmode < 0 || d->dctmode > 10) return set_error("Recalculate: dct must be between 0 and 10 (inclusive).");
    if (d->dctmode >= 5 && ad.nBlkSizeX == 16 && ad.nBlkSizeY == 2) return set_error("Recalculate: dct 5..10 cannot work with 16x2 blocks.");
    if (d->divide < 0 || d->divide > 2) return set_error("Recalculate: divide must be between 0 and 2 (inclusive).");
    {
        static const int ok[12][2] = {{4, 4}, {8, 4}, {8, 8}, {16, 2}, {16, 8}, {16, 16}, {32, 16}, {32, 32}, {64, 32}, {64, 64}, {128, 64}, {128, 128}};
        bool found = false;
        for (auto &s : ok)
            found |= (s[0] == ad.nBlkSizeX && s[1] == ad.nBlkSizeY);
        if (!found)
            return set_error("Recalculate: the block size must be 4x4, 8x4, 8x8, 16x2, 16x8, 16x16, 32x16, 32x32, 64x32, 64x64, 128x64, or 128x128.");
    }
    if (d->pnew < 0 || d->pnew > 256) return set_error("Recalculate: pnew must be between 0 and 256 (inclusive).");
    if (ad.nOverlapX < 0 || ad.nOverlapX > ad.nBlkSizeX / 2 || ad.nOverlapY < 0 || ad.nOverlapY > ad.nBlkSizeY / 2)
        return set_error("Recalculate: overlap must be at most half of blksize, overlapv must be at most half of blksizev, and they both need to be at least 0.");
    if (d->divide && (ad.nBlkSizeX < 8 || ad.nBlkSizeY < 8))
        return set_error("Recalculate: blksize and blksizev must be at least 8 when divide=True.");
    if (d->search == 1)
        d->nSearchParam = searchparam < 0 ? 0 : searchparam;
    else
        d->nSearchParam = searchparam < 1 ? 1 : searchparam;
    if (ad.nOverlapX % (1 << sp->ssw) || ad.nOverlapY % (1 << sp->ssh))
        return set_error("Recalculate: The requested overlap is incompatible with the super clip's subsampling.");
    if (d->divide && (ad.nOverlapX % (2 << sp->ssw) || ad.nOverlapY % (2 << sp->ssh)))
        return set_error("Recalculate: overlap and overlapv must be multiples of 2 or 4 when divide=True, depending on the super clip's subsampling.");
    if (sp->gray)
        d->chroma = 0;
    int nModeYUV = d->chroma ? MVB200_YUVPLANES : MVB200_YPLANE;
    if ((nModeYUV & sp->modeyuv) != nModeYUV)
        return set_error("Recalculate: super clip does not contain needed colour data.");
    d->old = *old;
    ad.yRatioUV = old->yRatioUV;
    ad.xRatioUV = old->xRatioUV;
    ad.nWidth = old->nWidth;
    ad.nHeight = old->nHeight;
    ad.nDeltaFrame = old->nDeltaFrame;
    ad.isBackward = old->isBackward;
    ad.bitsPerSample = sp->bits;
    int pixelMax = (1 << sp->bits) - 1;
    d->thSAD = (int64_t)((double)d->thSAD * pixelMax / 255.0 + 0.5);
    d->nLambda = (int)((double)d->nLambda * pixelMax / 255.0 + 0.5);
    d->thSAD = d->thSAD * (ad.nBlkSizeX * ad.nBlkSizeY) / 64;
    if (d->chroma)
        d->thSAD += d->thSAD / (ad.xRatioUV * ad.yRatioUV) * 2;
    ad.nMotionFlags = (opt ? 0x01 : 0) | (ad.isBackward ? 0x02 : 0) | (d->chroma ? 0x08 : 0);
    ad.nCPUFlags = 0;
    ad.nPel = sp->pel;
    if (sp->height != ad.nHeight || sp->super_width - 2 * sp->hpad != ad.nWidth)
        return set_error("Recalculate: wrong frame size.");
    ad.nHPadding = sp->hpad;
    ad.nVPadding = sp->vpad;
    ad.nBlkX = (ad.nWidth - ad.nOverlapX) / (ad.nBlkSizeX - ad.nOverlapX);
    ad.nBlkY = (ad.nHeight - ad.nOverlapY) / (ad.nBlkSizeY - ad.nOverlapY);
    ad.nLvCount = 1;
    if (d->divide) {
        d->ad_divided = ad;
        d->ad_divided.nBlkX = ad.nBlkX * 2;
        d->ad_divided.nBlkY = ad.nBlkY * 2;
        d->ad_divided.nBlkSizeX = ad.nBlkSizeX / 2;
        d->ad_divided.nBlkSizeY = ad.nBlkSizeY / 2;
        d->ad_divided.nOverlapX = ad.nOverlapX / 2;
        d->ad_divided.nOverlapY = ad.nOverlapY / 2;
        d->ad_divided.nLvCount = ad.nLvCount + 1;
    }
    d->super_hpad = sp->hpad;
    d->super_vpad = sp->vpad;
    d->super_pel = sp->pel;
    d->super_modeyuv = sp->modeyuv;
    d->super_levels = sp->levels;
    d->ssw = sp->ssw;
    d->ssh = sp->ssh;
    d->gray = sp->gray;
    d->num_frames = num_frames;
    int n = ad.nBlkX * ad.nBlkY;
    d->vectors_size = 8 + 4 + n * 16 + (d->divide ? 4 + n * 16 * 4 : 0);
    return 0;
}

// ------------------------------------------- SCDetection / Mask / Finest / Flow

int mvb200_scdetection_init(int64_t thscd1, int32_t thscd2, const mvb200_analysis_data *vd, mvb200_scdetection_params *d) {
    memset(d, 0, sizeof(*d));
    d->vd = *vd;
    d->nSCD1 = GIVEN64(thscd1) ? thscd1 : 400; // MV_DEFAULT_SCD1/2, MVAnalysisData.h
    d->nSCD2 = GIVEN(thscd2) ? thscd2 : 130;
    return scale_thscd(&d->nSCD1, &d->nSCD2, &d->vd, "SCDetection");
}

int mvb200_mask_init(const mvb200_mask_args *a, const mvb200_analysis_data *vd, int clip_bits, int clip_ssw, int clip_ssh, int clip_gray, mvb200_mask_params *d) {
    memset(d, 0, sizeof(*d));
    // MVMask.c:230-258
    d->ml = (a->ml == a->ml) ? (float)a->ml : 100.0f;
    d->fGamma = (a->gamma == a->gamma) ? (float)a->gamma : 1.0f;
    d->kind = GIVEN(a->kind) ? a->kind : 0;
    const double time = (a->time == a->time) ? a->time : 100.0;
    d->nSceneChangeValue = GIVEN(a->ysc) ? a->ysc : 0;
    d->nSCD1 = GIVEN64(a->thscd1) ? a->thscd1 : 400;
    d->nSCD2 = GIVEN(a->thscd2) ? a->thscd2 : 130;
    // :261-279
    if (d->fGamma < 0.0f)
        return set_error("Mask: gamma must not be negative.");
    if (d->kind < 0 || d->kind > 5)
        return set_error("Mask: kind must 0, 1, 2, 3, 4, or 5.");
    if (time < 0.0 || time > 100.0)
        return set_error("Mask: time must be between 0.0 and 100.0 (inclusive).");
    if (d->nSceneChangeValue < 0 || d->nSceneChangeValue > 255)
        return set_error("Mask: ysc must be between 0 and 255 (inclusive).");
    d->vd = *vd;
    if (scale_thscd(&d->nSCD1, &d->nSCD2, &d->vd, "Mask"))
        return 1;
    // :299-311 (float arithmetic, as in the reference)
    d->fMaskNormFactor = 1.0f / d->ml;
    d->fMaskNormFactor2 = d->fMaskNormFactor * d->fMaskNormFactor;
    d->fHalfGamma = d->fGamma * 0.5f;
    d->nWidth = vd->nWidth;
    d->nHeight = vd->nHeight;
    d->nWidthB = vd->nBlkX * (vd->nBlkSizeX - vd->nOverlapX) + vd->nOverlapX;
    d->nHeightB = vd->nBlkY * (vd->nBlkSizeY - vd->nOverlapY) + vd->nOverlapY;
    d->nHeightUV = vd->nHeight / vd->yRatioUV;
    d->nWidthUV = vd->nWidth / vd->xRatioUV;
    d->nHeightBUV = d->nHeightB / vd->yRatioUV;
    d->nWidthBUV = d->nWidthB / vd->xRatioUV;
    // :317-322
    if (clip_bits > 8 || clip_ssw > 1 || clip_ssh > 1)
        return set_error("Mask: input clip must be GRAY8, YUV420P8, YUV422P8, YUV440P8, or YUV444P8, with constant dimensions.");
    d->clip_gray = clip_gray != 0;
    d->time256 = (int)(time * 256 / 100);
    return 0;
}

int mvb200_flow_init(const mvb200_flow_args *a, const mvb200_super_params *sp, const mvb200_analysis_data *vd, int num_frames, mvb200_flow_params *d) {
    memset(d, 0, sizeof(*d));
    // MVFlow.cpp:403-438
    const double time = (a->time == a->time) ? a->time : 100.0;
    d->mode = GIVEN(a->mode) ? a->mode : 0;
    d->fields = GIVEN(a->fields) && a->fields != 0;
    d->nSCD1 = GIVEN64(a->thscd1) ? a->thscd1 : 400;
    d->nSCD2 = GIVEN(a->thscd2) ? a->thscd2 : 130;
    d->tff = GIVEN(a->tff) && a->tff != 0;
    d->tff_exists = GIVEN(a->tff);
    if (time < 0.0 || time > 100.0)
        return set_error("Flow: time must be between 0 and 100 %% (inclusive).");
    if (d->mode < 0 || d->mode > 1)
        return set_error("Flow: mode must be 0 or 1.");
    d->time256 = (int)(time * 256.0 / 100.0);
    d->vd = *vd;
    if (scale_thscd(&d->nSCD1, &d->nSCD2, &d->vd, "Flow"))
        return 1;
    // :514-533
    if (vd->nHeight != sp->height || vd->nWidth != sp->super_width - sp->hpad * 2 || vd->nPel != sp->pel)
        return set_error("Flow: wrong source or super clip frame size.");
    if (sp->bits > 16 || sp->ssw > 1 || sp->ssh > 1)
        return set_error("Flow: input clip must be GRAY, 420, 422, 440, or 444, up to 16 bits, with constant dimensions.");
    // :535-557
    d->nBlkXP = vd->nBlkX;
    d->nBlkYP = vd->nBlkY;
    while (d->nBlkXP * (vd->nBlkSizeX - vd->nOverlapX) + vd->nOverlapX < vd->nWidth)
        d->nBlkXP++;
    while (d->nBlkYP * (vd->nBlkSizeY - vd->nOverlapY) + vd->nOverlapY < vd->nHeight)
        d->nBlkYP++;
    d->nWidthP = d->nBlkXP * (vd->nBlkSizeX - vd->nOverlapX) + vd->nOverlapX;
    d->nHeightP = d->nBlkYP * (vd->nBlkSizeY - vd->nOverlapY) + vd->nOverlapY;
    d->nWidthPUV = d->nWidthP / vd->xRatioUV;
    d->nHeightPUV = d->nHeightP / vd->yRatioUV;
    d->nHeightUV = vd->nHeight / vd->yRatioUV;
    d->nWidthUV = vd->nWidth / vd->xRatioUV;
    d->nHPaddingUV = vd->nHPadding / vd->xRatioUV;
    d->nVPaddingUV = vd->nVPadding / vd->yRatioUV;
    d->VPitchY = (d->nWidthP + 15) & (~15);
    d->VPitchUV = (d->nWidthPUV + 15) & (~15);
    d->bits = sp->bits;
    d->ssw = sp->ssw;
    d->ssh = sp->ssh;
    d->gray = sp->gray;
    d->nplanes = sp->nplanes;
    d->super_hpad = sp->hpad;
    d->super_vpad = sp->vpad;
    d->super_pel = sp->pel;
    d->super_levels = sp->levels;
    d->num_frames = num_frames;
    return 0;
}

int mvb200_flowinter_init(const mvb200_flowinter_args *a, const mvb200_super_params *sp, const mvb200_analysis_data *vd_bw, const mvb200_analysis_data *vd_fw, int num_frames, mvb200_flowinter_params *d) {
    memset(d, 0, sizeof(*d));
    // MVFlowInter.c:466-503 (time and ml are floats there)
    const float time = (a->time == a->time) ? (float)a->time : 50.0f;
    const float ml = (a->ml == a->ml) ? (float)a->ml : 100.0f;
    d->blend = GIVEN(a->blend) ? a->blend != 0 : 1;
    d->nSCD1 = GIVEN64(a->thscd1) ? a->thscd1 : 400;
    d->nSCD2 = GIVEN(a->thscd2) ? a->thscd2 : 130;
    if (time < 0.0f || time > 100.0f)
        return set_error("FlowInter: time must be between 0 and 100 %% (inclusive).");
    if (ml <= 0.0f)
        return set_error("FlowInter: ml must be greater than 0.");
    d->time256 = (int)(time * 256.0f / 100.0f);
    d->ml = ml;
    d->vd = *vd_bw;
    // :537-580
    if (scale_thscd(&d->nSCD1, &d->nSCD2, &d->vd, "FlowInter"))
        return 1;
    if (check_similarity(vd_bw, vd_fw, "FlowInter", "mvbw", "mvfw"))
        return 1;
    if (vd_bw->nDeltaFrame <= 0 || vd_fw->nDeltaFrame <= 0)
        return set_error("FlowInter: cannot use motion vectors with absolute frame references.");
    if (vd_fw->nDeltaFrame != vd_bw->nDeltaFrame)
        return set_error("FlowInter: mvbw and mvfw must be generated with the same delta.");
    if (!vd_bw->isBackward)
        return set_error("FlowInter: mvbw must be generated with isb=True.");
    if (vd_fw->isBackward)
        return set_error("FlowInter: mvfw must be generated with isb=False.");
    // :617-635
    if (vd_bw->nHeight != sp->height || vd_bw->nWidth != sp->super_width - sp->hpad * 2 || vd_bw->nPel != sp->pel)
        return set_error("FlowInter: wrong source or super clip frame size.");
    if (sp->bits > 16 || sp->ssw > 1 || sp->ssh > 1)
        return set_error("FlowInter: input clip must be GRAY, 420, 422, 440, or 444, up to 16 bits, with constant dimensions.");
    d->mode_primary = 2;
    d->mode_fallback = 1;
    // :638-659
    const mvb200_analysis_data *vd = vd_bw;
    d->nBlkXP = vd->nBlkX;
    d->nBlkYP = vd->nBlkY;
    while (d->nBlkXP * (vd->nBlkSizeX - vd->nOverlapX) + vd->nOverlapX < vd->nWidth)
        d->nBlkXP++;
    while (d->nBlkYP * (vd->nBlkSizeY - vd->nOverlapY) + vd->nOverlapY < vd->nHeight)
        d->nBlkYP++;
    d->nWidthP = d->nBlkXP * (vd->nBlkSizeX - vd->nOverlapX) + vd->nOverlapX;
    d->nHeightP = d->nBlkYP * (vd->nBlkSizeY - vd->nOverlapY) + vd->nOverlapY;
    d->nWidthPUV = d->nWidthP / vd->xRatioUV;
    d->nHeightPUV = d->nHeightP / vd->yRatioUV;
    d->nHeightUV = vd->nHeight / vd->yRatioUV;
    d->nWidthUV = vd->nWidth / vd->xRatioUV;
    d->nHPaddingUV = vd->nHPadding / vd->xRatioUV;
    d->nVPaddingUV = vd->nVPadding / vd->yRatioUV;
    d->VPitchY = (d->nWidthP + 15) & (~15);
    d->VPitchUV = (d->nWidthPUV + 15) & (~15);
    d->bits = sp->bits;
    d->ssw = sp->ssw;
    d->ssh = sp->ssh;
    d->gray = sp->gray;
    d->nplanes = sp->nplanes;
    d->super_hpad = sp->hpad;
    d->super_vpad = sp->vpad;
    d->super_pel = sp->pel;
    d->super_levels = sp->levels;
    d->num_frames = num_frames;
    return 0;
}

static int64_t gcd64(int64_t x, int64_t y) { // MVFlowFPS.c gcd()
    while (y) {
        int64_t t = x % y;
        x = y;
        y = t;
    }
    return x;
}

int mvb200_flowfps_init(const mvb200_flowfps_args *a, const mvb200_super_params *sp, const mvb200_analysis_data *vd_bw, const mvb200_analysis_data *vd_fw, int64_t fps_num, int64_t fps_den, int clip_w, int clip_h, int num_frames,
                        mvb200_flowfps_params *d) {
    memset(d, 0, sizeof(*d));
    mvb200_flowinter_params *q = &d->inter;
    // MVFlowFPS.c:571-612
    const int64_t num = GIVEN64(a->num) ? a->num : 25, den = GIVEN64(a->den) ? a->den : 1;
    d->maskmode = GIVEN(a->mask) ? a->mask : 2;
    q->ml = (a->ml == a->ml) ? a->ml : 100.0;
    q->blend = GIVEN(a->blend) ? a->blend != 0 : 1;
    q->nSCD1 = GIVEN64(a->thscd1) ? a->thscd1 : 400;
    q->nSCD2 = GIVEN(a->thscd2) ? a->thscd2 : 130;
    if (d->maskmode < 0 || d->maskmode > 2)
        return set_error("FlowFPS: mask must be 0, 1, or 2.");
    if (q->ml <= 0.0)
        return set_error("FlowFPS: ml must be greater than 0.");
    q->vd = *vd_bw;
    // :650-690
    if (scale_thscd(&q->nSCD1, &q->nSCD2, &q->vd, "FlowFPS"))
        return 1;
    if (check_similarity(vd_bw, vd_fw, "FlowFPS", "mvbw", "mvfw"))
        return 1;
    if (vd_bw->nDeltaFrame != vd_fw->nDeltaFrame)
        return set_error("FlowFPS: mvbw and mvfw must be generated with the same delta.");
    if (!vd_bw->isBackward)
        return set_error("FlowFPS: mvbw must be generated with isb=True.");
    if (vd_fw->isBackward)
        return set_error("FlowFPS: mvfw must be generated with isb=False.");
    // :725-756
    if (fps_num == 0 || fps_den == 0)
        return set_error("FlowFPS: The input clip must have a frame rate. Invoke AssumeFPS if necessary.");
    int64_t numerator, denominator;
    if (num != 0 && den != 0) {
        numerator = num;
        denominator = den;
    } else {
        numerator = fps_num * 2;
        denominator = fps_den;
    }
    d->fa = denominator * fps_num;
    d->fb = numerator * fps_den;
    const int64_t g = gcd64(d->fa, d->fb);
    d->fa /= g;
    d->fb /= g;
    if (numerator <= 0 || denominator <= 0) { // setFPS, :547-561
        d->fps_num = 0;
        d->fps_den = 1;
    } else {
        const int64_t x = gcd64(numerator, denominator);
        d->fps_num = numerator / x;
        d->fps_den = denominator / x;
    }
    d->num_frames_in = num_frames;
    d->num_frames_out = (int)(1 + (num_frames - 1) * d->fb / d->fa);
    // :759-800
    if (vd_bw->nWidth != clip_w || vd_bw->nHeight != clip_h)
        return set_error("FlowFPS: inconsistent source and vector frame size.");
    if (vd_bw->nHeight != sp->height || vd_bw->nWidth != sp->super_width - sp->hpad * 2 || vd_bw->nPel != sp->pel)
        return set_error("FlowFPS: wrong source or super clip frame size.");
    if (!((vd_bw->nWidth + vd_bw->nHPadding * 2) == sp->super_width && (vd_bw->nHeight + vd_bw->nVPadding * 2) <= sp->super_height))
        return set_error("FlowFPS: inconsistent clips frame size! Incomprehensible error messages are the best, right?");
    if (sp->bits > 16 || sp->ssw > 1 || sp->ssh > 1)
        return set_error("FlowFPS: input clip must be GRAY, 420, 422, 440, or 444, up to 16 bits, with constant dimensions.");
    // which blend runs (MVFlowFPS.c:339-437): mask 2 -> Extra, or Simple when the extra vectors are unusable
    q->mode_primary = d->maskmode == 2 ? 2 : (d->maskmode == 1 ? 1 : 0);
    q->mode_fallback = d->maskmode == 1 ? 1 : 0;
    const mvb200_analysis_data *vd = vd_bw;
    q->nBlkXP = vd->nBlkX;
    q->nBlkYP = vd->nBlkY;
    while (q->nBlkXP * (vd->nBlkSizeX - vd->nOverlapX) + vd->nOverlapX < vd->nWidth)
        q->nBlkXP++;
    while (q->nBlkYP * (vd->nBlkSizeY - vd->nOverlapY) + vd->nOverlapY < vd->nHeight)
        q->nBlkYP++;
    q->nWidthP = q->nBlkXP * (vd->nBlkSizeX - vd->nOverlapX) + vd->nOverlapX;
    q->nHeightP = q->nBlkYP * (vd->nBlkSizeY - vd->nOverlapY) + vd->nOverlapY;
    q->nWidthPUV = q->nWidthP / vd->xRatioUV;
    q->nHeightPUV = q->nHeightP / vd->yRatioUV;
    q->nHeightUV = vd->nHeight / vd->yRatioUV;
    q->nWidthUV = vd->nWidth / vd->xRatioUV;
    q->nHPaddingUV = vd->nHPadding / vd->xRatioUV;
    q->nVPaddingUV = vd->nVPadding / vd->yRatioUV;
    q->VPitchY = (q->nWidthP + 15) & (~15);
    q->VPitchUV = (q->nWidthPUV + 15) & (~15);
    q->bits = sp->bits;
    q->ssw = sp->ssw;
    q->ssh = sp->ssh;
    q->gray = sp->gray;
    q->nplanes = sp->nplanes;
    q->super_hpad = sp->hpad;
    q->super_vpad = sp->vpad;
    q->super_pel = sp->pel;
    q->super_levels = sp->levels;
    q->num_frames = num_frames;
    return 0;
}

void mvb200_flowfps_position(const mvb200_flowfps_params *p, int n, int32_t *nleft, int32_t *time256) {
    const int off = p->inter.vd.nDeltaFrame;
    *nleft = (int)(n * p->fa / p->fb);
    int t = (int)(((double)n * p->fa / p->fb - *nleft) * 256 + 0.5);
    if (off > 1)
        t = t / off;
    *time256 = t;
}

int mvb200_flowblur_init(const mvb200_flowblur_args *a, const mvb200_super_params *sp, const mvb200_analysis_data *vd_bw, const mvb200_analysis_data *vd_fw, int num_frames, mvb200_flowblur_params *d) {
    memset(d, 0, sizeof(*d));
    // MVFlowBlur.c:353-384 (blur is a float there)
    const float blur = (a->blur == a->blur) ? (float)a->blur : 50.0f;
    d->prec = GIVEN(a->prec) ? a->prec : 1;
    d->nSCD1 = GIVEN64(a->thscd1) ? a->thscd1 : 400;
    d->nSCD2 = GIVEN(a->thscd2) ? a->thscd2 : 130;
    if (blur < 0.0f || blur > 200.0f)
        return set_error("FlowBlur: blur must be between 0 and 200 %% (inclusive).");
    if (d->prec < 1)
        return set_error("FlowBlur: prec must be at least 1.");
    d->blur256 = (int)(blur * 256.0f / 200.0f);
    d->vd = *vd_bw;
    // :418-461
    if (scale_thscd(&d->nSCD1, &d->nSCD2, &d->vd, "FlowBlur"))
        return 1;
    if (check_similarity(vd_bw, vd_fw, "FlowBlur", "mvbw", "mvfw"))
        return 1;
    if (vd_bw->nDeltaFrame <= 0 || vd_fw->nDeltaFrame <= 0)
        return set_error("FlowBlur: cannot use motion vectors with absolute frame references.");
    if (vd_fw->nDeltaFrame != vd_bw->nDeltaFrame)
        return set_error("FlowBlur: mvbw and mvfw must be generated with the same delta.");
    if (!vd_bw->isBackward)
        return set_error("FlowBlur: mvbw must be generated with isb=True.");
    if (vd_fw->isBackward)
        return set_error("FlowBlur: mvfw must be generated with isb=False.");
    // :498-516
    if (vd_bw->nHeight != sp->height || vd_bw->nWidth != sp->super_width - sp->hpad * 2 || vd_bw->nPel != sp->pel)
        return set_error("FlowBlur: wrong source or super clip frame size.");
    if (sp->bits > 16 || sp->ssw > 1 || sp->ssh > 1)
        return set_error("FlowBlur: input clip must be GRAY, 420, 422, 440, or 444, up to 16 bits, with constant dimensions.");
    // :519-529
    d->nHeightUV = vd_bw->nHeight / vd_bw->yRatioUV;
    d->nWidthUV = vd_bw->nWidth / vd_bw->xRatioUV;
    d->nHPaddingUV = vd_bw->nHPadding / vd_bw->xRatioUV;
    d->nVPaddingUV = vd_bw->nVPadding / vd_bw->yRatioUV;
    d->bits = sp->bits;
    d->ssw = sp->ssw;
    d->ssh = sp->ssh;
    d->gray = sp->gray;
    d->nplanes = sp->nplanes;
    d->super_hpad = sp->hpad;
    d->super_vpad = sp->vpad;
    d->super_pel = sp->pel;
    d->super_levels = sp->levels;
    d->num_frames = num_frames;
    return 0;
}

int mvb200_blockfps_init(const mvb200_blockfps_args *a, const mvb200_super_params *sp, const mvb200_analysis_data *vd_bw, const mvb200_analysis_data *vd_fw, int64_t fps_num, int64_t fps_den, int clip_w, int clip_h, int num_frames,
                         mvb200_blockfps_params *d) {
    memset(d, 0, sizeof(*d));
    // MVBlockFPS.c:744-782
    const int64_t num = GIVEN64(a->num) ? a->num : 25, den = GIVEN64(a->den) ? a->den : 1;
    d->mode = GIVEN(a->mode) ? a->mode : 3;
    d->ml = (a->ml == a->ml) ? a->ml : 100.0;
    d->blend = GIVEN(a->blend) ? a->blend != 0 : 1;
    d->nSCD1 = GIVEN64(a->thscd1) ? a->thscd1 : 400;
    d->nSCD2 = GIVEN(a->thscd2) ? a->thscd2 : 130;
    if (d->mode < 0 || d->mode > 8)
        return set_error("BlockFPS: mode must be between 0 and 8 (inclusive).");
    d->vd = *vd_bw;
    // :823-868
    if (scale_thscd(&d->nSCD1, &d->nSCD2, &d->vd, "BlockFPS"))
        return 1;
    if (check_similarity(vd_bw, vd_fw, "BlockFPS", "mvbw", "mvfw"))
        return 1;
    if (vd_bw->nDeltaFrame <= 0 || vd_fw->nDeltaFrame <= 0)
        return set_error("BlockFPS: cannot use motion vectors with absolute frame references.");
    if (vd_bw->nDeltaFrame != vd_fw->nDeltaFrame)
        return set_error("BlockFPS: mvbw and mvfw must be generated with the same delta.");
    if (!vd_bw->isBackward)
        return set_error("BlockFPS: mvbw must be generated with isb=True.");
    if (vd_fw->isBackward)
        return set_error("BlockFPS: mvfw must be generated with isb=False.");
    // :875-910
    if (fps_num == 0 || fps_den == 0)
        return set_error("BlockFPS: The input clip must have a frame rate. Invoke AssumeFPS if necessary.");
    int64_t numerator, denominator;
    if (num != 0 && den != 0) {
        numerator = num;
        denominator = den;
    } else {
        numerator = fps_num * 2;
        denominator = fps_den;
    }
    d->fa = denominator * fps_num;
    d->fb = numerator * fps_den;
    const int64_t g = gcd64(d->fa, d->fb);
    d->fa /= g;
    d->fb /= g;
    if (numerator <= 0 || denominator <= 0) {
        d->fps_num = 0;
        d->fps_den = 1;
    } else {
        const int64_t x = gcd64(numerator, denominator);
        d->fps_num = numerator / x;
        d->fps_den = denominator / x;
    }
    d->num_frames_in = num_frames;
    d->num_frames_out = (int)(1 + (num_frames - 1) * d->fb / d->fa);
    // :913-935
    if (vd_bw->nHeight != sp->height || vd_bw->nWidth != sp->super_width - sp->hpad * 2 || vd_bw->nWidth != clip_w || vd_bw->nHeight != clip_h || vd_bw->nPel != sp->pel)
        return set_error("BlockFPS: wrong source or super clip frame size.");
    if (sp->bits > 16 || sp->ssw > 1 || sp->ssh > 1)
        return set_error("BlockFPS: input clip must be GRAY, 420, 422, 440, or 444, up to 16 bits, with constant dimensions.");
    // :938-956
    const mvb200_analysis_data *vd = vd_bw;
    d->nBlkXP = vd->nBlkX;
    d->nBlkYP = vd->nBlkY;
    while (d->nBlkXP * (vd->nBlkSizeX - vd->nOverlapX) + vd->nOverlapX < vd->nWidth)
        d->nBlkXP++;
    while (d->nBlkYP * (vd->nBlkSizeY - vd->nOverlapY) + vd->nOverlapY < vd->nHeight)
        d->nBlkYP++;
    d->nWidthP = d->nBlkXP * (vd->nBlkSizeX - vd->nOverlapX) + vd->nOverlapX;
    d->nHeightP = d->nBlkYP * (vd->nBlkSizeY - vd->nOverlapY) + vd->nOverlapY;
    d->nWidthPUV = d->nWidthP / vd->xRatioUV;
    d->nHeightPUV = d->nHeightP / vd->yRatioUV;
    d->nHeightUV = vd->nHeight / vd->yRatioUV;
    d->nWidthUV = vd->nWidth / vd->xRatioUV;
    d->nPitchY = (d->nWidthP + 15) & (~15);
    d->nPitchUV = (d->nWidthPUV + 15) & (~15);
    d->bits = sp->bits;
    d->ssw = sp->ssw;
    d->ssh = sp->ssh;
    d->gray = sp->gray;
    d->nplanes = sp->nplanes;
    d->super_hpad = sp->hpad;
    d->super_vpad = sp->vpad;
    d->super_pel = sp->pel;
    d->super_modeyuv = sp->modeyuv;
    d->super_levels = sp->levels;
    return 0;
}

void mvb200_blockfps_position(const mvb200_blockfps_params *p, int n, int32_t *nleft, int32_t *time256) {
    const int off = p->vd.nDeltaFrame;
    *nleft = (int)(n * p->fa / p->fb);
    int t = (int)(((double)n * p->fa / p->fb - *nleft) * 256 + 0.5);
    if (off > 1)
        t = t / off;
    *time256 = t;
}

// InitTables, SimpleResize.cpp:26-56 (single-precision arithmetic, evaluated in the reference's order)
void mvb200_simple_resize_tables(int out, int in, int32_t *offsets, int32_t *weights) {
    const float leftmost = 0.5f, rightmost = in - 0.5f;
    const int leftmost_idx = (int)leftmost > 0 ? (int)leftmost : 0;
    const int rightmost_idx = (int)rightmost < in - 1 ? (int)rightmost : in - 1;
    for (int i = 0; i < out; i++) {
        volatile float position = (i + 0.5f) * (float)in; // (volatile: no contraction / excess precision on any host)
        position = position / (float)out;
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
            volatile float t = position - leftmost;
            weight = t - offset;
        }
        offsets[i] = offset;
        weights[i] = (int)(weight * (1 << 14));
    }
}

static long long round_up256(long long v) { return (v + 255) / 256 * 256; }

void mvb200_mask_layout(const mvb200_mask_params *p, mvb200_frame_layout *l) {
    memset(l, 0, sizeof(*l));
    l->nplanes = 3;
    l->bytes_per_sample = 1;
    long long off = 0;
    for (int i = 0; i < 3; i++) {
        l->plane_w[i] = i ? p->nWidthUV : p->nWidth;
        l->plane_h[i] = i ? p->nHeightUV : p->nHeight;
        l->pitch[i] = round_up256(l->plane_w[i]);
        l->offset[i] = off;
        off += l->pitch[i] * l->plane_h[i];
    }
    l->frame_bytes = round_up256(off);
}

// mvfinestCreate, MVFinest.c:190-194: (W + 2 hpad) pel x (H + 2 vpad) pel; chroma planes by the subsampling
void mvb200_finest_layout(const mvb200_super_params *sp, mvb200_frame_layout *l) {
    memset(l, 0, sizeof(*l));
    l->nplanes = sp->nplanes;
    l->bytes_per_sample = (sp->bits + 7) / 8;
    const int w = (sp->width + 2 * sp->hpad) * sp->pel, h = (sp->height + 2 * sp->vpad) * sp->pel;
    long long off = 0;
    for (int i = 0; i < sp->nplanes; i++) {
        l->plane_w[i] = i ? w >> sp->ssw : w;
        l->plane_h[i] = i ? h >> sp->ssh : h;
        l->pitch[i] = round_up256((long long)l->plane_w[i] * l->bytes_per_sample);
        l->offset[i] = off;
        off += l->pitch[i] * l->plane_h[i];
    }
    l->frame_bytes = round_up256(off);
}

} // extern "C"

