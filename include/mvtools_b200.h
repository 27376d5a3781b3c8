/*
 * mvtools_b200.h -- C ABI of the B200 (sm_100a) engine behind the mv.* hot path.
 *
 * One shared library, libmvtools_b200.so, exports exactly the functions declared
 * here: plain C, POD structs, raw pointers and sizes -- no C++ and no torch types
 * cross this boundary.  Each entry point names the reference interface it
 * replaces (file:line under dubhater/vapoursynth-mvtools @ e516e90).
 *
 *   reference (CPU)                                   this ABI (GPU)
 *   ------------------------------------------------  -----------------------------
 *   mvsuperCreate      MVSuper.c:140-275              mvb200_super_init
 *   mvsuperGetFrame    MVSuper.c:43-126               mvb200_super_frame / _batch
 *   mvanalyseCreate    MVAnalyse.c:267-635            mvb200_analyse_init
 *   mvanalyseGetFrame  MVAnalyse.c:76-254             mvb200_analyse_frame / _batch
 *     gopSearchMVs     GroupOfPlanes.c:69-125           (inside the search kernel)
 *     doPobSearchMVs   PlaneOfBlocks.cpp:971-1131       (inside the search kernel)
 *   mvdegrainCreate<r> MVDegrains.cpp:512-809         mvb200_degrain_init
 *   mvdegrainGetFrame  MVDegrains.cpp:85-330          mvb200_degrain_frame / _batch
 *   mvcompensateCreate/GetFrame MVCompensate.c:73-575 mvb200_compensate_init / _frame
 *   mvrecalculateCreate/GetFrame MVRecalculate.c:68-553 mvb200_recalculate_init / _frame
 *   -- consumers of the vector clips --
 *   mvscdetectionCreate/GetFrame MVSCDetection.c:43-131 mvb200_scdetection_init / _frame / _batch
 *   mvmaskCreate/GetFrame        MVMask.c:74-344        mvb200_mask_init / _frame / _batch
 *   mvfinestCreate/GetFrame      MVFinest.c:47-207      mvb200_finest_layout / _frame / _batch
 *   mvflowCreate/GetFrame        MVFlow.cpp:157-593     mvb200_flow_init / _frame / _batch
 *   mvflowinterCreate/GetFrame   MVFlowInter.c:83-690   mvb200_flowinter_init / _frame / _batch, mvb200_blend_frame
 *   mvflowfpsCreate/GetFrame     MVFlowFPS.c:86-928     mvb200_flowfps_init / _position + mvb200_flowinter_frame
 *   mvflowblurCreate/GetFrame    MVFlowBlur.c:144-545   mvb200_flowblur_init / _frame / _batch
 *   mvblockfpsCreate/GetFrame    MVBlockFPS.c:240-1015  mvb200_blockfps_init / _position / _frame
 *
 * The VapourSynth plugin (vapoursynth-mvtools_b200/plugin, exporting
 * VapourSynthPluginInit2 like EntryPoint.c:28-50) is a thin host-side glue over
 * these functions; INTEGRATION.md shows the binding a maintainer of the
 * reference would add to call them from the existing filters.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; the message is
 *     retrievable with mvb200_last_error() (thread-local) and, for the *_init
 *     functions, is byte-identical to the reference's mapSetError text;
 *   - "host" pointers are ordinary CPU memory (pinned or not); "dev" pointers are
 *     CUDA device memory on the context's device;
 *   - there is NO CPU fallback: without a usable CUDA device every compute entry
 *     point fails loudly.
 */
#ifndef MVTOOLS_B200_H
#define MVTOOLS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MVB200_API __attribute__((visibility("default")))
#else
#define MVB200_API
#endif

/* "argument not given" marker for every optional filter argument below */
#define MVB200_UNSET INT32_MIN
#define MVB200_UNSET64 INT64_MIN

#define MVB200_MAX_LEVELS 32
#define MVB200_MAX_RADIUS 6

/* MVPlaneSet, MVFrame.h:23-28 */
enum { MVB200_YPLANE = 1, MVB200_UPLANE = 2, MVB200_VPLANE = 4, MVB200_YUVPLANES = 7 };

/* VECTOR, MVAnalysisData.h:40-44 -- 16 bytes, the unit of the MVTools_vectors blob */
typedef struct mvb200_vector {
    int32_t x;
    int32_t y;
    int64_t sad;
} mvb200_vector;

/* MVAnalysisData, MVAnalysisData.h:83-134 -- the 84-byte MVTools_MVAnalysisData prop */
typedef struct mvb200_analysis_data {
    int32_t nMagicKey, nVersion;
    int32_t nBlkSizeX, nBlkSizeY, nPel, nLvCount, nDeltaFrame, isBackward, nCPUFlags, nMotionFlags;
    int32_t nWidth, nHeight, nOverlapX, nOverlapY, nBlkX, nBlkY, bitsPerSample, yRatioUV, xRatioUV;
    int32_t nHPadding, nVPadding;
} mvb200_analysis_data;

/* ------------------------------------------------------------------ Super */

/* Arguments of mv.Super (MVSuper.c:279-291) + the input clip's format.
 * Set the "in" fields (MVB200_UNSET for "argument not given"), call mvb200_super_init. */
typedef struct mvb200_super_params {
    /* in: clip format */
    int32_t width, height, bits;  /* bits 8..16 */
    int32_t ssw, ssh;             /* log2 chroma subsampling (0 or 1) */
    int32_t gray;                 /* 1 = single plane */
    /* in: filter arguments, MVB200_UNSET = default (MVSuper.c:149-177) */
    int32_t hpad, vpad, pel, levels, chroma, sharp, rfilter;
    /* out (derived): */
    int32_t nplanes;              /* 1 or 3 */
    int32_t xRatioUV, yRatioUV;
    int32_t modeyuv;              /* Super_modeyuv */
    int32_t super_width, super_height; /* luma size of the super frame (MVSuper.c:254-262) */
} mvb200_super_params;

/* Geometry of one (plane, level) inside a super frame.  All rows of a plane share
 * one pitch; `row0` is the first row of the level's sub-plane 0 inside the plane
 * (PlaneSuperOffset, MVFrame.cpp:1229-1247, divided by the pitch). */
typedef struct mvb200_level_geom {
    int32_t width, height;        /* interior */
    int32_t hpad, vpad;
    int32_t padded_w, padded_h;
    int32_t pel;                  /* pel of this level (level>0: 1) */
    int32_t row0;
} mvb200_level_geom;

typedef struct mvb200_super_geom {
    int32_t nplanes, levels;
    int32_t plane_w[3], plane_h[3];            /* super-frame plane sizes in samples/rows */
    mvb200_level_geom lv[3][MVB200_MAX_LEVELS];
} mvb200_super_geom;

/* ---------------------------------------------------------------- Analyse */

/* Arguments of mv.Analyse (MVAnalyse.c:639-671).  MVB200_UNSET = not given. */
typedef struct mvb200_analyse_args {
    int32_t blksize, blksizev, levels, search, searchparam, pelsearch, isb, lambda, chroma, delta,
        truemotion, lsad, plevel, global, pnew, pzero, pglobal, overlap, overlapv, divide, badsad,
        badrange, opt, meander, trymany, fields, tff, search_coarse, dct;
} mvb200_analyse_args;

/* Fully resolved search parameters (what MVAnalyseData holds after mvanalyseCreate). */
typedef struct mvb200_analyse_params {
    mvb200_analysis_data ad;      /* the prop written on every vector frame */
    int32_t divide;               /* divideExtra */
    mvb200_analysis_data ad_divided;
    int32_t search, search_coarse, nSearchParam, nPelSearch;
    int32_t nLambda, lsad, pnew, plevel, global, pglobal, pzero;
    int64_t badSAD;
    int32_t badrange, meander, tryMany, dctmode, chroma;
    int32_t fields, tff, tff_exists;
    /* super-clip facts (Super_* props) */
    int32_t super_hpad, super_vpad, super_pel, super_modeyuv, super_levels;
    int32_t ssw, ssh, gray;
    int32_t num_frames;
    /* level geometry of the search (gopInit, GroupOfPlanes.c:25-56) */
    int32_t nBlkX[MVB200_MAX_LEVELS], nBlkY[MVB200_MAX_LEVELS];
    int32_t vectors_size;         /* gopGetArraySize: bytes of the MVTools_vectors blob */
} mvb200_analyse_params;

/* ---------------------------------------------------------------- Degrain */

typedef struct mvb200_degrain_args {
    int32_t radius;               /* 1..6 (Degrain1..Degrain6) */
    int64_t thsad, thsadc;        /* MVB200_UNSET64 = default */
    int32_t plane;                /* MVB200_UNSET = default (4) */
    int32_t limit, limitc;        /* MVB200_UNSET = default */
    int64_t thscd1;               /* MVB200_UNSET64 = default */
    int32_t thscd2;               /* MVB200_UNSET = default */
} mvb200_degrain_args;

typedef struct mvb200_degrain_params {
    int32_t radius;
    mvb200_analysis_data vd[2 * MVB200_MAX_RADIUS]; /* per vector clip, order mvbw,mvfw,mvbw2,... */
    int64_t thSAD[3];
    int32_t nLimit[3];
    int64_t nSCD1;
    int32_t nSCD2;
    int32_t process[3];
    int32_t YUVplanes;
    int32_t bits, ssw, ssh, gray, nplanes;
    int32_t super_hpad, super_vpad, super_pel, super_modeyuv, super_levels;
    int32_t nWidth[3], nHeight[3], nOverlapX[3], nOverlapY[3], nBlkSizeX[3], nBlkSizeY[3], nWidth_B[3], nHeight_B[3];
} mvb200_degrain_params;

/* ------------------------------------------------------- Compensate / Recalculate */

typedef struct mvb200_compensate_args {
    int32_t scbehavior;           /* MVB200_UNSET = default (1) */
    int64_t thsad;                /* MVB200_UNSET64 = default (10000) */
    int32_t fields;               /* MVB200_UNSET = default (0) */
    double time;                  /* NaN = default (100.0) */
    int64_t thscd1;               /* MVB200_UNSET64 = default */
    int32_t thscd2;
    int32_t tff;                  /* MVB200_UNSET = not given */
} mvb200_compensate_args;

typedef struct mvb200_compensate_params {
    mvb200_analysis_data vd;
    int32_t scBehavior, fields, tff, tff_exists, time256;
    int64_t thSAD;
    int64_t nSCD1;
    int32_t nSCD2;
    int32_t bits, ssw, ssh, gray, nplanes;
    int32_t super_hpad, super_vpad, super_pel, super_modeyuv, super_levels;
    int32_t nWidth[3], nHeight[3], nOverlapX[3], nOverlapY[3], nBlkSizeX[3], nBlkSizeY[3], nWidth_B[3], nHeight_B[3];
    int32_t num_frames;
} mvb200_compensate_params;

typedef struct mvb200_recalculate_args {
    int64_t thsad;                /* MVB200_UNSET64 = default (200) */
    int32_t smooth, blksize, blksizev, search, searchparam, lambda, chroma, truemotion, pnew, overlap,
        overlapv, divide, opt, meander, fields, tff, dct;
} mvb200_recalculate_args;

typedef struct mvb200_recalculate_params {
    mvb200_analysis_data ad;      /* new analysis data (written as the prop) */
    mvb200_analysis_data ad_divided;
    mvb200_analysis_data old;     /* analysis data of the input vector clip */
    int32_t divide, search, nSearchParam, nLambda, pnew, smooth, meander, dctmode, chroma;
    int64_t thSAD;
    int32_t fields, tff, tff_exists;
    int32_t super_hpad, super_vpad, super_pel, super_modeyuv, super_levels;
    int32_t ssw, ssh, gray;
    int32_t num_frames;
    int32_t vectors_size;
} mvb200_recalculate_params;


/* ------------------------------------------ SCDetection / Mask / Finest / Flow */
/* Consumers of the vector clips.  Float arguments: NaN = "not given". */

typedef struct mvb200_scdetection_params {
    mvb200_analysis_data vd;
    int64_t nSCD1;                /* scaled (scaleThSCD, MVAnalysisData.c:7-31) */
    int32_t nSCD2;
} mvb200_scdetection_params;

typedef struct mvb200_mask_args {
    double ml, gamma;             /* NaN = default (100.0, 1.0) */
    int32_t kind;                 /* MVB200_UNSET = default (0) */
    double time;                  /* NaN = default (100.0) */
    int32_t ysc;                  /* MVB200_UNSET = default (0) */
    int64_t thscd1;               /* MVB200_UNSET64 = default */
    int32_t thscd2;               /* MVB200_UNSET = default */
} mvb200_mask_args;

typedef struct mvb200_mask_params {
    mvb200_analysis_data vd;
    int32_t kind, time256, nSceneChangeValue;
    float ml, fGamma, fMaskNormFactor, fMaskNormFactor2, fHalfGamma; /* MVMask.c:299-302 (float arithmetic) */
    int64_t nSCD1;
    int32_t nSCD2;
    int32_t clip_gray;            /* the input clip was GRAY8: the output is YUV444P8 (MVMask.c:323-324) */
    int32_t nWidth, nHeight, nWidthUV, nHeightUV, nWidthB, nHeightB, nWidthBUV, nHeightBUV;
} mvb200_mask_params;

typedef struct mvb200_flow_args {
    double time;                  /* NaN = default (100.0) */
    int32_t mode, fields;         /* MVB200_UNSET = default (0) */
    int64_t thscd1;
    int32_t thscd2;
    int32_t tff;                  /* MVB200_UNSET = not given */
} mvb200_flow_args;

typedef struct mvb200_flow_params {
    mvb200_analysis_data vd;
    int32_t time256, mode, fields, tff, tff_exists;
    int64_t nSCD1;
    int32_t nSCD2;
    int32_t bits, ssw, ssh, gray, nplanes;
    int32_t super_hpad, super_vpad, super_pel, super_levels;
    /* MVFlow.cpp:529-548 */
    int32_t nBlkXP, nBlkYP, nWidthP, nHeightP, nWidthPUV, nHeightPUV, nWidthUV, nHeightUV, nHPaddingUV, nVPaddingUV, VPitchY, VPitchUV;
    int32_t num_frames;
} mvb200_flow_params;

/* mv.FlowInter (MVFlowInter.c) and the per-frame core of mv.FlowFPS (MVFlowFPS.c): motion-compensated interpolation
 * between a "left" frame and the frame `delta` later, at time256 / 256 of the way. */
typedef struct mvb200_flowinter_args {
    double time, ml;              /* NaN = default (50.0, 100.0) */
    int32_t blend;                /* MVB200_UNSET = default (1) */
    int64_t thscd1;
    int32_t thscd2;
} mvb200_flowinter_args;

typedef struct mvb200_flowinter_params {
    mvb200_analysis_data vd;      /* of mvbw (mvfw is checked to be similar) */
    int32_t time256, blend;
    double ml;                    /* MVFlowInter.c keeps a float, MVFlowFPS.c a double: stored widened */
    int64_t nSCD1;
    int32_t nSCD2;
    /* which of FlowInterSimple (0) / FlowInter (1) / FlowInterExtra (2) runs when the extra vector pair (mvfw of the
     * left frame, mvbw of the right frame) is usable, and which otherwise.  FlowInter: 2 / 1 (MVFlowInter.c:285,357);
     * FlowFPS mask=0: 0 / 0, mask=1: 1 / 1, mask=2: 2 / 0 (MVFlowFPS.c:339-437) */
    int32_t mode_primary, mode_fallback;
    int32_t bits, ssw, ssh, gray, nplanes;
    int32_t super_hpad, super_vpad, super_pel, super_levels;
    int32_t nBlkXP, nBlkYP, nWidthP, nHeightP, nWidthPUV, nHeightPUV, nWidthUV, nHeightUV, nHPaddingUV, nVPaddingUV, VPitchY, VPitchUV;
    int32_t num_frames;
} mvb200_flowinter_params;

/* mv.FlowFPS (MVFlowFPS.c:563-928): frame-rate conversion; the per-frame work is mvb200_flowinter_frame */
typedef struct mvb200_flowfps_args {
    int64_t num, den;             /* MVB200_UNSET64 = default (25, 1) */
    int32_t mask;                 /* MVB200_UNSET = default (2) */
    double ml;                    /* NaN = default (100.0) */
    int32_t blend;
    int64_t thscd1;
    int32_t thscd2;
} mvb200_flowfps_args;

typedef struct mvb200_flowfps_params {
    mvb200_flowinter_params inter; /* time256 is per output frame (mvb200_flowfps_position) */
    int32_t maskmode;
    int64_t fa, fb;               /* output frame n lies at n fa / fb in the input clip (MVFlowFPS.c:747-751) */
    int64_t fps_num, fps_den;     /* of the output clip (setFPS, :547-561) */
    int32_t num_frames_in, num_frames_out;
} mvb200_flowfps_params;

/* mv.FlowBlur (MVFlowBlur.c): motion blur of frame n along its forward and backward vectors */
typedef struct mvb200_flowblur_args {
    double blur;                  /* NaN = default (50.0) */
    int32_t prec;                 /* MVB200_UNSET = default (1) */
    int64_t thscd1;
    int32_t thscd2;
} mvb200_flowblur_args;

typedef struct mvb200_flowblur_params {
    mvb200_analysis_data vd;      /* of mvbw */
    int32_t blur256, prec;
    int64_t nSCD1;
    int32_t nSCD2;
    int32_t bits, ssw, ssh, gray, nplanes;
    int32_t super_hpad, super_vpad, super_pel, super_levels;
    int32_t nWidthUV, nHeightUV, nHPaddingUV, nVPaddingUV;
    int32_t num_frames;
} mvb200_flowblur_params;

/* mv.BlockFPS (MVBlockFPS.c): block-based frame-rate conversion */
typedef struct mvb200_blockfps_args {
    int64_t num, den;             /* MVB200_UNSET64 = default (25, 1) */
    int32_t mode;                 /* MVB200_UNSET = default (3) */
    double ml;                    /* NaN = default (100.0) */
    int32_t blend;
    int64_t thscd1;
    int32_t thscd2;
} mvb200_blockfps_args;

typedef struct mvb200_blockfps_params {
    mvb200_analysis_data vd;      /* of mvbw */
    int32_t mode, blend;
    double ml;
    int64_t nSCD1;
    int32_t nSCD2;
    int32_t bits, ssw, ssh, gray, nplanes;
    int32_t super_hpad, super_vpad, super_pel, super_modeyuv, super_levels;
    int32_t nBlkXP, nBlkYP, nWidthP, nHeightP, nWidthPUV, nHeightPUV, nWidthUV, nHeightUV, nPitchY, nPitchUV;
    int64_t fa, fb, fps_num, fps_den;
    int32_t num_frames_in, num_frames_out;
} mvb200_blockfps_params;

/* ------------------------------------------------------------- host logic */
/* (pure host code: usable without a GPU; restates the reference's *Create) */

MVB200_API const char *mvb200_last_error(void);
MVB200_API const char *mvb200_version(void);

/* MVSuper.c:140-264.  Returns non-zero and sets the reference's message on bad input. */
MVB200_API int mvb200_super_init(mvb200_super_params *p);
MVB200_API void mvb200_super_geometry(const mvb200_super_params *p, mvb200_super_geom *g);

/* MVAnalyse.c:267-624.  `sp` describes the super clip (the Super_* props), num_frames its length. */
MVB200_API int mvb200_analyse_init(const mvb200_analyse_args *a, const mvb200_super_params *sp, int num_frames, mvb200_analyse_params *out);

/* MVDegrains.cpp:512-782.  vd[2*radius]: MVAnalysisData of mvbw,mvfw,mvbw2,...;
 * (clip_w, clip_h) the source clip size. */
MVB200_API int mvb200_degrain_init(const mvb200_degrain_args *a, const mvb200_super_params *sp, const mvb200_analysis_data *vd, int clip_w, int clip_h, mvb200_degrain_params *out);

/* MVCompensate.c:419-575 */
MVB200_API int mvb200_compensate_init(const mvb200_compensate_args *a, const mvb200_super_params *sp, const mvb200_analysis_data *vd, int clip_w, int clip_h, int num_frames, mvb200_compensate_params *out);

/* MVRecalculate.c:263-553 */
MVB200_API int mvb200_recalculate_init(const mvb200_recalculate_args *a, const mvb200_super_params *sp, const mvb200_analysis_data *old, int num_frames, mvb200_recalculate_params *out);


/* MVSCDetection.c:89-131.  thscd1 / thscd2: MVB200_UNSET64 / MVB200_UNSET = default. */
MVB200_API int mvb200_scdetection_init(int64_t thscd1, int32_t thscd2, const mvb200_analysis_data *vd, mvb200_scdetection_params *out);

/* MVMask.c:221-344.  The input clip's format: bits, log2 subsampling, gray. */
MVB200_API int mvb200_mask_init(const mvb200_mask_args *a, const mvb200_analysis_data *vd, int clip_bits, int clip_ssw, int clip_ssh, int clip_gray, mvb200_mask_params *out);

/* MVFlow.cpp:395-575.  `sp` describes the super clip, (clip_w, clip_h) the source clip. */
MVB200_API int mvb200_flow_init(const mvb200_flow_args *a, const mvb200_super_params *sp, const mvb200_analysis_data *vd, int num_frames, mvb200_flow_params *out);

/* MVFlowInter.c:458-690.  vd_bw / vd_fw: the analysis data of mvbw / mvfw. */
MVB200_API int mvb200_flowinter_init(const mvb200_flowinter_args *a, const mvb200_super_params *sp, const mvb200_analysis_data *vd_bw, const mvb200_analysis_data *vd_fw, int num_frames, mvb200_flowinter_params *out);

/* MVFlowFPS.c:563-800.  (fps_num, fps_den, clip_w, clip_h, num_frames) describe the input clip. */
MVB200_API int mvb200_flowfps_init(const mvb200_flowfps_args *a, const mvb200_super_params *sp, const mvb200_analysis_data *vd_bw, const mvb200_analysis_data *vd_fw, int64_t fps_num, int64_t fps_den, int clip_w, int clip_h,
                                   int num_frames, mvb200_flowfps_params *out);
/* MVFlowFPS.c:118-141: output frame n lies between input frames *nleft and *nleft + delta, *time256 / 256 of the way
 * (0 and 256 mean "that input frame itself") */
MVB200_API void mvb200_flowfps_position(const mvb200_flowfps_params *p, int n, int32_t *nleft, int32_t *time256);

/* MVFlowBlur.c:345-545 */
MVB200_API int mvb200_flowblur_init(const mvb200_flowblur_args *a, const mvb200_super_params *sp, const mvb200_analysis_data *vd_bw, const mvb200_analysis_data *vd_fw, int num_frames, mvb200_flowblur_params *out);

/* MVBlockFPS.c:736-990 */
MVB200_API int mvb200_blockfps_init(const mvb200_blockfps_args *a, const mvb200_super_params *sp, const mvb200_analysis_data *vd_bw, const mvb200_analysis_data *vd_fw, int64_t fps_num, int64_t fps_den, int clip_w, int clip_h,
                                    int num_frames, mvb200_blockfps_params *out);
MVB200_API void mvb200_blockfps_position(const mvb200_blockfps_params *p, int n, int32_t *nleft, int32_t *time256); /* MVBlockFPS.c:275-289 */

/* overInit, Overlap.cpp:40-125: the nine overlap windows (TL, TM, TR, ML, MM, MR, BL, BM, BR) of nx * ny int16 each,
 * built on the host (the reference's cosf(double) expression); the Degrain / Compensate / BlockFPS kernels use them */
MVB200_API void mvb200_overlap_windows(int nx, int ny, int ox, int oy, int16_t *out9);

/* InitTables, SimpleResize.cpp:26-56: offsets[out] and weights[out] (14-bit) of the bilinear resizer the Mask / Flow
 * filters upsize their per-block fields with (float arithmetic, restated on the host; uploaded once per configuration). */
MVB200_API void mvb200_simple_resize_tables(int out, int in, int32_t *offsets, int32_t *weights);

/* ---------------------------------------------------------------- context */

typedef struct mvb200_ctx mvb200_ctx;

/* Binds a context to CUDA device `device` (own stream, scratch pools).  Fails if
 * the device is missing or is not compute capability 10.x. */
MVB200_API int mvb200_ctx_create(int device, mvb200_ctx **out);
/* Number of CUDA devices this engine can run on (compute capability 10.x), for callers that spread independent frame
 * windows over every GPU of a box (the plugin's fused paths: one context per device, no data-path collective -- frames are
 * independent units, readme.rst:36).  0 when there is none (and mvb200_last_error says why). */
MVB200_API int mvb200_device_count(void);
/* total memory of a device in bytes (0 on failure): callers that keep several windows of a clip resident size them with it */
MVB200_API int64_t mvb200_device_memory(int device);
MVB200_API void mvb200_ctx_destroy(mvb200_ctx *ctx);
MVB200_API int mvb200_ctx_synchronize(mvb200_ctx *ctx);
/* the cudaStream_t the context launches on (for CUDA-event timing by the caller) */
MVB200_API void *mvb200_ctx_stream(mvb200_ctx *ctx);
/* number of kernels this context has launched so far */
MVB200_API int64_t mvb200_ctx_launch_count(mvb200_ctx *ctx);

/* Completion marker for the asynchronous entry points (mvb200_degrain_clip_host_async): mvb200_ctx_record() marks
 * everything enqueued on the context so far (its compute stream and both copy streams).  The reference has no
 * counterpart -- its GetFrame functions are synchronous (MVDegrains.cpp:85-330); this is what lets the plugin keep the next
 * window of frames in flight on the GPU while the frame server's threads still collect the current one.
 * mvb200_event_wait blocks the calling thread until that work is done; mvb200_event_query returns 0 when done, 1 while
 * pending; both may be called from any thread without holding whatever lock serialises the context's other calls. */
typedef struct mvb200_event mvb200_event;
MVB200_API int mvb200_ctx_record(mvb200_ctx *ctx, mvb200_event **out);
MVB200_API int mvb200_event_wait(mvb200_event *ev);
MVB200_API int mvb200_event_query(mvb200_event *ev);
MVB200_API void mvb200_event_destroy(mvb200_event *ev);

/* Which search kernel a mv.Analyse configuration runs on and how many of its serial search chains (one per frame and
 * vector clip: doPobSearchMVs walks a plane's blocks strictly in order, PlaneOfBlocks.cpp:971-1131) fill the GPU once.
 * kernel: 0 = generic (every option), 1 = register-blocked, 2 = TMA-staged windows (the production kernel).  Pure host
 * code: sm_count <= 0 assumes a B200 (148 SMs).  Callers size their batches with it: frames per launch =
 * chains_per_wave / number of vector clips. */
typedef struct mvb200_search_plan {
    int32_t kernel;
    int32_t chains_per_sm;
    int32_t chains_per_wave;
    int32_t smem_per_chain;
} mvb200_search_plan;
MVB200_API int mvb200_search_plan_query(const mvb200_analyse_params *ap, const mvb200_super_params *sp, int sm_count, mvb200_search_plan *out);

/* How a caller that streams windows of frames through the *_clip_host entry points should size them (the plugin's fused paths and
 * bench.py share this).  A window of nvec vector clips per frame that fills the search kernel's chain slots once has
 * chains_per_wave / nvec frames (at most 768); two such waves in flight per device are what pays.  The windows' device buffers
 * (source / output clips double buffered, super clip, vectors) must fit 60 % of device_memory and the pinned host windows (one per
 * context, one being served, one recycled; src_frame_bytes + out_frame_bytes per frame) pinned_budget (<= 0: an eighth of the
 * machine's memory, between 2 and 32 GiB).  lookahead > 0: the window length is given.  contexts_per_device > 0: the number of
 * engine contexts per device is given (reduced only when the device cannot hold them); <= 0: chosen here, 1 to 4 -- a configuration
 * whose window cannot be a whole wave (4K frames: the window is what the pinned budget allows) gets more, smaller windows on more
 * contexts, since the chains in flight are contexts x window. */
typedef struct mvb200_window_plan {
    int32_t frames;               /* frames per window */
    int32_t contexts_per_device;  /* engine contexts to create on each device */
    int64_t pinned_budget;        /* the budget that was applied */
} mvb200_window_plan;
MVB200_API int mvb200_window_plan_query(const mvb200_analyse_params *ap, const mvb200_super_params *sp, int nvec, int lookahead, int64_t src_frame_bytes, int64_t out_frame_bytes, int ndevices,
                                        int contexts_per_device, int64_t pinned_budget, int64_t device_memory, mvb200_window_plan *out);

/* Device frame layout used by the *_batch functions: planes are stored one after
 * another, each with its own pitch (bytes), frames one after another.
 * pitch[p] = round_up(plane_w[p]*bytes, 256); offsets are from the frame start. */
typedef struct mvb200_frame_layout {
    int32_t nplanes, bytes_per_sample;
    int32_t plane_w[3], plane_h[3];
    int64_t pitch[3], offset[3];
    int64_t frame_bytes;
} mvb200_frame_layout;

MVB200_API void mvb200_clip_layout(const mvb200_super_params *p, mvb200_frame_layout *l);  /* source / output frames */
MVB200_API void mvb200_super_layout(const mvb200_super_params *p, mvb200_frame_layout *l); /* super frames */

/* ------------------------------------------- device-resident batch entry points */
/* n frames, contiguous in device memory in the layouts above. */

/* mvsuperGetFrame (MVSuper.c:43-126) for frames [0,n). */
MVB200_API int mvb200_super_batch(mvb200_ctx *ctx, const mvb200_super_params *p, const void *dev_src, void *dev_super, int n);

/* mvanalyseGetFrame (MVAnalyse.c:76-254) for frames [first, first+n) of a clip whose
 * super frames [0,num_frames) are resident at dev_super.  dev_vectors receives n
 * blobs of ap->vectors_size bytes each (stride vectors_stride bytes, >= size). */
MVB200_API int mvb200_analyse_batch(mvb200_ctx *ctx, const mvb200_analyse_params *ap, const mvb200_super_params *sp, const void *dev_super, int first, int n, void *dev_vectors, int64_t vectors_stride);

/* Several mv.Analyse instances over the same super clip in ONE launch (the
 * 2*radius vector clips a DegrainN consumes): nap parameter sets, dev_vectors[i]
 * receives the n blobs of instance i. */
MVB200_API int mvb200_analyse_multi_batch(mvb200_ctx *ctx, const mvb200_analyse_params *const *ap, int nap, const mvb200_super_params *sp, const void *dev_super, int first, int n, void *const *dev_vectors, int64_t vectors_stride);

/* mvdegrainGetFrame<radius> (MVDegrains.cpp:85-330) for frames [first, first+n).
 * dev_vectors[r] = blobs of vector clip r for frames [first, first+n) (same stride). */
MVB200_API int mvb200_degrain_batch(mvb200_ctx *ctx, const mvb200_degrain_params *dp, const mvb200_super_params *sp, const void *dev_src, const void *dev_super, int num_frames, int first, int n, const void *const *dev_vectors, int64_t vectors_stride, void *dev_dst);

/* mv.Recalculate / mv.Compensate over frames [first, first + n) of a device-resident super clip (gopRecalculateMVs per
 * frame, MVRecalculate.c:68-254; mvcompensateGetFrame, MVCompensate.c:73-374).  Old blobs are old_stride bytes apart, new
 * ones vectors_stride (>= rp->vectors_size); compensated frames are written in the clip layout, frame_bytes apart.
 * Reference frames follow the filters' own rule (n +/- delta, or the fixed frame -delta), an out-of-range or invalid
 * one gives the default blob / the scene-change copy; field shifts need an explicit tff (frame props are the plugin's). */
MVB200_API int mvb200_recalculate_batch(mvb200_ctx *ctx, const mvb200_recalculate_params *rp, const mvb200_super_params *sp, const void *dev_super, int first, int n, const void *dev_old_vectors, int64_t old_stride, void *dev_vectors, int64_t vectors_stride);
MVB200_API int mvb200_compensate_batch(mvb200_ctx *ctx, const mvb200_compensate_params *cp, const mvb200_super_params *sp, const void *dev_super, int first, int n, const void *dev_vectors, int64_t vectors_stride, void *dev_dst);

/* Host-buffer clip pipeline: mv.Super + nap x mv.Analyse + mv.DegrainN (nap == 2 * radius) over a whole
 * segment of n frames held in HOST memory in the clip layout (mvb200_clip_layout; frames frame_bytes
 * apart; use pinned memory for truly asynchronous copies); frames [first, first + count) are denoised into
 * host_dst (count frames), the rest of the segment only serves as references (the halo of a sharded clip:
 * vapoursynth-mvtools_b200/sharding.py).  Replaces, for a segment, the
 * mvsuperGetFrame / mvanalyseGetFrame / mvdegrainGetFrame calls a frame server would make per frame
 * (MVSuper.c:43-126, MVAnalyse.c:76-254, MVDegrains.cpp:85-330).  ASYNCHRONOUS: returns once the copies
 * and kernels are enqueued; host_dst is valid after mvb200_ctx_synchronize().  Consecutive calls
 * overlap (H2D of segment k+1 and D2H of segment k-1 beside the kernels of segment k). */
MVB200_API int mvb200_degrain_clip_host_async(mvb200_ctx *ctx, const mvb200_super_params *sp, const mvb200_analyse_params *const *ap, int nap,
                                              const mvb200_degrain_params *dp, const void *host_src, void *host_dst, int n, int first, int count);

/* The same clip pipeline with mv.Recalculate between the searches and the blend (the refinemotion shape of denoising
 * scripts: Analyse -> Recalculate (-> Recalculate) -> DegrainN): rp[i * nrp + k], k = 0 .. nrp - 1 (NULL = none), are the
 * refinement passes of vector clip i, applied in order (mvrecalculateGetFrame, MVRecalculate.c:68-254); dp was derived
 * from the analysis data of the LAST pass of each clip.  rp == NULL or nrp == 0: exactly mvb200_degrain_clip_host_async. */
MVB200_API int mvb200_degrain_recalc_clip_host_async(mvb200_ctx *ctx, const mvb200_super_params *sp, const mvb200_analyse_params *const *ap, int nap,
                                                     const mvb200_recalculate_params *const *rp, int nrp, const mvb200_degrain_params *dp, const void *host_src, void *host_dst,
                                                     int n, int first, int count);

/* Page-locked host memory for the buffers handed to mvb200_degrain_clip_host_async (pageable memory works too, but its
 * copies are not asynchronous).  mvb200_pinned_alloc returns NULL on failure. */
MVB200_API void *mvb200_pinned_alloc(size_t bytes);
MVB200_API void mvb200_pinned_free(void *p);

/* --------------------------------------------- host-buffer per-frame entry points */
/* (what the VapourSynth plugin calls from GetFrame; H2D/D2H copies inside) */

MVB200_API int mvb200_super_frame(mvb200_ctx *ctx, const mvb200_super_params *p, const uint8_t *const src[3], const ptrdiff_t src_stride[3], uint8_t *const dst[3], const ptrdiff_t dst_stride[3]);
/* mv.Super's `pelclip` (MVSuper.c:229-253 create-time checks; :91-104 + mvpRefineExt, MVFrame.cpp:1530-1631 per frame):
 * the check takes the pel clip's dimensions (same_format: it has the input clip's format and constant dimensions) and
 * says whether the clip is used (pel >= 2) and whether it already carries the padding; the frame call takes the pel
 * clip's frame planes beside the source frame's. */
MVB200_API int mvb200_super_pelclip_check(const mvb200_super_params *p, int same_format, int pel_width, int pel_height, int *use_pelclip, int *pel_padded);
MVB200_API int mvb200_super_frame_pel(mvb200_ctx *ctx, const mvb200_super_params *p, const uint8_t *const src[3], const ptrdiff_t src_stride[3], const uint8_t *const pel[3], const ptrdiff_t pel_stride[3], int pel_padded, uint8_t *const dst[3], const ptrdiff_t dst_stride[3]);

/* src_super/ref_super: host super frames (ref may be NULL when nref is out of range:
 * the default blob of gopWriteDefaultToArray, GroupOfPlanes.c:150-164, is written). */
MVB200_API int mvb200_analyse_frame(mvb200_ctx *ctx, const mvb200_analyse_params *ap, const mvb200_super_params *sp, const uint8_t *const src_super[3], const ptrdiff_t src_stride[3], const uint8_t *const ref_super[3], const ptrdiff_t ref_stride[3], int field_shift, uint8_t *vectors_out);

/* Several frames of ONE mv.Analyse instance in one call (the plugin's look-ahead: a search is a serial chain, so the
 * only way to use the GPU from a frame server is to have many frames' chains in flight).  `nframes` distinct host super
 * frames are uploaded once (frames[3 i + p] / strides[3 i + p] = plane p of frame i); job k searches frame src_index[k]
 * against frame ref_index[k] (< 0: no reference, the default blob of GroupOfPlanes.c:150-164) with field_shift[k], and
 * writes ap->vectors_size bytes to vectors_out[k].  One launch, njobs chains. */
MVB200_API int mvb200_analyse_frames(mvb200_ctx *ctx, const mvb200_analyse_params *ap, const mvb200_super_params *sp, int nframes, const uint8_t *const *frames, const ptrdiff_t *strides, int njobs, const int32_t *src_index,
                                     const int32_t *ref_index, const int32_t *field_shift, uint8_t *const *vectors_out);

/* refs[r] may be NULL (frame out of range); vectors[r] are host blobs. */
MVB200_API int mvb200_degrain_frame(mvb200_ctx *ctx, const mvb200_degrain_params *dp, const mvb200_super_params *sp, const uint8_t *const src[3], const ptrdiff_t src_stride[3], const uint8_t *const *const ref_super /* [2r][3] flattened */, const ptrdiff_t *ref_stride /* [2r][3] */, const uint8_t *const *vectors /* [2r] */, uint8_t *const dst[3], const ptrdiff_t dst_stride[3]);

MVB200_API int mvb200_compensate_frame(mvb200_ctx *ctx, const mvb200_compensate_params *cp, const mvb200_super_params *sp, const uint8_t *const src_super[3], const ptrdiff_t src_stride[3], const uint8_t *const ref_super[3], const ptrdiff_t ref_stride[3], const uint8_t *vectors, int field_shift, uint8_t *const dst[3], const ptrdiff_t dst_stride[3]);

MVB200_API int mvb200_recalculate_frame(mvb200_ctx *ctx, const mvb200_recalculate_params *rp, const mvb200_super_params *sp, const uint8_t *const src_super[3], const ptrdiff_t src_stride[3], const uint8_t *const ref_super[3], const ptrdiff_t ref_stride[3], const uint8_t *old_vectors, int field_shift, uint8_t *vectors_out);


/* ------------------------------------- SCDetection / Mask / Finest / Flow entry points */

/* mvscdetectionGetFrame (MVSCDetection.c:43-74): scene_change[i] = !fgopIsUsable(blob i) for n blobs (host or device
 * blobs; the flags always land in HOST memory).  The plugin writes the flag as _SceneChangePrev / _SceneChangeNext. */
MVB200_API int mvb200_scdetection_frame(mvb200_ctx *ctx, const mvb200_scdetection_params *p, const uint8_t *vectors, int32_t *scene_change);
MVB200_API int mvb200_scdetection_batch(mvb200_ctx *ctx, const mvb200_scdetection_params *p, const void *dev_vectors, int64_t vectors_stride, int n, int32_t *scene_change);

/* mvmaskGetFrame (MVMask.c:74-203).  The output is 8-bit, three planes (plane sizes nWidth x nHeight and
 * nWidthUV x nHeightUV of the params); src_luma is only read for kind = 5 (luma is passed through). */
MVB200_API int mvb200_mask_frame(mvb200_ctx *ctx, const mvb200_mask_params *p, const uint8_t *src_luma, ptrdiff_t src_stride, const uint8_t *vectors, uint8_t *const dst[3], const ptrdiff_t dst_stride[3]);
/* device resident: n blobs -> n masks; planes back to back with pitch round_up(w, 256) (mvb200_mask_layout);
 * dev_src_luma (kind 5 only): n luma planes, src_pitch / src_frame_bytes apart. */
MVB200_API void mvb200_mask_layout(const mvb200_mask_params *p, mvb200_frame_layout *l);
MVB200_API int mvb200_mask_batch(mvb200_ctx *ctx, const mvb200_mask_params *p, const void *dev_src_luma, int64_t src_pitch, int64_t src_frame_bytes, const void *dev_vectors, int64_t vectors_stride, int n, void *dev_dst);

/* mvfinestGetFrame (MVFinest.c:47-125): the pel^2 sub-planes of level 0 interleaved into one plane of
 * (W + 2 hpad) pel x (H + 2 vpad) pel samples per plane (pel 1: level 0 as it is). */
MVB200_API void mvb200_finest_layout(const mvb200_super_params *sp, mvb200_frame_layout *l);
MVB200_API int mvb200_finest_frame(mvb200_ctx *ctx, const mvb200_super_params *sp, const uint8_t *const super[3], const ptrdiff_t super_stride[3], uint8_t *const dst[3], const ptrdiff_t dst_stride[3]);
MVB200_API int mvb200_finest_batch(mvb200_ctx *ctx, const mvb200_super_params *sp, const void *dev_super, int n, void *dev_dst);

/* mvflowGetFrame (MVFlow.cpp:157-368): per-pixel motion compensation of the reference frame by the upsized vector
 * field.  The engine reads the pel sub-planes of the reference SUPER frame directly (the reference materialises them
 * as an mv.Finest clip first; same samples).  *usable = 0 (scene change, invalid vectors or ref_super == NULL) leaves
 * dst untouched: the reference returns the source frame itself then (MVFlow.cpp:362-366). */
MVB200_API int mvb200_flow_frame(mvb200_ctx *ctx, const mvb200_flow_params *fp, const mvb200_super_params *sp, const uint8_t *const ref_super[3], const ptrdiff_t ref_stride[3], const uint8_t *vectors, int field_shift, uint8_t *const dst[3], const ptrdiff_t dst_stride[3], int32_t *usable);
/* device resident, frames [first, first + n): unusable frames are copies of the source clip's frame (dev_src, clip layout). */
MVB200_API int mvb200_flow_batch(mvb200_ctx *ctx, const mvb200_flow_params *fp, const mvb200_super_params *sp, const void *dev_src, const void *dev_super, int first, int n, const void *dev_vectors, int64_t vectors_stride, void *dev_dst);

/* mvflowinterGetFrame (MVFlowInter.c:83-448) / the core of mvflowfpsGetFrame (MVFlowFPS.c:118-470) for one output
 * frame: left_super / right_super are the super frames of frames nleft and nleft + delta; vec_b = mvbw of nleft,
 * vec_f = mvfw of nleft + delta, vec_bb = mvbw of nleft + delta and vec_ff = mvfw of nleft (the "extra" pair, may be
 * NULL when the mode never uses it).  *usable = 0 (scene change / invalid vectors): dst is left untouched and the
 * caller blends the two SOURCE frames (mvb200_blend_frame) or returns the left one, as the reference does. */
MVB200_API int mvb200_flowinter_frame(mvb200_ctx *ctx, const mvb200_flowinter_params *p, const mvb200_super_params *sp, int time256, const uint8_t *const left_super[3], const ptrdiff_t left_stride[3],
                                      const uint8_t *const right_super[3], const ptrdiff_t right_stride[3], const uint8_t *vec_b, const uint8_t *vec_f, const uint8_t *vec_bb, const uint8_t *vec_ff,
                                      uint8_t *const dst[3], const ptrdiff_t dst_stride[3], int32_t *usable);
/* Blend (MaskFun.cpp:349-373): dst = (src (256 - time256) + ref time256) >> 8 on frames of the clip `sp` describes */
MVB200_API int mvb200_blend_frame(mvb200_ctx *ctx, const mvb200_super_params *sp, int time256, const uint8_t *const src[3], const ptrdiff_t src_stride[3], const uint8_t *const ref[3], const ptrdiff_t ref_stride[3],
                                  uint8_t *const dst[3], const ptrdiff_t dst_stride[3]);
/* mv.FlowInter over frames [first, first + n) of a device-resident clip: dev_bw / dev_fw hold the blobs of ALL
 * num_frames frames (frame f's blob at f * vectors_stride); unusable frames are blended (or copied) from dev_src. */
MVB200_API int mvb200_flowinter_batch(mvb200_ctx *ctx, const mvb200_flowinter_params *p, const mvb200_super_params *sp, const void *dev_src, const void *dev_super, int first, int n, const void *dev_bw, const void *dev_fw,
                                      int64_t vectors_stride, void *dev_dst);

/* mvflowblurGetFrame (MVFlowBlur.c:144-337): super = the super frame of frame n, vec_b = mvbw of frame n - delta,
 * vec_f = mvfw of frame n + delta (NULL when out of range).  *usable = 0: dst untouched, the caller returns the source. */
MVB200_API int mvb200_flowblur_frame(mvb200_ctx *ctx, const mvb200_flowblur_params *p, const mvb200_super_params *sp, const uint8_t *const super[3], const ptrdiff_t super_stride[3], const uint8_t *vec_b, const uint8_t *vec_f,
                                     uint8_t *const dst[3], const ptrdiff_t dst_stride[3], int32_t *usable);
/* frames [first, first + n) of a device-resident clip (blobs of all frames, frame f's at f * vectors_stride);
 * unusable frames are copies of the source clip's frames */
MVB200_API int mvb200_flowblur_batch(mvb200_ctx *ctx, const mvb200_flowblur_params *p, const mvb200_super_params *sp, const void *dev_src, const void *dev_super, int first, int n, const void *dev_bw, const void *dev_fw,
                                     int64_t vectors_stride, void *dev_dst);

/* mvblockfpsGetFrame (MVBlockFPS.c:240-640) for one output frame: left_super / right_super are the super frames of
 * frames nleft and nleft + delta, vec_b = mvbw of nleft, vec_f = mvfw of nleft + delta.  *usable = 0: dst untouched
 * (the caller blends the source frames with mvb200_blend_frame, or returns the left one). */
MVB200_API int mvb200_blockfps_frame(mvb200_ctx *ctx, const mvb200_blockfps_params *p, const mvb200_super_params *sp, int time256, const uint8_t *const left_super[3], const ptrdiff_t left_stride[3],
                                     const uint8_t *const right_super[3], const ptrdiff_t right_stride[3], const uint8_t *vec_b, const uint8_t *vec_f, uint8_t *const dst[3], const ptrdiff_t dst_stride[3], int32_t *usable);

/* The whole mv.Super -> 2 x mv.Analyse -> mv.FlowFPS chain for a window of OUTPUT frames in one call, on HOST buffers in
 * the clip layout (the plugin's fused path for the canonical frame-rate conversion script).  host_src holds the n
 * source frames [src_first, src_first + n) of the input clip (the window's left / right frames and a halo of 2 x delta
 * frames after and delta before them, clipped to the clip: the caller computes it with mvb200_flowfps_position);
 * frames [out_first, out_first + out_count) of the OUTPUT clip are written to host_dst.  ap_bw / ap_fw: the two
 * mv.Analyse instances (their num_frames is replaced by n internally).  Unusable frames are blended / copied from the
 * source frames as the filter does; usable_out (may be NULL): 1 = the frame was interpolated from the vectors (the
 * reference's frame then carries the LEFT super frame's properties), 0 = an input frame, a blend or a copy.  Synchronous. */
MVB200_API int mvb200_flowfps_clip_host(mvb200_ctx *ctx, const mvb200_super_params *sp, const mvb200_analyse_params *ap_bw, const mvb200_analyse_params *ap_fw, const mvb200_flowfps_params *p, const void *host_src, int src_first,
                                        int n, int out_first, int out_count, void *host_dst, int32_t *usable_out);

/* mv.Super -> mv.Analyse -> mv.Compensate for output frames [out_first, out_first + out_count) in one call, on HOST
 * buffers in the clip layout (the plugin's fused path).  host_src holds the n source frames [src_first, src_first + n):
 * the output frames and delta frames on the side the vectors point to (more does not hurt).  ap / cp: the filters'
 * parameters (their num_frames is replaced by n internally; relative delta only, no field shifts).  Synchronous. */
MVB200_API int mvb200_compensate_clip_host(mvb200_ctx *ctx, const mvb200_super_params *sp, const mvb200_analyse_params *ap, const mvb200_compensate_params *cp, const void *host_src, int src_first, int n, int out_first,
                                           int out_count, void *host_dst);

/* A window of a VECTOR clip from source frames: mv.Super of the n source frames [src_first, src_first + n) in host_src and
 * mv.Analyse of frames [out_first, out_first + out_count) (halo: delta frames on the side the vectors point to).  Writes, per
 * output frame, the SUPER frame (the pixels an mv.Analyse frame carries, MVAnalyse.c:224: super layout, mvb200_super_layout)
 * into host_super (out_count frames, frame_bytes apart) and the MVTools_vectors blob into host_vectors (vectors_stride apart, >=
 * ap->vectors_size).  Replaces mvsuperGetFrame + mvanalyseGetFrame (MVSuper.c:43-126, MVAnalyse.c:76-254) for a window of a
 * script whose vector clip is read by something else than this library's fused consumers.  Relative delta, no field shifts.
 * Synchronous. */
MVB200_API int mvb200_analyse_clip_host(mvb200_ctx *ctx, const mvb200_super_params *sp, const mvb200_analyse_params *ap, const void *host_src, int src_first, int n, int out_first, int out_count,
                                        void *host_super, void *host_vectors, int64_t vectors_stride);

/* ... with nrp mv.Recalculate passes (rp[0 .. nrp - 1], applied in order; MVRecalculate.c:68-254) between the search and the
 * compensation; cp was derived from the last pass's analysis data. */
MVB200_API int mvb200_compensate_recalc_clip_host(mvb200_ctx *ctx, const mvb200_super_params *sp, const mvb200_analyse_params *ap, const mvb200_recalculate_params *const *rp, int nrp,
                                                  const mvb200_compensate_params *cp, const void *host_src, int src_first, int n, int out_first, int out_count, void *host_dst);

/* The two above without the final wait: the call returns when the work is enqueued on the context (the source frames move
 * host -> device in pieces on a copy stream beside the previous call's kernels, the output pieces leave on another one);
 * the host buffers (pinned memory for truly asynchronous copies) are valid after mvb200_ctx_synchronize() or after an
 * event recorded behind the call (mvb200_ctx_record / mvb200_event_wait).  This is what the plugin's windows use. */
MVB200_API int mvb200_compensate_recalc_clip_host_async(mvb200_ctx *ctx, const mvb200_super_params *sp, const mvb200_analyse_params *ap, const mvb200_recalculate_params *const *rp, int nrp,
                                                        const mvb200_compensate_params *cp, const void *host_src, int src_first, int n, int out_first, int out_count, void *host_dst);
MVB200_API int mvb200_analyse_clip_host_async(mvb200_ctx *ctx, const mvb200_super_params *sp, const mvb200_analyse_params *ap, const void *host_src, int src_first, int n, int out_first, int out_count,
                                              void *host_super, void *host_vectors, int64_t vectors_stride);

/* The same for the per-pixel consumers: mv.Super -> mv.Analyse (one or two instances) -> mv.Flow / mv.FlowInter /
 * mv.FlowBlur for output frames [out_first, out_first + out_count), from the n source frames [src_first, src_first + n)
 * in host_src (halo: delta frames on both sides for Flow and FlowBlur, delta before and 2 x delta after for FlowInter).
 * Relative delta only, no field shifts.  Synchronous.  Replaces, for a window of frames, mvflowGetFrame (MVFlow.cpp:157-368),
 * mvflowinterGetFrame (MVFlowInter.c:83-448) and mvflowblurGetFrame (MVFlowBlur.c:144-337) together with the
 * mvsuperGetFrame / mvanalyseGetFrame calls that feed them.  usable_out (Flow and FlowInter, may be NULL): one flag per
 * output frame, 1 = made from the vectors (the reference's frame then carries the properties of a SUPER frame: the
 * reference frame's for Flow, frame n's for FlowInter), 0 = the copy / blend of source frames the reference returns at
 * scene changes and clip ends (MVFlow.cpp:362-366, MVFlowInter.c:412-447). */
MVB200_API int mvb200_flow_clip_host(mvb200_ctx *ctx, const mvb200_super_params *sp, const mvb200_analyse_params *ap, const mvb200_flow_params *fp, const void *host_src, int src_first, int n, int out_first, int out_count,
                                     void *host_dst, int32_t *usable_out);
MVB200_API int mvb200_flowinter_clip_host(mvb200_ctx *ctx, const mvb200_super_params *sp, const mvb200_analyse_params *ap_bw, const mvb200_analyse_params *ap_fw, const mvb200_flowinter_params *p, const void *host_src, int src_first,
                                          int n, int out_first, int out_count, void *host_dst, int32_t *usable_out);
MVB200_API int mvb200_flowblur_clip_host(mvb200_ctx *ctx, const mvb200_super_params *sp, const mvb200_analyse_params *ap_bw, const mvb200_analyse_params *ap_fw, const mvb200_flowblur_params *p, const void *host_src, int src_first,
                                         int n, int out_first, int out_count, void *host_dst);

/* The same for mv.BlockFPS (mvblockfpsGetFrame, MVBlockFPS.c:240-680, with the mvsuperGetFrame / mvanalyseGetFrame calls
 * that feed it): output frames [out_first, out_first + out_count) of the rate-converted clip from the n source frames
 * [src_first, src_first + n) -- every window frame's left and right source frames (mvb200_blockfps_position; no further
 * halo: the two searches only look at each other's frame).  usable_out (may be NULL): 1 = interpolated from the vectors
 * (the reference's frame then carries the LEFT super frame's properties), 0 = an input frame itself, or the blend /
 * copy of the source frames the reference returns for poor estimations (:636-670). */
MVB200_API int mvb200_blockfps_clip_host(mvb200_ctx *ctx, const mvb200_super_params *sp, const mvb200_analyse_params *ap_bw, const mvb200_analyse_params *ap_fw, const mvb200_blockfps_params *p, const void *host_src, int src_first,
                                         int n, int out_first, int out_count, void *host_dst, int32_t *usable_out);

#ifdef __cplusplus
}
#endif

#endif /* MVTOOLS_B200_H */
