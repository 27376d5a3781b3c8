/* TEST INFRASTRUCTURE -- CPU oracle for the mv.* hot path.  See mvtools_oracle.c.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load
 * this; the product (libmvtools_b200.so, the plugin) never does. */
#ifndef MVTOOLS_ORACLE_H
#define MVTOOLS_ORACLE_H

#include "../include/mvtools_b200.h" /* POD parameter structs only */

#ifdef __cplusplus
extern "C" {
#endif

#define MVO_API __attribute__((visibility("default")))

MVO_API const char *mvo_last_error(void);

/* mvsuperGetFrame, MVSuper.c:43-126 */
MVO_API int mvo_super_frame(const mvb200_super_params *p, const uint8_t *const src[3], const ptrdiff_t src_stride[3], uint8_t *const dst[3], const ptrdiff_t dst_stride[3]);

/* the same with a pelclip: mvpRefineExt, MVFrame.cpp:1530-1631 (pel planes of level 0 taken from an upsized clip) */
MVO_API int mvo_super_frame_pel(const mvb200_super_params *p, const uint8_t *const src[3], const ptrdiff_t src_stride[3], const uint8_t *const pel[3], const ptrdiff_t pel_stride[3], int pel_padded, uint8_t *const dst[3], const ptrdiff_t dst_stride[3]);

/* mvanalyseGetFrame, MVAnalyse.c:76-254 (ref_super == NULL -> default blob) */
MVO_API int mvo_analyse_frame(const mvb200_analyse_params *ap, const mvb200_super_params *sp, const uint8_t *const src_super[3], const ptrdiff_t src_stride[3], const uint8_t *const ref_super[3], const ptrdiff_t ref_stride[3], int field_shift, uint8_t *vectors_out);

/* mvdegrainGetFrame<radius>, MVDegrains.cpp:85-330 */
MVO_API int mvo_degrain_frame(const mvb200_degrain_params *dp, const mvb200_super_params *sp, const uint8_t *const src[3], const ptrdiff_t src_stride[3], const uint8_t *const *ref_super, const ptrdiff_t *ref_stride, const uint8_t *const *vectors, uint8_t *const dst[3], const ptrdiff_t dst_stride[3]);

/* mvrecalculateGetFrame, MVRecalculate.c:68-254 */
MVO_API int mvo_recalculate_frame(const mvb200_recalculate_params *rp, const mvb200_super_params *sp, const uint8_t *const src_super[3], const ptrdiff_t src_stride[3], const uint8_t *const ref_super[3], const ptrdiff_t ref_stride[3], const uint8_t *old_vectors, int field_shift, uint8_t *vectors_out);

/* mvcompensateGetFrame, MVCompensate.c:73-374 */
MVO_API int mvo_compensate_frame(const mvb200_compensate_params *cp, const mvb200_super_params *sp, const uint8_t *const src_super[3], const ptrdiff_t src_stride[3], const uint8_t *const ref_super[3], const ptrdiff_t ref_stride[3], const uint8_t *vectors, int field_shift, uint8_t *const dst[3], const ptrdiff_t dst_stride[3]);

/* overInit, Overlap.cpp:40-125: 9 windows of nx*ny int16 */
MVO_API void mvo_overlap_windows(int nx, int ny, int ox, int oy, int16_t *out9);

/* ---- consumers of the vector clips (mvtools_oracle_flow.c) ---- */
/* mvscdetectionGetFrame, MVSCDetection.c:43-74 */
MVO_API int mvo_scdetection_frame(const mvb200_scdetection_params *p, const uint8_t *vectors, int32_t *scene_change);
/* mvmaskGetFrame, MVMask.c:74-203 */
MVO_API int mvo_mask_frame(const mvb200_mask_params *p, const uint8_t *src_luma, ptrdiff_t src_stride, const uint8_t *vectors, uint8_t *const dst[3], const ptrdiff_t dst_stride[3]);
/* mvfinestGetFrame, MVFinest.c:47-125 */
MVO_API int mvo_finest_frame(const mvb200_super_params *sp, const uint8_t *const super[3], const ptrdiff_t super_stride[3], uint8_t *const dst[3], const ptrdiff_t dst_stride[3]);
/* mvflowGetFrame, MVFlow.cpp:157-368 */
MVO_API int mvo_flow_frame(const mvb200_flow_params *fp, const mvb200_super_params *sp, const uint8_t *const ref_super[3], const ptrdiff_t ref_stride[3], const uint8_t *vectors, int field_shift, uint8_t *const dst[3], const ptrdiff_t dst_stride[3], int32_t *usable);
/* mvflowinterGetFrame, MVFlowInter.c:83-448 (also the core of mvflowfpsGetFrame) and Blend, MaskFun.cpp:349-373 */
MVO_API int mvo_flowinter_frame(const mvb200_flowinter_params *p, const mvb200_super_params *sp, int time256, const uint8_t *const left_super[3], const ptrdiff_t left_stride[3], const uint8_t *const right_super[3],
                                const ptrdiff_t right_stride[3], const uint8_t *vec_b, const uint8_t *vec_f, const uint8_t *vec_bb, const uint8_t *vec_ff, uint8_t *const dst[3], const ptrdiff_t dst_stride[3], int32_t *usable);
MVO_API int mvo_blend_frame(const mvb200_super_params *sp, int time256, const uint8_t *const src[3], const ptrdiff_t src_stride[3], const uint8_t *const ref[3], const ptrdiff_t ref_stride[3], uint8_t *const dst[3],
                            const ptrdiff_t dst_stride[3]);
/* mvflowblurGetFrame, MVFlowBlur.c:144-337 */
MVO_API int mvo_flowblur_frame(const mvb200_flowblur_params *p, const mvb200_super_params *sp, const uint8_t *const super[3], const ptrdiff_t super_stride[3], const uint8_t *vec_b, const uint8_t *vec_f, uint8_t *const dst[3],
                               const ptrdiff_t dst_stride[3], int32_t *usable);
/* mvblockfpsGetFrame, MVBlockFPS.c:240-640 */
MVO_API int mvo_blockfps_frame(const mvb200_blockfps_params *p, const mvb200_super_params *sp, int time256, const uint8_t *const left_super[3], const ptrdiff_t left_stride[3], const uint8_t *const right_super[3],
                               const ptrdiff_t right_stride[3], const uint8_t *vec_b, const uint8_t *vec_f, uint8_t *const dst[3], const ptrdiff_t dst_stride[3], int32_t *usable);
/* InitTables, SimpleResize.cpp:26-56 */
MVO_API void mvo_simple_resize_tables(int out, int in, int32_t *offsets, int32_t *weights);

/* kernels exposed for unit parity tests */
MVO_API unsigned mvo_sad(const uint8_t *a, ptrdiff_t apitch, const uint8_t *b, ptrdiff_t bpitch, int w, int h, int bytes);
MVO_API unsigned mvo_satd(const uint8_t *a, ptrdiff_t apitch, const uint8_t *b, ptrdiff_t bpitch, int w, int h, int bytes);

#ifdef __cplusplus
}
#endif
#endif
