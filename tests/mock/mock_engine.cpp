// TEST INFRASTRUCTURE -- a stand-in for the ENGINE half of the C ABI, for testing the VapourSynth plugin's host-side
// glue (request patterns, reference-frame selection, field shifts, frame properties, fall-back paths at scene
// changes) on a machine without a GPU.
//
// tests/mock/libmvtools_mockengine.so = the product's plugin.cpp, unchanged, linked against THIS file + the product's
// host_logic.cpp instead of libmvtools_b200.so.  The per-frame entry points below forward to the CPU oracle
// (oracle/_ref/liboracle.so), so a script driven through that plugin by the mini host must give the reference's
// goldens if -- and only if -- the glue hands the right frames, blobs and parameters to the engine.
// Nothing here is shipped, and the product plugin (vapoursynth-mvtools_b200/libmvtools.so) cannot reach it: it links
// libmvtools_b200.so by $ORIGIN rpath and has no CPU path.  Only `-m "not gpu"` tests load this library.
#include "../../include/mvtools_b200.h"
#include "../../oracle/mvtools_oracle.h"

#include <cstring>

struct mvb200_ctx {
    int dummy;
};

namespace mvb200 {
int set_error(const char *fmt, ...);
}

static int fwd(int rc) {
    if (rc)
        mvb200::set_error("%s", mvo_last_error());
    return rc;
}

extern "C" {

MVB200_API int mvb200_ctx_create(int, mvb200_ctx **out) {
    *out = new mvb200_ctx{0};
    return 0;
}
MVB200_API void mvb200_ctx_destroy(mvb200_ctx *ctx) { delete ctx; }
MVB200_API int mvb200_ctx_synchronize(mvb200_ctx *) { return 0; }

MVB200_API int mvb200_super_frame(mvb200_ctx *, const mvb200_super_params *p, const uint8_t *const src[3], const ptrdiff_t src_stride[3], uint8_t *const dst[3], const ptrdiff_t dst_stride[3]) {
    return fwd(mvo_super_frame(p, src, src_stride, dst, dst_stride));
}
MVB200_API int mvb200_super_frame_pel(mvb200_ctx *, const mvb200_super_params *p, const uint8_t *const src[3], const ptrdiff_t src_stride[3], const uint8_t *const pel[3], const ptrdiff_t pel_stride[3], int pel_padded, uint8_t *const dst[3],
                                      const ptrdiff_t dst_stride[3]) {
    return fwd(mvo_super_frame_pel(p, src, src_stride, pel, pel_stride, pel_padded, dst, dst_stride));
}
MVB200_API int mvb200_analyse_frame(mvb200_ctx *, const mvb200_analyse_params *ap, const mvb200_super_params *sp, const uint8_t *const src_super[3], const ptrdiff_t src_stride[3], const uint8_t *const ref_super[3],
                                    const ptrdiff_t ref_stride[3], int field_shift, uint8_t *vectors_out) {
    return fwd(mvo_analyse_frame(ap, sp, src_super, src_stride, ref_super, ref_stride, field_shift, vectors_out));
}
MVB200_API int mvb200_analyse_frames(mvb200_ctx *, const mvb200_analyse_params *ap, const mvb200_super_params *sp, int nframes, const uint8_t *const *frames, const ptrdiff_t *strides, int njobs, const int32_t *src_index,
                                     const int32_t *ref_index, const int32_t *field_shift, uint8_t *const *vectors_out) {
    for (int k = 0; k < njobs; k++) {
        if (src_index[k] < 0 || src_index[k] >= nframes || ref_index[k] >= nframes)
            return mvb200::set_error("mock: bad frame index");
        const int r = ref_index[k];
        if (fwd(mvo_analyse_frame(ap, sp, frames + 3 * src_index[k], strides + 3 * src_index[k], r >= 0 ? frames + 3 * r : nullptr, r >= 0 ? strides + 3 * r : nullptr, r >= 0 ? field_shift[k] : 0, vectors_out[k])))
            return 1;
    }
    return 0;
}
MVB200_API int mvb200_degrain_frame(mvb200_ctx *, const mvb200_degrain_params *dp, const mvb200_super_params *sp, const uint8_t *const src[3], const ptrdiff_t src_stride[3], const uint8_t *const *const ref_super, const ptrdiff_t *ref_stride,
                                    const uint8_t *const *vectors, uint8_t *const dst[3], const ptrdiff_t dst_stride[3]) {
    return fwd(mvo_degrain_frame(dp, sp, src, src_stride, ref_super, ref_stride, vectors, dst, dst_stride));
}
MVB200_API int mvb200_compensate_frame(mvb200_ctx *, const mvb200_compensate_params *cp, const mvb200_super_params *sp, const uint8_t *const src_super[3], const ptrdiff_t src_stride[3], const uint8_t *const ref_super[3],
                                       const ptrdiff_t ref_stride[3], const uint8_t *vectors, int field_shift, uint8_t *const dst[3], const ptrdiff_t dst_stride[3]) {
    return fwd(mvo_compensate_frame(cp, sp, src_super, src_stride, ref_super, ref_stride, vectors, field_shift, dst, dst_stride));
}
MVB200_API int mvb200_recalculate_frame(mvb200_ctx *, const mvb200_recalculate_params *rp, const mvb200_super_params *sp, const uint8_t *const src_super[3], const ptrdiff_t src_stride[3], const uint8_t *const ref_super[3],
                                        const ptrdiff_t ref_stride[3], const uint8_t *old_vectors, int field_shift, uint8_t *vectors_out) {
    return fwd(mvo_recalculate_frame(rp, sp, src_super, src_stride, ref_super, ref_stride, old_vectors, field_shift, vectors_out));
}
MVB200_API int mvb200_scdetection_frame(mvb200_ctx *, const mvb200_scdetection_params *p, const uint8_t *vectors, int32_t *scene_change) { return fwd(mvo_scdetection_frame(p, vectors, scene_change)); }
MVB200_API int mvb200_mask_frame(mvb200_ctx *, const mvb200_mask_params *p, const uint8_t *src_luma, ptrdiff_t src_stride, const uint8_t *vectors, uint8_t *const dst[3], const ptrdiff_t dst_stride[3]) {
    return fwd(mvo_mask_frame(p, src_luma, src_stride, vectors, dst, dst_stride));
}
MVB200_API int mvb200_finest_frame(mvb200_ctx *, const mvb200_super_params *sp, const uint8_t *const super[3], const ptrdiff_t super_stride[3], uint8_t *const dst[3], const ptrdiff_t dst_stride[3]) {
    return fwd(mvo_finest_frame(sp, super, super_stride, dst, dst_stride));
}
MVB200_API int mvb200_flow_frame(mvb200_ctx *, const mvb200_flow_params *fp, const mvb200_super_params *sp, const uint8_t *const ref_super[3], const ptrdiff_t ref_stride[3], const uint8_t *vectors, int field_shift, uint8_t *const dst[3],
                                 const ptrdiff_t dst_stride[3], int32_t *usable) {
    return fwd(mvo_flow_frame(fp, sp, ref_super, ref_stride, vectors, field_shift, dst, dst_stride, usable));
}
MVB200_API int mvb200_flowinter_frame(mvb200_ctx *, const mvb200_flowinter_params *p, const mvb200_super_params *sp, int time256, const uint8_t *const left_super[3], const ptrdiff_t left_stride[3], const uint8_t *const right_super[3],
                                      const ptrdiff_t right_stride[3], const uint8_t *vec_b, const uint8_t *vec_f, const uint8_t *vec_bb, const uint8_t *vec_ff, uint8_t *const dst[3], const ptrdiff_t dst_stride[3], int32_t *usable) {
    return fwd(mvo_flowinter_frame(p, sp, time256, left_super, left_stride, right_super, right_stride, vec_b, vec_f, vec_bb, vec_ff, dst, dst_stride, usable));
}
MVB200_API int mvb200_blend_frame(mvb200_ctx *, const mvb200_super_params *sp, int time256, const uint8_t *const src[3], const ptrdiff_t src_stride[3], const uint8_t *const ref[3], const ptrdiff_t ref_stride[3], uint8_t *const dst[3],
                                  const ptrdiff_t dst_stride[3]) {
    return fwd(mvo_blend_frame(sp, time256, src, src_stride, ref, ref_stride, dst, dst_stride));
}
MVB200_API int mvb200_flowblur_frame(mvb200_ctx *, const mvb200_flowblur_params *p, const mvb200_super_params *sp, const uint8_t *const super[3], const ptrdiff_t super_stride[3], const uint8_t *vec_b, const uint8_t *vec_f,
                                     uint8_t *const dst[3], const ptrdiff_t dst_stride[3], int32_t *usable) {
    return fwd(mvo_flowblur_frame(p, sp, super, super_stride, vec_b, vec_f, dst, dst_stride, usable));
}
MVB200_API int mvb200_blockfps_frame(mvb200_ctx *, const mvb200_blockfps_params *p, const mvb200_super_params *sp, int time256, const uint8_t *const left_super[3], const ptrdiff_t left_stride[3], const uint8_t *const right_super[3],
                                     const ptrdiff_t right_stride[3], const uint8_t *vec_b, const uint8_t *vec_f, uint8_t *const dst[3], const ptrdiff_t dst_stride[3], int32_t *usable) {
    return fwd(mvo_blockfps_frame(p, sp, time256, left_super, left_stride, right_super, right_stride, vec_b, vec_f, dst, dst_stride, usable));
}

} // extern "C"
