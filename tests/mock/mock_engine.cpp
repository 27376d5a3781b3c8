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

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <unistd.h>

static std::atomic<long> g_ctx_created{0}; // mvb200_ctx_create calls so far (one context per device the plugin uses)
static std::atomic<long> g_clip_calls{0}; // mvb200_degrain_clip_host_async calls so far (the glue tests count the windows)

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
    g_ctx_created++;
    return 0;
}
MVB200_API void mvb200_ctx_destroy(mvb200_ctx *ctx) { delete ctx; }
// MVB200_MOCK_DEVICES: how many "devices" the stand-in pretends to have (the plugin then spreads its windows over them)
MVB200_API int mvb200_device_count(void) {
    const char *e = getenv("MVB200_MOCK_DEVICES");
    return e ? atoi(e) : 1;
}
MVB200_API int64_t mvb200_device_memory(int) { return (int64_t)180 << 30; }
// (the stand-in's window plan: the 16-frame minimum unless the look-ahead is given; every context the plugin asks for)
MVB200_API int mvb200_window_plan_query(const mvb200_analyse_params *, const mvb200_super_params *, int, int lookahead, int64_t, int64_t, int, int contexts_per_device, int64_t pinned_budget, int64_t,
                                        mvb200_window_plan *out) {
    out->frames = lookahead > 0 ? lookahead : 16;
    out->contexts_per_device = contexts_per_device < 1 ? 2 : contexts_per_device; // (unset: two, the engine's usual answer)
    out->pinned_budget = pinned_budget;
    return 0;
}
MVB200_API long mvb200_mock_contexts(void) { return g_ctx_created.load(); }
MVB200_API int mvb200_ctx_synchronize(mvb200_ctx *) { return 0; }
// the stand-in computes synchronously: an event is complete when it is recorded
struct mvb200_event {
    int dummy;
};
MVB200_API int mvb200_ctx_record(mvb200_ctx *, mvb200_event **out) {
    *out = new mvb200_event{0};
    return 0;
}
MVB200_API long mvb200_mock_clip_calls(void) { return g_clip_calls.load(); }
MVB200_API int mvb200_event_wait(mvb200_event *) { return 0; }
MVB200_API int mvb200_event_query(mvb200_event *) { return 0; }
MVB200_API void mvb200_event_destroy(mvb200_event *ev) { delete ev; }
// the automatic window of the fused paths: small here, so that the default configuration exercises several windows per test clip
MVB200_API int mvb200_search_plan_query(const mvb200_analyse_params *, const mvb200_super_params *, int, mvb200_search_plan *out) {
    out->kernel = 0;
    out->chains_per_sm = 1;
    out->chains_per_wave = 6;
    out->smem_per_chain = 0;
    return 0;
}
static std::atomic<int> g_pinned_allocs{0};
MVB200_API int mvb200_mock_pinned_allocs() { return g_pinned_allocs.load(); }
MVB200_API void *mvb200_pinned_alloc(size_t bytes) {
    g_pinned_allocs.fetch_add(1);
    return malloc(bytes);
}
MVB200_API void mvb200_pinned_free(void *p) { free(p); }

// the layouts of the device-resident clips, restated for the stand-in (the product's live in engine.cu)
static long long round_up256(long long v) { return (v + 255) / 256 * 256; }
MVB200_API void mvb200_clip_layout(const mvb200_super_params *p, mvb200_frame_layout *l) {
    memset(l, 0, sizeof(*l));
    l->nplanes = p->nplanes;
    l->bytes_per_sample = (p->bits + 7) / 8;
    long long off = 0;
    for (int i = 0; i < p->nplanes; i++) {
        l->plane_w[i] = i ? p->width >> p->ssw : p->width;
        l->plane_h[i] = i ? p->height >> p->ssh : p->height;
        l->pitch[i] = round_up256((long long)l->plane_w[i] * l->bytes_per_sample);
        l->offset[i] = off;
        off += l->pitch[i] * l->plane_h[i];
    }
    l->frame_bytes = round_up256(off);
}

// (only the plugin's estimate of a window's device footprint reads it: the super frame's planes at the same pitch rule)
MVB200_API void mvb200_super_layout(const mvb200_super_params *p, mvb200_frame_layout *l) {
    mvb200_super_geom g;
    mvb200_super_geometry(p, &g);
    memset(l, 0, sizeof(*l));
    l->nplanes = p->nplanes;
    l->bytes_per_sample = (p->bits + 7) / 8;
    long long off = 0;
    for (int i = 0; i < p->nplanes; i++) {
        l->plane_w[i] = g.plane_w[i];
        l->plane_h[i] = g.plane_h[i];
        l->pitch[i] = round_up256((long long)g.plane_w[i] * l->bytes_per_sample + 4);
        l->offset[i] = off;
        off += l->pitch[i] * g.plane_h[i];
    }
    l->frame_bytes = round_up256(off);
}

// The clip pipeline on host memory, frame by frame through the oracle: Super of every frame of the segment, the nap
// searches and DegrainN for frames [first, first + count).  Reference frames follow the analyse parameters with
// num_frames = the segment length, as in the engine.
MVB200_API int mvb200_degrain_recalc_clip_host_async(mvb200_ctx *, const mvb200_super_params *sp, const mvb200_analyse_params *const *ap, int nap, const mvb200_recalculate_params *const *rp, int nrp,
                                                     const mvb200_degrain_params *dp, const void *host_src, void *host_dst, int n, int first, int count) {
    const long call = ++g_clip_calls;
    if (const char *fail = getenv("MVB200_MOCK_FAIL_CLIP_CALL")) // (the glue's error path: the k-th window's engine call fails)
        if (atol(fail) == call)
            return mvb200::set_error("mock: injected failure of clip call %ld", call);
    mvb200_frame_layout cl;
    mvb200_clip_layout(sp, &cl);
    mvb200_super_geom g;
    mvb200_super_geometry(sp, &g);
    const int bps = cl.bytes_per_sample;
    std::vector<std::vector<uint8_t>> sup((size_t)n * 3);
    ptrdiff_t sstride[3] = {0, 0, 0};
    auto src_planes = [&](int f, const uint8_t *p[3], ptrdiff_t st[3]) {
        for (int i = 0; i < 3; i++) {
            p[i] = i < cl.nplanes ? (const uint8_t *)host_src + (size_t)cl.frame_bytes * f + cl.offset[i] : nullptr;
            st[i] = i < cl.nplanes ? (ptrdiff_t)cl.pitch[i] : 0;
        }
    };
    for (int f = 0; f < n; f++) {
        const uint8_t *p[3];
        ptrdiff_t st[3];
        uint8_t *d[3] = {nullptr, nullptr, nullptr};
        src_planes(f, p, st);
        for (int i = 0; i < sp->nplanes; i++) {
            sstride[i] = (ptrdiff_t)g.plane_w[i] * bps;
            sup[(size_t)f * 3 + i].assign((size_t)sstride[i] * g.plane_h[i], 0);
            d[i] = sup[(size_t)f * 3 + i].data();
        }
        if (fwd(mvo_super_frame(sp, p, st, d, sstride)))
            return 1;
    }
    auto sup_planes = [&](int f, const uint8_t *p[3]) {
        for (int i = 0; i < 3; i++)
            p[i] = i < sp->nplanes ? sup[(size_t)f * 3 + i].data() : nullptr;
    };
    for (int k = 0; k < count; k++) {
        const int f = first + k;
        std::vector<std::vector<uint8_t>> blobs((size_t)nap);
        std::vector<const uint8_t *> vecs((size_t)nap), refs((size_t)nap * 3, nullptr);
        std::vector<ptrdiff_t> rstr((size_t)nap * 3, 0);
        for (int i = 0; i < nap; i++) {
            const int d = ap[i]->ad.nDeltaFrame;
            int nref = d > 0 ? f + (ap[i]->ad.isBackward ? d : -d) : -d;
            if (nref < 0 || nref >= n)
                nref = -1;
            const uint8_t *s[3], *r[3];
            sup_planes(f, s);
            if (nref >= 0)
                sup_planes(nref, r);
            blobs[i].assign((size_t)ap[i]->vectors_size, 0);
            if (fwd(mvo_analyse_frame(ap[i], sp, s, sstride, nref >= 0 ? r : nullptr, nref >= 0 ? sstride : nullptr, 0, blobs[i].data())))
                return 1;
            for (int k2 = 0; rp && k2 < nrp; k2++) { // the refinement passes of this vector clip, in order
                if (!rp[i * nrp + k2])
                    continue;
                std::vector<uint8_t> nb((size_t)rp[i * nrp + k2]->vectors_size, 0);
                if (fwd(mvo_recalculate_frame(rp[i * nrp + k2], sp, s, sstride, nref >= 0 ? r : nullptr, nref >= 0 ? sstride : nullptr, blobs[i].data(), 0, nb.data())))
                    return 1;
                blobs[i].swap(nb);
            }
            vecs[i] = blobs[i].data();
            if (nref >= 0)
                for (int pl = 0; pl < sp->nplanes; pl++) {
                    refs[(size_t)i * 3 + pl] = sup[(size_t)nref * 3 + pl].data();
                    rstr[(size_t)i * 3 + pl] = sstride[pl];
                }
        }
        const uint8_t *p[3];
        ptrdiff_t st[3];
        uint8_t *d[3];
        src_planes(f, p, st);
        for (int i = 0; i < 3; i++)
            d[i] = i < cl.nplanes ? (uint8_t *)host_dst + (size_t)cl.frame_bytes * k + cl.offset[i] : nullptr;
        if (fwd(mvo_degrain_frame(dp, sp, p, st, refs.data(), rstr.data(), vecs.data(), d, st)))
            return 1;
    }
    return 0;
}

MVB200_API int mvb200_degrain_clip_host_async(mvb200_ctx *ctx, const mvb200_super_params *sp, const mvb200_analyse_params *const *ap, int nap, const mvb200_degrain_params *dp, const void *host_src, void *host_dst, int n, int first,
                                              int count) {
    return mvb200_degrain_recalc_clip_host_async(ctx, sp, ap, nap, nullptr, 0, dp, host_src, host_dst, n, first, count);
}

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

// mvb200_flowfps_clip_host on host memory through the oracle, with the engine's conventions: the segment is the clip for
// the two searches, output frames are placed by mvb200_flowfps_position
MVB200_API int mvb200_flowfps_clip_host(mvb200_ctx *, const mvb200_super_params *sp, const mvb200_analyse_params *ap_bw, const mvb200_analyse_params *ap_fw, const mvb200_flowfps_params *p, const void *host_src, int src_first, int n,
                                        int out_first, int out_count, void *host_dst, int32_t *usable_out) {
    for (int k = 0; usable_out && k < out_count; k++)
        usable_out[k] = 0;
    mvb200_frame_layout cl;
    mvb200_clip_layout(sp, &cl);
    mvb200_super_geom g;
    mvb200_super_geometry(sp, &g);
    const int bps = cl.bytes_per_sample, off = p->inter.vd.nDeltaFrame, N = p->num_frames_in;
    std::vector<std::vector<uint8_t>> sup((size_t)n * 3);
    ptrdiff_t sstride[3] = {0, 0, 0};
    auto src_planes = [&](int local, const uint8_t *q[3], ptrdiff_t st[3]) {
        for (int i = 0; i < 3; i++) {
            q[i] = i < cl.nplanes ? (const uint8_t *)host_src + (size_t)cl.frame_bytes * local + cl.offset[i] : nullptr;
            st[i] = i < cl.nplanes ? (ptrdiff_t)cl.pitch[i] : 0;
        }
    };
    for (int f = 0; f < n; f++) {
        const uint8_t *q[3];
        ptrdiff_t st[3];
        uint8_t *d[3] = {nullptr, nullptr, nullptr};
        src_planes(f, q, st);
        for (int i = 0; i < sp->nplanes; i++) {
            sstride[i] = (ptrdiff_t)g.plane_w[i] * bps;
            sup[(size_t)f * 3 + i].assign((size_t)sstride[i] * g.plane_h[i], 0);
            d[i] = sup[(size_t)f * 3 + i].data();
        }
        if (fwd(mvo_super_frame(sp, q, st, d, sstride)))
            return 1;
    }
    auto sup_planes = [&](int local, const uint8_t *q[3]) {
        for (int i = 0; i < 3; i++)
            q[i] = i < sp->nplanes ? sup[(size_t)local * 3 + i].data() : nullptr;
    };
    const mvb200_analyse_params *aps[2] = {ap_bw, ap_fw};
    std::vector<std::vector<uint8_t>> vec[2];
    for (int k = 0; k < 2; k++) {
        vec[k].resize((size_t)n);
        for (int f = 0; f < n; f++) {
            const int d = aps[k]->ad.nDeltaFrame;
            int nref = f + (aps[k]->ad.isBackward ? d : -d);
            if (nref < 0 || nref >= n)
                nref = -1;
            const uint8_t *s_[3], *r_[3];
            sup_planes(f, s_);
            if (nref >= 0)
                sup_planes(nref, r_);
            vec[k][f].assign((size_t)aps[k]->vectors_size, 0);
            if (fwd(mvo_analyse_frame(aps[k], sp, s_, sstride, nref >= 0 ? r_ : nullptr, nref >= 0 ? sstride : nullptr, 0, vec[k][f].data())))
                return 1;
        }
    }
    for (int k = 0; k < out_count; k++) {
        int32_t nleft, time256;
        mvb200_flowfps_position(p, out_first + k, &nleft, &time256);
        const int nright = nleft + off, fl = nleft < N - 1 ? nleft : N - 1, fr = nright < N - 1 ? nright : N - 1;
        uint8_t *d[3];
        ptrdiff_t dstr[3];
        for (int i = 0; i < 3; i++) {
            d[i] = i < cl.nplanes ? (uint8_t *)host_dst + (size_t)cl.frame_bytes * k + cl.offset[i] : nullptr;
            dstr[i] = i < cl.nplanes ? (ptrdiff_t)cl.pitch[i] : 0;
        }
        auto in_seg = [&](int f) { return f >= src_first && f < src_first + n; };
        const uint8_t *ql[3], *qr[3];
        ptrdiff_t stl[3], str_[3];
        if (!in_seg(fl) || !in_seg(fr))
            return mvb200::set_error("mock: source frame outside the segment");
        src_planes(fl - src_first, ql, stl);
        src_planes(fr - src_first, qr, str_);
        auto copy = [&](const uint8_t *const q[3], const ptrdiff_t st[3]) {
            for (int i = 0; i < cl.nplanes; i++)
                for (int y = 0; y < cl.plane_h[i]; y++)
                    memcpy(d[i] + y * dstr[i], q[i] + y * st[i], (size_t)cl.plane_w[i] * bps);
        };
        if (time256 == 0) {
            copy(ql, stl);
            continue;
        }
        if (time256 == 256) {
            copy(qr, str_);
            continue;
        }
        int32_t usable = 0;
        if (nleft < N && nright < N) {
            if (!in_seg(nright))
                return mvb200::set_error("mock: source frame outside the segment");
            const int L = nleft - src_first, R = nright - src_first;
            const uint8_t *sl[3], *sr[3];
            sup_planes(L, sl);
            sup_planes(R, sr);
            const bool extra = p->maskmode == 2;
            if (fwd(mvo_flowinter_frame(&p->inter, sp, time256, sl, sstride, sr, sstride, vec[0][L].data(), vec[1][R].data(), extra ? vec[0][R].data() : nullptr, extra ? vec[1][L].data() : nullptr, d, dstr, &usable)))
                return 1;
        }
        if (usable_out)
            usable_out[k] = usable;
        if (!usable) {
            if (p->inter.blend) {
                if (fwd(mvo_blend_frame(sp, time256, ql, stl, qr, str_, d, dstr)))
                    return 1;
            } else {
                copy(ql, stl);
            }
        }
    }
    return 0;
}

// A window of a vector clip from source frames, frame by frame through the oracle: the super frames of the output range (super
// layout) and their blobs.
MVB200_API int mvb200_analyse_clip_host(mvb200_ctx *, const mvb200_super_params *sp, const mvb200_analyse_params *ap, const void *host_src, int src_first, int n, int out_first, int out_count, void *host_super,
                                        void *host_vectors, int64_t vectors_stride) {
    if (const char *nc = getenv("MVB200_MOCK_NOCOMPUTE")) { // (tools/bench_glue.py: the plugin's host plumbing alone, outputs left as they are; the value = a call's latency in ms)
        if (atoi(nc) > 1)
            usleep((useconds_t)atoi(nc) * 1000);
        return 0;
    }
    mvb200_frame_layout cl, sl;
    mvb200_clip_layout(sp, &cl);
    mvb200_super_layout(sp, &sl);
    mvb200_super_geom g;
    mvb200_super_geometry(sp, &g);
    const int bps = cl.bytes_per_sample;
    std::vector<std::vector<uint8_t>> sup((size_t)n * 3);
    ptrdiff_t sstride[3] = {0, 0, 0};
    for (int f = 0; f < n; f++) {
        const uint8_t *q[3];
        ptrdiff_t st[3];
        uint8_t *d[3] = {nullptr, nullptr, nullptr};
        for (int i = 0; i < 3; i++) {
            q[i] = i < cl.nplanes ? (const uint8_t *)host_src + (size_t)cl.frame_bytes * f + cl.offset[i] : nullptr;
            st[i] = i < cl.nplanes ? (ptrdiff_t)cl.pitch[i] : 0;
        }
        for (int i = 0; i < sp->nplanes; i++) {
            sstride[i] = (ptrdiff_t)g.plane_w[i] * bps;
            sup[(size_t)f * 3 + i].assign((size_t)sstride[i] * g.plane_h[i], 0);
            d[i] = sup[(size_t)f * 3 + i].data();
        }
        if (fwd(mvo_super_frame(sp, q, st, d, sstride)))
            return 1;
    }
    for (int k = 0; k < out_count; k++) {
        const int f = out_first - src_first + k, dlt = ap->ad.nDeltaFrame;
        int nref = f + (ap->ad.isBackward ? dlt : -dlt);
        if (nref < 0 || nref >= n)
            nref = -1;
        const uint8_t *s_[3], *r_[3];
        for (int i = 0; i < 3; i++) {
            s_[i] = i < sp->nplanes ? sup[(size_t)f * 3 + i].data() : nullptr;
            r_[i] = i < sp->nplanes && nref >= 0 ? sup[(size_t)nref * 3 + i].data() : nullptr;
        }
        if (fwd(mvo_analyse_frame(ap, sp, s_, sstride, nref >= 0 ? r_ : nullptr, nref >= 0 ? sstride : nullptr, 0, (uint8_t *)host_vectors + (size_t)vectors_stride * k)))
            return 1;
        for (int i = 0; i < sp->nplanes; i++)
            for (int y = 0; y < g.plane_h[i]; y++)
                memcpy((uint8_t *)host_super + (size_t)sl.frame_bytes * k + sl.offset[i] + (size_t)sl.pitch[i] * y, sup[(size_t)f * 3 + i].data() + (size_t)sstride[i] * y, (size_t)sstride[i]);
    }
    return 0;
}

MVB200_API int mvb200_compensate_recalc_clip_host(mvb200_ctx *, const mvb200_super_params *sp, const mvb200_analyse_params *ap, const mvb200_recalculate_params *const *rp, int nrp, const mvb200_compensate_params *cp,
                                                  const void *host_src, int src_first, int n, int out_first, int out_count, void *host_dst) {
    mvb200_frame_layout cl;
    mvb200_clip_layout(sp, &cl);
    mvb200_super_geom g;
    mvb200_super_geometry(sp, &g);
    const int bps = cl.bytes_per_sample;
    std::vector<std::vector<uint8_t>> sup((size_t)n * 3);
    ptrdiff_t sstride[3] = {0, 0, 0};
    for (int f = 0; f < n; f++) {
        const uint8_t *q[3];
        ptrdiff_t st[3];
        uint8_t *d[3] = {nullptr, nullptr, nullptr};
        for (int i = 0; i < 3; i++) {
            q[i] = i < cl.nplanes ? (const uint8_t *)host_src + (size_t)cl.frame_bytes * f + cl.offset[i] : nullptr;
            st[i] = i < cl.nplanes ? (ptrdiff_t)cl.pitch[i] : 0;
        }
        for (int i = 0; i < sp->nplanes; i++) {
            sstride[i] = (ptrdiff_t)g.plane_w[i] * bps;
            sup[(size_t)f * 3 + i].assign((size_t)sstride[i] * g.plane_h[i], 0);
            d[i] = sup[(size_t)f * 3 + i].data();
        }
        if (fwd(mvo_super_frame(sp, q, st, d, sstride)))
            return 1;
    }
    for (int k = 0; k < out_count; k++) {
        const int f = out_first - src_first + k, dlt = ap->ad.nDeltaFrame;
        int nref = f + (ap->ad.isBackward ? dlt : -dlt);
        if (nref < 0 || nref >= n)
            nref = -1;
        const uint8_t *s_[3], *r_[3];
        for (int i = 0; i < 3; i++) {
            s_[i] = i < sp->nplanes ? sup[(size_t)f * 3 + i].data() : nullptr;
            r_[i] = i < sp->nplanes && nref >= 0 ? sup[(size_t)nref * 3 + i].data() : nullptr;
        }
        std::vector<uint8_t> blob((size_t)ap->vectors_size, 0);
        if (fwd(mvo_analyse_frame(ap, sp, s_, sstride, nref >= 0 ? r_ : nullptr, nref >= 0 ? sstride : nullptr, 0, blob.data())))
            return 1;
        for (int k2 = 0; rp && k2 < nrp; k2++) {
            if (!rp[k2])
                continue;
            std::vector<uint8_t> nb((size_t)rp[k2]->vectors_size, 0);
            if (fwd(mvo_recalculate_frame(rp[k2], sp, s_, sstride, nref >= 0 ? r_ : nullptr, nref >= 0 ? sstride : nullptr, blob.data(), 0, nb.data())))
                return 1;
            blob.swap(nb);
        }
        uint8_t *d[3];
        ptrdiff_t dstr[3];
        for (int i = 0; i < 3; i++) {
            d[i] = i < cl.nplanes ? (uint8_t *)host_dst + (size_t)cl.frame_bytes * k + cl.offset[i] : nullptr;
            dstr[i] = i < cl.nplanes ? (ptrdiff_t)cl.pitch[i] : 0;
        }
        if (fwd(mvo_compensate_frame(cp, sp, s_, sstride, nref >= 0 ? r_ : nullptr, nref >= 0 ? sstride : nullptr, blob.data(), 0, d, dstr)))
            return 1;
    }
    return 0;
}

} // extern "C"

// The front of the per-pixel consumers' clip calls on host memory: Super of every frame of the segment and the blobs of
// nap mv.Analyse instances for every frame (the segment is the clip for the searches, as in the engine).
namespace {
struct MockSegment {
    mvb200_frame_layout cl;
    mvb200_super_geom g;
    const mvb200_super_params *sp;
    const uint8_t *host_src;
    int n, bps;
    std::vector<std::vector<uint8_t>> sup;
    std::vector<std::vector<uint8_t>> vec[2];
    ptrdiff_t sstride[3] = {0, 0, 0};
    void src_planes(int local, const uint8_t *q[3], ptrdiff_t st[3]) const {
        for (int i = 0; i < 3; i++) {
            q[i] = i < cl.nplanes ? host_src + (size_t)cl.frame_bytes * local + cl.offset[i] : nullptr;
            st[i] = i < cl.nplanes ? (ptrdiff_t)cl.pitch[i] : 0;
        }
    }
    void sup_planes(int local, const uint8_t *q[3]) const {
        for (int i = 0; i < 3; i++)
            q[i] = i < sp->nplanes ? sup[(size_t)local * 3 + i].data() : nullptr;
    }
    void dst_planes(void *host_dst, int k, uint8_t *d[3], ptrdiff_t dstr[3]) const {
        for (int i = 0; i < 3; i++) {
            d[i] = i < cl.nplanes ? (uint8_t *)host_dst + (size_t)cl.frame_bytes * k + cl.offset[i] : nullptr;
            dstr[i] = i < cl.nplanes ? (ptrdiff_t)cl.pitch[i] : 0;
        }
    }
    void copy_src(int local, uint8_t *const d[3], const ptrdiff_t dstr[3]) const {
        const uint8_t *q[3];
        ptrdiff_t st[3];
        src_planes(local, q, st);
        for (int i = 0; i < cl.nplanes; i++)
            for (int y = 0; y < cl.plane_h[i]; y++)
                memcpy(d[i] + y * dstr[i], q[i] + y * st[i], (size_t)cl.plane_w[i] * bps);
    }
    int build(const mvb200_super_params *sp_, const mvb200_analyse_params *const *ap, int nap, const void *src, int n_) {
        sp = sp_;
        host_src = (const uint8_t *)src;
        n = n_;
        mvb200_clip_layout(sp, &cl);
        mvb200_super_geometry(sp, &g);
        bps = cl.bytes_per_sample;
        sup.resize((size_t)n * 3);
        for (int f = 0; f < n; f++) {
            const uint8_t *q[3];
            ptrdiff_t st[3];
            uint8_t *d[3] = {nullptr, nullptr, nullptr};
            src_planes(f, q, st);
            for (int i = 0; i < sp->nplanes; i++) {
                sstride[i] = (ptrdiff_t)g.plane_w[i] * bps;
                sup[(size_t)f * 3 + i].assign((size_t)sstride[i] * g.plane_h[i], 0);
                d[i] = sup[(size_t)f * 3 + i].data();
            }
            if (fwd(mvo_super_frame(sp, q, st, d, sstride)))
                return 1;
        }
        for (int k = 0; k < nap; k++) {
            vec[k].resize((size_t)n);
            for (int f = 0; f < n; f++) {
                const int d = ap[k]->ad.nDeltaFrame;
                int nref = f + (ap[k]->ad.isBackward ? d : -d);
                if (nref < 0 || nref >= n)
                    nref = -1;
                const uint8_t *s_[3], *r_[3];
                sup_planes(f, s_);
                if (nref >= 0)
                    sup_planes(nref, r_);
                vec[k][f].assign((size_t)ap[k]->vectors_size, 0);
                if (fwd(mvo_analyse_frame(ap[k], sp, s_, sstride, nref >= 0 ? r_ : nullptr, nref >= 0 ? sstride : nullptr, 0, vec[k][f].data())))
                    return 1;
            }
        }
        return 0;
    }
};
} // namespace

extern "C" {

MVB200_API int mvb200_compensate_clip_host(mvb200_ctx *ctx, const mvb200_super_params *sp, const mvb200_analyse_params *ap, const mvb200_compensate_params *cp, const void *host_src, int src_first, int n, int out_first, int out_count,
                                           void *host_dst) {
    return mvb200_compensate_recalc_clip_host(ctx, sp, ap, nullptr, 0, cp, host_src, src_first, n, out_first, out_count, host_dst);
}

MVB200_API int mvb200_flow_clip_host(mvb200_ctx *, const mvb200_super_params *sp, const mvb200_analyse_params *ap, const mvb200_flow_params *fp, const void *host_src, int src_first, int n, int out_first, int out_count,
                                     void *host_dst, int32_t *usable_out) {
    MockSegment S;
    if (S.build(sp, &ap, 1, host_src, n))
        return 1;
    const int dlt = fp->vd.nDeltaFrame;
    for (int k = 0; k < out_count; k++) {
        const int f = out_first - src_first + k;
        int nref = fp->vd.isBackward ? f + dlt : f - dlt;
        if (nref < 0 || nref >= n)
            nref = -1;
        uint8_t *d[3];
        ptrdiff_t dstr[3];
        S.dst_planes(host_dst, k, d, dstr);
        const uint8_t *r_[3];
        int32_t usable = 0;
        if (nref >= 0) {
            S.sup_planes(nref, r_);
            if (fwd(mvo_flow_frame(fp, sp, r_, S.sstride, S.vec[0][f].data(), 0, d, dstr, &usable)))
                return 1;
        }
        if (!usable)
            S.copy_src(f, d, dstr);
        if (usable_out)
            usable_out[k] = usable;
    }
    return 0;
}

MVB200_API int mvb200_flowinter_clip_host(mvb200_ctx *, const mvb200_super_params *sp, const mvb200_analyse_params *ap_bw, const mvb200_analyse_params *ap_fw, const mvb200_flowinter_params *p, const void *host_src, int src_first,
                                          int n, int out_first, int out_count, void *host_dst, int32_t *usable_out) {
    MockSegment S;
    const mvb200_analyse_params *aps[2] = {ap_bw, ap_fw};
    if (S.build(sp, aps, 2, host_src, n))
        return 1;
    const int off = p->vd.nDeltaFrame;
    for (int k = 0; k < out_count; k++) {
        const int L = out_first - src_first + k, R = L + off, ref = R < n - 1 ? R : n - 1;
        uint8_t *d[3];
        ptrdiff_t dstr[3];
        S.dst_planes(host_dst, k, d, dstr);
        int32_t usable = 0;
        if (R < n) {
            const uint8_t *sl[3], *sr[3];
            S.sup_planes(L, sl);
            S.sup_planes(R, sr);
            if (fwd(mvo_flowinter_frame(p, sp, p->time256, sl, S.sstride, sr, S.sstride, S.vec[0][L].data(), S.vec[1][R].data(), S.vec[0][R].data(), S.vec[1][L].data(), d, dstr, &usable)))
                return 1;
        }
        if (usable_out)
            usable_out[k] = usable;
        if (!usable) {
            if (p->blend) {
                const uint8_t *ql[3], *qr[3];
                ptrdiff_t stl[3], str_[3];
                S.src_planes(L, ql, stl);
                S.src_planes(ref, qr, str_);
                if (fwd(mvo_blend_frame(sp, p->time256, ql, stl, qr, str_, d, dstr)))
                    return 1;
            } else {
                S.copy_src(L, d, dstr);
            }
        }
    }
    return 0;
}

MVB200_API int mvb200_flowblur_clip_host(mvb200_ctx *, const mvb200_super_params *sp, const mvb200_analyse_params *ap_bw, const mvb200_analyse_params *ap_fw, const mvb200_flowblur_params *p, const void *host_src, int src_first,
                                         int n, int out_first, int out_count, void *host_dst) {
    MockSegment S;
    const mvb200_analyse_params *aps[2] = {ap_bw, ap_fw};
    if (S.build(sp, aps, 2, host_src, n))
        return 1;
    const int off = p->vd.nDeltaFrame;
    for (int k = 0; k < out_count; k++) {
        const int f = out_first - src_first + k;
        uint8_t *d[3];
        ptrdiff_t dstr[3];
        S.dst_planes(host_dst, k, d, dstr);
        int32_t usable = 0;
        if (f - off >= 0 && f + off < n) {
            const uint8_t *s_[3];
            S.sup_planes(f, s_);
            if (fwd(mvo_flowblur_frame(p, sp, s_, S.sstride, S.vec[0][f - off].data(), S.vec[1][f + off].data(), d, dstr, &usable)))
                return 1;
        }
        if (!usable)
            S.copy_src(f, d, dstr);
    }
    return 0;
}

MVB200_API int mvb200_blockfps_clip_host(mvb200_ctx *, const mvb200_super_params *sp, const mvb200_analyse_params *ap_bw, const mvb200_analyse_params *ap_fw, const mvb200_blockfps_params *p, const void *host_src, int src_first,
                                         int n, int out_first, int out_count, void *host_dst, int32_t *usable_out) {
    MockSegment S;
    const mvb200_analyse_params *aps[2] = {ap_bw, ap_fw};
    if (S.build(sp, aps, 2, host_src, n))
        return 1;
    const int off = p->vd.nDeltaFrame, N = p->num_frames_in;
    auto in_seg = [&](int f) { return f >= src_first && f < src_first + n; };
    for (int k = 0; k < out_count; k++) {
        int32_t nleft, time256;
        mvb200_blockfps_position(p, out_first + k, &nleft, &time256);
        const int nright = nleft + off, fl = nleft < N - 1 ? nleft : N - 1, fr = nright < N - 1 ? nright : N - 1;
        uint8_t *d[3];
        ptrdiff_t dstr[3];
        S.dst_planes(host_dst, k, d, dstr);
        if (usable_out)
            usable_out[k] = 0;
        if (time256 == 0 || time256 == 256) {
            const int f = time256 == 0 ? fl : fr;
            if (!in_seg(f))
                return mvb200::set_error("mock: source frame outside the segment");
            S.copy_src(f - src_first, d, dstr);
            continue;
        }
        if (!in_seg(fl) || !in_seg(fr))
            return mvb200::set_error("mock: source frame outside the segment");
        int32_t usable = 0;
        if (nleft < N && nright < N) {
            const int L = nleft - src_first, R = nright - src_first;
            const uint8_t *sl[3], *sr[3];
            S.sup_planes(L, sl);
            S.sup_planes(R, sr);
            if (fwd(mvo_blockfps_frame(p, sp, time256, sl, S.sstride, sr, S.sstride, S.vec[0][L].data(), S.vec[1][R].data(), d, dstr, &usable)))
                return 1;
        }
        if (usable_out)
            usable_out[k] = usable;
        if (!usable) {
            if (p->blend) {
                const uint8_t *ql[3], *qr[3];
                ptrdiff_t stl[3], str_[3];
                S.src_planes(fl - src_first, ql, stl);
                S.src_planes(fr - src_first, qr, str_);
                if (fwd(mvo_blend_frame(sp, time256, ql, stl, qr, str_, d, dstr)))
                    return 1;
            } else {
                S.copy_src(fl - src_first, d, dstr);
            }
        }
    }
    return 0;
}

} // extern "C"

// (the stand-in computes at once: "asynchronous" is the synchronous call)
MVB200_API int mvb200_compensate_recalc_clip_host_async(mvb200_ctx *ctx, const mvb200_super_params *sp, const mvb200_analyse_params *ap, const mvb200_recalculate_params *const *rp, int nrp,
                                                        const mvb200_compensate_params *cp, const void *host_src, int src_first, int n, int out_first, int out_count, void *host_dst) {
    return mvb200_compensate_recalc_clip_host(ctx, sp, ap, rp, nrp, cp, host_src, src_first, n, out_first, out_count, host_dst);
}
MVB200_API int mvb200_analyse_clip_host_async(mvb200_ctx *ctx, const mvb200_super_params *sp, const mvb200_analyse_params *ap, const void *host_src, int src_first, int n, int out_first, int out_count,
                                              void *host_super, void *host_vectors, int64_t vectors_stride) {
    return mvb200_analyse_clip_host(ctx, sp, ap, host_src, src_first, n, out_first, out_count, host_super, host_vectors, vectors_stride);
}
