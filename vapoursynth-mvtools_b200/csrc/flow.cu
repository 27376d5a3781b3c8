// The consumers of the vector clips that SURVEY.md section 8f ranks next: mv.SCDetection, mv.Mask, mv.Finest and
// mv.Flow (MVSCDetection.c, MVMask.c + MaskFun.cpp, MVFinest.c, MVFlow.cpp + SimpleResize.cpp of the reference).
//
// All of them are HBM-bound per-pixel (or per-block) maps with no serial chain.  The reference stages full-size
// intermediates (int16 vector planes, a pel-times upsized "finest" clip); here every output sample is computed from
// the per-block field (a few KB, L1/L2 resident) and the reference SUPER frame in one pass: algorithmic bytes per
// output frame are one read of the touched reference samples + one write of the output (Flow, fetch mode:
// 2 P; Mask: P8 written only; Finest: S0 read + S0 written, S0 = the pel^2 level-0 sub-planes).
// The per-thread arithmetic lives in flow_device.h (also compiled for the host by the emulation test).
#include "engine.h"
#include "flow_device.h"
#include "flow_setup.h"

#include <cstring>
#include <vector>

namespace mvb200 {

namespace {

using namespace flowdev;

// fgopIsUsable (Fakery.c:52-58,144-146) for blob blockIdx.x: usable[i] = valid && !(count(sad > nSCD1) > nSCD2) && haveRef
__global__ void k_flow_scd(const uint8_t *vectors, long long stride, int nBlk, long long nSCD1, int nSCD2, int haveRef, int *usable) {
    const uint8_t *blob = vectors + blockIdx.x * stride;
    const int *l0 = blob_level0(blob, nBlk);
    int cnt = 0;
    for (int i = threadIdx.x; i < nBlk; i += blockDim.x)
        cnt += vec_sad(l0 + 4 * i) > nSCD1 ? 1 : 0;
    __shared__ int s;
    if (threadIdx.x == 0)
        s = 0;
    __syncthreads();
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0)
        atomicAdd(&s, cnt);
    __syncthreads();
    if (threadIdx.x == 0)
        usable[blockIdx.x] = (((const int *)blob)[1] == 1) && !(s > nSCD2) && haveRef;
}

// ------------------------------------------------------------------- Flow

struct SmallFields {
    short *xY, *yY, *xUV, *yUV;
};

__global__ void k_small_vectors(const uint8_t *blob, int nBlkX, int nBlkY, int nBlkXP, int nBlkYP, int fieldShift, int xRatioUV, int yRatioUV, SmallFields out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nBlkXP * nBlkYP)
        return;
    const SmallVec s = small_vector(blob_level0(blob, nBlkX * nBlkY), nBlkX, nBlkY, k % nBlkXP, k / nBlkXP, fieldShift, xRatioUV, yRatioUV);
    out.xY[k] = s.xY;
    out.yY[k] = s.yY;
    out.xUV[k] = s.xUV;
    out.yUV[k] = s.yUV;
}

// fetch mode; an unusable frame is a copy of the source clip's frame when one is given (batch path)
template <typename T>
__global__ void k_flow_fetch(FlowPlane F, const uint8_t *__restrict__ ref_super, uint8_t *__restrict__ dst, long long dst_pitch, const uint8_t *__restrict__ src, const int *__restrict__ usable) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= F.W || y >= F.H)
        return;
    T *d = (T *)(dst + (long long)y * dst_pitch) + x;
    if (!*usable) {
        if (src)
            *d = *((const T *)(src + (long long)y * dst_pitch) + x);
        return;
    }
    *d = flow_fetch_sample<T>(F, ref_super, x, y);
}

// shift mode, pass 1: every source sample claims its destination; the largest source index wins (= the last writer
// of the reference's raster loop)
__global__ void k_flow_shift_mark(FlowPlane F, int *__restrict__ owner, const int *__restrict__ usable) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= F.W || y >= F.H || !*usable)
        return;
    const int t = flow_shift_target(F, x, y);
    if (t >= 0)
        atomicMax(owner + t, y * F.W + x + 1);
}

// pass 2: destinations nobody claimed keep the fill value (pixel_max, MVFlow.cpp:310-311)
template <typename T>
__global__ void k_flow_shift_resolve(FlowPlane F, const uint8_t *__restrict__ ref_super, uint8_t *__restrict__ dst, long long dst_pitch, const uint8_t *__restrict__ src, const int *__restrict__ owner, int pixel_max,
                                     const int *__restrict__ usable) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= F.W || y >= F.H)
        return;
    T *d = (T *)(dst + (long long)y * dst_pitch) + x;
    if (!*usable) {
        if (src)
            *d = *((const T *)(src + (long long)y * dst_pitch) + x);
        return;
    }
    const int o = owner[y * F.W + x];
    if (o == 0) {
        *d = (T)pixel_max;
        return;
    }
    const int sx = (o - 1) % F.W, sy = (o - 1) / F.W;
    *d = finest_sample<T>(ref_super, F.S, F.logPel, (sx + F.S.hpad) << F.logPel, (sy + F.S.vpad) << F.logPel);
}

// ------------------------------------------------------------------- Mask

__global__ void k_mask_small(MaskSmallParams P, const uint8_t *blob, int *small, int *smallV) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.nBlkX * P.nBlkY)
        return;
    const int *l0 = blob_level0(blob, P.nBlkX * P.nBlkY);
    if (P.kind == 2) {
        occlusion_contrib(l0, P.isBackward, P.nBlkX, P.nBlkY, 1.0 / P.fMaskNormFactor, P.fGamma, P.nPel, P.time256, P.stepX, P.stepY, j % P.nBlkX, j / P.nBlkX, [small](int idx, int val) { atomicMax(small + idx, val); });
    } else {
        mask_small_direct(P, l0, j, small, smallV);
    }
}

__global__ void k_mask_plane(MaskPlane B, int kind, int ysc, int nBlkX, const int *__restrict__ small, const int *__restrict__ smallV, const uint8_t *__restrict__ src_luma, long long src_pitch, uint8_t *__restrict__ dst,
                             long long dst_pitch, const int *__restrict__ usable) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= B.W || y >= B.H)
        return;
    dst[(long long)y * dst_pitch + x] = (uint8_t)mask_sample(B, kind, *usable, ysc, nBlkX, small, smallV, src_luma, src_pitch, x, y);
}

// ----------------------------------------------------------------- Finest

template <typename T>
__global__ void k_finest(SuperPlane S, int logPel, int W, int H, const uint8_t *__restrict__ super_frame, uint8_t *__restrict__ dst, long long dst_pitch) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H)
        return;
    *((T *)(dst + (long long)y * dst_pitch) + x) = finest_sample<T>(super_frame, S, logPel, x, y);
}

// -------------------------------------------------------------- FlowInter

// MakeVectorOcclusionMaskTime into a zeroed nBlkX x nBlkY int array (gamma = 1: MVFlowInter.c:203-206)
__global__ void k_occ_small(const uint8_t *blob, int isBackward, int nBlkX, int nBlkY, double ml, int nPel, int time256, int stepX, int stepY, int *raw) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nBlkX * nBlkY)
        return;
    occlusion_contrib(blob_level0(blob, nBlkX * nBlkY), isBackward, nBlkX, nBlkY, ml, 1.0, nPel, time256, stepX, stepY, j % nBlkX, j / nBlkX, [raw](int idx, int val) { atomicMax(raw + idx, val); });
}

__global__ void k_pad_mask(const int *raw, int nBlkX, int nBlkY, int nBlkXP, int nBlkYP, int *out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nBlkXP * nBlkYP)
        out[k] = padded_mask_entry(raw, nBlkX, nBlkY, k % nBlkXP, k / nBlkXP);
}

// flags: usable(vec_b), usable(vec_f), usable(vec_bb), usable(vec_ff).  Not usable: the blend of (or the copy from) the
// SOURCE clip's frames when they are given (batch path), else dst is left alone.
template <typename T>
__global__ void k_flow_inter(InterPlane P, int mode_primary, int mode_fallback, const uint8_t *__restrict__ right_super, const uint8_t *__restrict__ left_super, uint8_t *__restrict__ dst, long long dst_pitch,
                             const uint8_t *__restrict__ src, const uint8_t *__restrict__ ref, int blend, const int *__restrict__ flags) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= P.W || y >= P.H)
        return;
    T *d = (T *)(dst + (long long)y * dst_pitch) + x;
    if (!(flags[0] && flags[1])) {
        if (src) {
            const int a = *((const T *)(src + (long long)y * dst_pitch) + x);
            *d = (T)(blend ? blend_sample(a, *((const T *)(ref + (long long)y * dst_pitch) + x), P.time256) : a);
        }
        return;
    }
    *d = (T)flow_inter_sample<T>(P, (flags[2] && flags[3]) ? mode_primary : mode_fallback, right_super, left_super, x, y);
}

template <typename T>
__global__ void k_blend(int W, int H, int time256, const uint8_t *__restrict__ src, const uint8_t *__restrict__ ref, uint8_t *__restrict__ dst, long long pitch) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H)
        return;
    const long long o = (long long)y * pitch;
    *((T *)(dst + o) + x) = (T)blend_sample(*((const T *)(src + o) + x), *((const T *)(ref + o) + x), time256);
}

// What the reference returns when the vectors are not usable (MVBlockFPS.c:636-670): the blend of the two source frames, or
// the left one.  flags (may be null = unconditionally): two ints, both set = the frame was interpolated, leave it.
template <typename T>
__global__ void k_fallback(int W, int H, int time256, int blend, const uint8_t *__restrict__ src, const uint8_t *__restrict__ ref, uint8_t *__restrict__ dst, long long pitch, const int *__restrict__ flags) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H || (flags && flags[0] && flags[1]))
        return;
    const long long o = (long long)y * pitch;
    const int a = *((const T *)(src + o) + x);
    *((T *)(dst + o) + x) = (T)(blend ? blend_sample(a, *((const T *)(ref + o) + x), time256) : a);
}

// --------------------------------------------------------------- FlowBlur
template <typename T>
__global__ void k_flow_blur(BlurPlane P, const uint8_t *__restrict__ super_frame, uint8_t *__restrict__ dst, long long dst_pitch, const uint8_t *__restrict__ src, const int *__restrict__ flags) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= P.W || y >= P.H)
        return;
    T *d = (T *)(dst + (long long)y * dst_pitch) + x;
    if (!(flags[0] && flags[1])) {
        if (src)
            *d = *((const T *)(src + (long long)y * dst_pitch) + x);
        return;
    }
    *d = (T)flow_blur_sample<T>(P, super_frame, x, y);
}

// --------------------------------------------------------------- BlockFPS
__global__ void k_sad_small(const uint8_t *blob, int nBlkX, int nBlkY, double dSADNormFactor, int nPel, int time256, int stepX, int stepY, int shift, int *raw) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < nBlkX * nBlkY)
        raw[j] = sad_mask_entry(blob_level0(blob, nBlkX * nBlkY), nBlkX, nBlkY, j, dSADNormFactor, nPel, time256, stepX, stepY, shift);
}

__global__ void k_mult_masks(const int *f, const int *b, int n, int *o) { // MultMasks, MVBlockFPS.c:90-98
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n)
        o[k] = (f[k] * b[k]) / 255;
}

template <typename T>
__global__ void k_blockfps(BfpsPlane P, const uint8_t *__restrict__ vec_b, const uint8_t *__restrict__ vec_f, const uint8_t *__restrict__ right_super, const uint8_t *__restrict__ left_super, uint8_t *__restrict__ dst,
                           long long dst_pitch, const int *__restrict__ flags) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= P.W || y >= P.H || !(flags[0] && flags[1]))
        return;
    const int nb = P.nBlkX * P.nBlkY;
    *((T *)(dst + (long long)y * dst_pitch) + x) = (T)bfps_sample<T>(P, blob_level0(vec_b, nb), blob_level0(vec_f, nb), right_super, left_super, x, y);
}

// ------------------------------------------------------------ host helpers

constexpr int kSlotTables = 16, kSlotSmall = 17, kSlotStage = 18;

SuperPlane super_plane_of(const SuperDesc &sd, int plane) { return make_super_plane(sd.pitch[plane], sd.offset[plane], sd.lv[plane][0]); }

// The SimpleResize tables of (luma, chroma) upsizers with a common source size, resident on the device until the
// configuration changes (first use: built on the host -- InitTables is float arithmetic -- and uploaded synchronously).
int get_resize_tables(Ctx *c, int dw0, int dh0, int dw1, int dh1, int sw, int sh, ResizeTables t[2]) {
    const int key[8] = {dw0, dh0, dw1, dh1, sw, sh, 1, 0};
    const size_t n0 = tables_ints(dw0, dh0), n1 = tables_ints(dw1, dh1);
    if (!c->rz_valid || memcmp(key, c->rz_key, sizeof(key)) != 0) {
        std::vector<int32_t> host(n0 + n1);
        fill_tables(host.data(), dw0, dh0, sw, sh);
        if (n1)
            fill_tables(host.data() + n0, dw1, dh1, sw, sh);
        int *dev = (int *)c->get_scratch(kSlotTables, (n0 + n1) * sizeof(int32_t) + 256);
        if (!dev)
            return 1;
        MVB_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        MVB_CUDA_CHECK(upload_table(c, dev, host.data(), (n0 + n1) * sizeof(int32_t)));
        c->rz_dev = dev;
        memcpy(c->rz_key, key, sizeof(key));
        c->rz_valid = true;
    }
    t[0] = tables_at(c->rz_dev, dw0, dh0);
    t[1] = tables_at(c->rz_dev + n0, dw1, dh1);
    return 0;
}

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

} // namespace

// mvflowGetFrame, MVFlow.cpp:157-368, for one frame on the device.  dev_ref_super == nullptr: not usable.
// dev_src (clip layout, may be null): what an unusable frame becomes.  usable_dev: where the flag is left.
static int launch_flow(Ctx *c, const mvb200_flow_params *fp, const SuperDesc &sd, const ClipDesc &cd, const uint8_t *dev_ref_super, const uint8_t *dev_vectors, int field_shift, const uint8_t *dev_src, uint8_t *dev_dst, int **usable_dev) {
    cudaStream_t st = c->stream;
    const mvb200_analysis_data &vd = fp->vd;
    const int ns = fp->nBlkXP * fp->nBlkYP;
    const size_t field_bytes = align256((size_t)ns * sizeof(short));
    const size_t owner_bytes = fp->mode == 1 ? align256((size_t)vd.nWidth * vd.nHeight * sizeof(int)) : 0;
    uint8_t *scr = (uint8_t *)c->get_scratch(kSlotSmall, 256 + 4 * field_bytes + owner_bytes);
    if (!scr)
        return 1;
    int *usable = (int *)scr;
    SmallFields sf;
    sf.xY = (short *)(scr + 256);
    sf.yY = (short *)(scr + 256 + field_bytes);
    sf.xUV = (short *)(scr + 256 + 2 * field_bytes);
    sf.yUV = (short *)(scr + 256 + 3 * field_bytes);
    int *owner = (int *)(scr + 256 + 4 * field_bytes);
    *usable_dev = usable;
    ResizeTables t[2];
    if (get_resize_tables(c, fp->nWidthP, fp->nHeightP, fp->nplanes > 1 ? fp->nWidthPUV : 0, fp->nplanes > 1 ? fp->nHeightPUV : 0, fp->nBlkXP, fp->nBlkYP, t))
        return 1;
    k_flow_scd<<<1, 256, 0, st>>>(dev_vectors, 0, vd.nBlkX * vd.nBlkY, fp->nSCD1, fp->nSCD2, dev_ref_super != nullptr, usable);
    c->launches++;
    k_small_vectors<<<(ns + 255) / 256, 256, 0, st>>>(dev_vectors, vd.nBlkX, vd.nBlkY, fp->nBlkXP, fp->nBlkYP, field_shift, vd.xRatioUV, vd.yRatioUV, sf);
    c->launches++;
    const uint8_t *ref = dev_ref_super ? dev_ref_super : dev_vectors; // (never dereferenced when unusable)
    for (int plane = 0; plane < fp->nplanes; plane++) {
        const FlowPlane F = make_flow_plane(fp, plane, plane ? sf.xUV : sf.xY, plane ? sf.yUV : sf.yY, t[plane ? 1 : 0], super_plane_of(sd, plane));
        uint8_t *dst = dev_dst + cd.offset[plane];
        const uint8_t *src = dev_src ? dev_src + cd.offset[plane] : nullptr;
        dim3 blk(32, 8), grid((F.W + 31) / 32, (F.H + 7) / 8);
        if (fp->mode == 0) {
            if (sd.bps == 1)
                k_flow_fetch<uint8_t><<<grid, blk, 0, st>>>(F, ref, dst, cd.pitch[plane], src, usable);
            else
                k_flow_fetch<uint16_t><<<grid, blk, 0, st>>>(F, ref, dst, cd.pitch[plane], src, usable);
            c->launches++;
        } else {
            MVB_CUDA_CHECK(cudaMemsetAsync(owner, 0, (size_t)F.W * F.H * sizeof(int), st));
            k_flow_shift_mark<<<grid, blk, 0, st>>>(F, owner, usable);
            if (sd.bps == 1)
                k_flow_shift_resolve<uint8_t><<<grid, blk, 0, st>>>(F, ref, dst, cd.pitch[plane], src, owner, (1 << fp->bits) - 1, usable);
            else
                k_flow_shift_resolve<uint16_t><<<grid, blk, 0, st>>>(F, ref, dst, cd.pitch[plane], src, owner, (1 << fp->bits) - 1, usable);
            c->launches += 2;
        }
    }
    MVB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// mvmaskGetFrame, MVMask.c:74-203, for one frame on the device (dst: mvb200_mask_layout)
static int launch_mask(Ctx *c, const mvb200_mask_params *mp, const mvb200_frame_layout &ml, const uint8_t *dev_src_luma, long long src_pitch, const uint8_t *dev_vectors, uint8_t *dev_dst) {
    cudaStream_t st = c->stream;
    const mvb200_analysis_data &vd = mp->vd;
    const int n = vd.nBlkX * vd.nBlkY;
    const size_t small_bytes = align256((size_t)n * sizeof(int));
    uint8_t *scr = (uint8_t *)c->get_scratch(kSlotSmall, 256 + 2 * small_bytes);
    if (!scr)
        return 1;
    int *usable = (int *)scr, *small = (int *)(scr + 256), *smallV = (int *)(scr + 256 + small_bytes);
    ResizeTables t[2];
    if (get_resize_tables(c, mp->nWidthB, mp->nHeightB, mp->nWidthBUV, mp->nHeightBUV, vd.nBlkX, vd.nBlkY, t))
        return 1;
    if (mp->kind == 5 && !dev_src_luma)
        return set_error("Mask: kind=5 needs the source clip's luma plane");
    k_flow_scd<<<1, 256, 0, st>>>(dev_vectors, 0, n, mp->nSCD1, mp->nSCD2, 1, usable);
    c->launches++;
    const MaskSmallParams P = make_mask_small_params(mp);
    if (mp->kind == 2)
        MVB_CUDA_CHECK(cudaMemsetAsync(small, 0, (size_t)n * sizeof(int), st));
    k_mask_small<<<(n + 127) / 128, 128, 0, st>>>(P, dev_vectors, small, smallV);
    c->launches++;
    for (int plane = 0; plane < 3; plane++) {
        const MaskPlane B = make_mask_plane(mp, plane, t[plane ? 1 : 0]);
        dim3 blk(32, 8), grid((B.W + 31) / 32, (B.H + 7) / 8);
        k_mask_plane<<<grid, blk, 0, st>>>(B, mp->kind, mp->nSceneChangeValue, vd.nBlkX, small, smallV, dev_src_luma, src_pitch, dev_dst + ml.offset[plane], ml.pitch[plane], usable);
        c->launches++;
    }
    MVB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// mvfinestGetFrame, MVFinest.c:47-125, for one frame on the device (dst: mvb200_finest_layout)
static int launch_finest(Ctx *c, const mvb200_super_params *sp, const SuperDesc &sd, const mvb200_frame_layout &fl, const uint8_t *dev_super, uint8_t *dev_dst) {
    const int lp = ilog2_pel(sp->pel);
    for (int plane = 0; plane < sd.nplanes; plane++) {
        if (plane && !(sp->modeyuv & (MVB200_UPLANE | MVB200_VPLANE)))
            continue; // (the reference leaves planes the super clip does not carry untouched)
        const SuperPlane S = super_plane_of(sd, plane);
        const int W = S.padded_w * sp->pel, H = S.padded_h * sp->pel;
        dim3 blk(32, 8), grid((W + 31) / 32, (H + 7) / 8);
        if (sd.bps == 1)
            k_finest<uint8_t><<<grid, blk, 0, c->stream>>>(S, lp, W, H, dev_super, dev_dst + fl.offset[plane], fl.pitch[plane]);
        else
            k_finest<uint16_t><<<grid, blk, 0, c->stream>>>(S, lp, W, H, dev_super, dev_dst + fl.offset[plane], fl.pitch[plane]);
        c->launches++;
    }
    MVB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// mvflowinterGetFrame (MVFlowInter.c:83-448) / the core of mvflowfpsGetFrame for one output frame on the device.
// Any of the blobs / super frames may be null (frame out of range): the frame is then not usable.
static int launch_flowinter(Ctx *c, const mvb200_flowinter_params *p, const SuperDesc &sd, const ClipDesc &cd, int time256, const uint8_t *left_super, const uint8_t *right_super, const uint8_t *vec_b, const uint8_t *vec_f,
                            const uint8_t *vec_bb, const uint8_t *vec_ff, const uint8_t *dev_src, const uint8_t *dev_ref, uint8_t *dev_dst, int **flags_dev) {
    cudaStream_t st = c->stream;
    const mvb200_analysis_data &vd = p->vd;
    const int ns = p->nBlkXP * p->nBlkYP, nb = vd.nBlkX * vd.nBlkY;
    const size_t field_bytes = align256((size_t)ns * sizeof(short)), mask_bytes = align256((size_t)ns * sizeof(int)), raw_bytes = align256((size_t)nb * sizeof(int));
    uint8_t *scr = (uint8_t *)c->get_scratch(kSlotSmall, 256 + 16 * field_bytes + 2 * mask_bytes + 2 * raw_bytes);
    if (!scr)
        return 1;
    int *flags = (int *)scr;
    *flags_dev = flags;
    short *fld[16];
    for (int i = 0; i < 16; i++)
        fld[i] = (short *)(scr + 256 + i * field_bytes);
    int *mask[2] = {(int *)(scr + 256 + 16 * field_bytes), (int *)(scr + 256 + 16 * field_bytes + mask_bytes)};
    int *raw[2] = {(int *)(scr + 256 + 16 * field_bytes + 2 * mask_bytes), (int *)(scr + 256 + 16 * field_bytes + 2 * mask_bytes + raw_bytes)};
    ResizeTables t[2];
    if (get_resize_tables(c, p->nWidthP, p->nHeightP, p->nplanes > 1 ? p->nWidthPUV : 0, p->nplanes > 1 ? p->nHeightPUV : 0, p->nBlkXP, p->nBlkYP, t))
        return 1;
    const bool frames_ok = left_super && right_super && vec_b && vec_f;
    const uint8_t *blobs[4] = {vec_b, vec_f, vec_bb, vec_ff};
    for (int i = 0; i < 4; i++) {
        const bool have = blobs[i] != nullptr && frames_ok;
        if (have) {
            k_flow_scd<<<1, 256, 0, st>>>(blobs[i], 0, nb, p->nSCD1, p->nSCD2, 1, flags + i);
            c->launches++;
            SmallFields sf = {fld[4 * i], fld[4 * i + 1], fld[4 * i + 2], fld[4 * i + 3]};
            k_small_vectors<<<(ns + 255) / 256, 256, 0, st>>>(blobs[i], vd.nBlkX, vd.nBlkY, p->nBlkXP, p->nBlkYP, 0, vd.xRatioUV, vd.yRatioUV, sf);
            c->launches++;
        } else {
            MVB_CUDA_CHECK(cudaMemsetAsync(flags + i, 0, sizeof(int), st));
        }
    }
    if (frames_ok) {
        const int stepX = vd.nBlkSizeX - vd.nOverlapX, stepY = vd.nBlkSizeY - vd.nOverlapY;
        MVB_CUDA_CHECK(cudaMemsetAsync(raw[0], 0, 2 * raw_bytes, st));
        k_occ_small<<<(nb + 127) / 128, 128, 0, st>>>(vec_b, 1, vd.nBlkX, vd.nBlkY, p->ml, vd.nPel, 256 - time256, stepX, stepY, raw[0]);
        k_occ_small<<<(nb + 127) / 128, 128, 0, st>>>(vec_f, 0, vd.nBlkX, vd.nBlkY, p->ml, vd.nPel, time256, stepX, stepY, raw[1]);
        k_pad_mask<<<(ns + 255) / 256, 256, 0, st>>>(raw[0], vd.nBlkX, vd.nBlkY, p->nBlkXP, p->nBlkYP, mask[0]);
        k_pad_mask<<<(ns + 255) / 256, 256, 0, st>>>(raw[1], vd.nBlkX, vd.nBlkY, p->nBlkXP, p->nBlkYP, mask[1]);
        c->launches += 4;
    }
    if (!frames_ok && !dev_src) { // nothing to write: the caller handles the unusable frame
        MVB_CUDA_CHECK(cudaGetLastError());
        return 0;
    }
    for (int plane = 0; plane < p->nplanes; plane++) {
        const int u = plane ? 2 : 0; // fields: [4 i + 0/1] = luma x / y, [4 i + 2/3] = chroma x / y of blob i (b, f, bb, ff)
        const InterSmall v = {fld[0 + u], fld[1 + u], fld[4 + u], fld[5 + u], fld[8 + u], fld[9 + u], fld[12 + u], fld[13 + u]};
        const InterPlane P = make_inter_plane(p, time256, plane, v, mask[0], mask[1], t[plane ? 1 : 0], super_plane_of(sd, plane));
        uint8_t *dst = dev_dst + cd.offset[plane];
        const uint8_t *src = dev_src ? dev_src + cd.offset[plane] : nullptr, *ref = dev_ref ? dev_ref + cd.offset[plane] : src;
        const uint8_t *ls = left_super ? left_super : (const uint8_t *)flags, *rs = right_super ? right_super : (const uint8_t *)flags;
        dim3 blk(32, 8), grid((P.W + 31) / 32, (P.H + 7) / 8);
        if (sd.bps == 1)
            k_flow_inter<uint8_t><<<grid, blk, 0, st>>>(P, p->mode_primary, p->mode_fallback, rs, ls, dst, cd.pitch[plane], src, ref, p->blend, flags);
        else
            k_flow_inter<uint16_t><<<grid, blk, 0, st>>>(P, p->mode_primary, p->mode_fallback, rs, ls, dst, cd.pitch[plane], src, ref, p->blend, flags);
        c->launches++;
    }
    MVB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// mvflowblurGetFrame (MVFlowBlur.c:144-337) for one frame on the device; vec_b / vec_f null: not usable
static int launch_flowblur(Ctx *c, const mvb200_flowblur_params *p, const SuperDesc &sd, const ClipDesc &cd, const uint8_t *super_frame, const uint8_t *vec_b, const uint8_t *vec_f, const uint8_t *dev_src, uint8_t *dev_dst,
                           int **flags_dev) {
    cudaStream_t st = c->stream;
    const mvb200_analysis_data &vd = p->vd;
    const int nb = vd.nBlkX * vd.nBlkY;
    const size_t field_bytes = align256((size_t)nb * sizeof(short));
    uint8_t *scr = (uint8_t *)c->get_scratch(kSlotSmall, 256 + 8 * field_bytes);
    if (!scr)
        return 1;
    int *flags = (int *)scr;
    *flags_dev = flags;
    short *fld[8];
    for (int i = 0; i < 8; i++)
        fld[i] = (short *)(scr + 256 + i * field_bytes);
    ResizeTables t[2];
    if (get_resize_tables(c, vd.nWidth, vd.nHeight, p->nplanes > 1 ? p->nWidthUV : 0, p->nplanes > 1 ? p->nHeightUV : 0, vd.nBlkX, vd.nBlkY, t))
        return 1;
    const bool ok = vec_b && vec_f;
    const uint8_t *blobs[2] = {vec_b, vec_f};
    for (int i = 0; i < 2; i++) {
        if (ok) {
            k_flow_scd<<<1, 256, 0, st>>>(blobs[i], 0, nb, p->nSCD1, p->nSCD2, 1, flags + i);
            SmallFields sf = {fld[4 * i], fld[4 * i + 1], fld[4 * i + 2], fld[4 * i + 3]};
            k_small_vectors<<<(nb + 255) / 256, 256, 0, st>>>(blobs[i], vd.nBlkX, vd.nBlkY, vd.nBlkX, vd.nBlkY, 0, vd.xRatioUV, vd.yRatioUV, sf);
            c->launches += 2;
        } else {
            MVB_CUDA_CHECK(cudaMemsetAsync(flags + i, 0, sizeof(int), st));
        }
    }
    if (!ok && !dev_src) {
        MVB_CUDA_CHECK(cudaGetLastError());
        return 0;
    }
    for (int plane = 0; plane < p->nplanes; plane++) {
        const int u = plane ? 2 : 0;
        const BlurPlane P = make_blur_plane(p, plane, fld[0 + u], fld[1 + u], fld[4 + u], fld[5 + u], t[plane ? 1 : 0], super_plane_of(sd, plane));
        const uint8_t *src = dev_src ? dev_src + cd.offset[plane] : nullptr;
        dim3 blk(32, 8), grid((P.W + 31) / 32, (P.H + 7) / 8);
        if (sd.bps == 1)
            k_flow_blur<uint8_t><<<grid, blk, 0, st>>>(P, super_frame, dev_dst + cd.offset[plane], cd.pitch[plane], src, flags);
        else
            k_flow_blur<uint16_t><<<grid, blk, 0, st>>>(P, super_frame, dev_dst + cd.offset[plane], cd.pitch[plane], src, flags);
        c->launches++;
    }
    MVB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

void overlap_windows_host(int nx, int ny, int ox, int oy, std::vector<short> &out); // degrain.cu

// mvblockfpsGetFrame (MVBlockFPS.c:240-640), usable case, for one output frame on the device
static int launch_blockfps(Ctx *c, const mvb200_blockfps_params *p, const SuperDesc &sd, const ClipDesc &cd, int time256, const uint8_t *left_super, const uint8_t *right_super, const uint8_t *vec_b, const uint8_t *vec_f, uint8_t *dev_dst,
                           int **flags_dev) {
    cudaStream_t st = c->stream;
    const mvb200_analysis_data &vd = p->vd;
    const int ns = p->nBlkXP * p->nBlkYP, nb = vd.nBlkX * vd.nBlkY;
    const int planes = (p->super_modeyuv & (MVB200_UPLANE | MVB200_VPLANE)) ? 3 : 1;
    const size_t mask_bytes = align256((size_t)ns * sizeof(int)), raw_bytes = align256((size_t)nb * sizeof(int));
    uint8_t *scr = (uint8_t *)c->get_scratch(kSlotSmall, 256 + 3 * mask_bytes + 2 * raw_bytes);
    if (!scr)
        return 1;
    int *flags = (int *)scr;
    *flags_dev = flags;
    int *mask[3] = {(int *)(scr + 256), (int *)(scr + 256 + mask_bytes), (int *)(scr + 256 + 2 * mask_bytes)}; // B, F, O
    int *raw[2] = {(int *)(scr + 256 + 3 * mask_bytes), (int *)(scr + 256 + 3 * mask_bytes + raw_bytes)};
    ResizeTables t[2];
    if (get_resize_tables(c, p->nWidthP, p->nHeightP, planes > 1 ? p->nWidthPUV : 0, planes > 1 ? p->nHeightPUV : 0, p->nBlkXP, p->nBlkYP, t))
        return 1;
    const bool overlap = vd.nOverlapX || vd.nOverlapY;
    if (overlap) { // the nine windows per plane type, resident until the configuration changes
        const int key[8] = {vd.nBlkSizeX, vd.nBlkSizeY, vd.nOverlapX, vd.nOverlapY, planes > 1 ? vd.xRatioUV : 0, planes > 1 ? vd.yRatioUV : 0, 1, 0};
        if (!c->bfw_valid || memcmp(key, c->bfw_key, sizeof(key)) != 0) {
            std::vector<short> wl, wc;
            overlap_windows_host(vd.nBlkSizeX, vd.nBlkSizeY, vd.nOverlapX, vd.nOverlapY, wl);
            if (planes > 1)
                overlap_windows_host(vd.nBlkSizeX / vd.xRatioUV, vd.nBlkSizeY / vd.yRatioUV, vd.nOverlapX / vd.xRatioUV, vd.nOverlapY / vd.yRatioUV, wc);
            const size_t bl = wl.size() * sizeof(short), bc = wc.size() * sizeof(short), offc = align256(bl);
            uint8_t *dev = (uint8_t *)c->get_scratch(19, offc + bc + 256);
            if (!dev)
                return 1;
            MVB_CUDA_CHECK(cudaStreamSynchronize(st));
            MVB_CUDA_CHECK(upload_table(c, dev, wl.data(), bl));
            if (bc)
                MVB_CUDA_CHECK(upload_table(c, dev + offc, wc.data(), bc));
            c->bfw_dev[0] = (short *)dev;
            c->bfw_dev[1] = (short *)(dev + offc);
            memcpy(c->bfw_key, key, sizeof(key));
            c->bfw_valid = true;
        }
    }
    k_flow_scd<<<1, 256, 0, st>>>(vec_b, 0, nb, p->nSCD1, p->nSCD2, 1, flags);
    k_flow_scd<<<1, 256, 0, st>>>(vec_f, 0, nb, p->nSCD1, p->nSCD2, 1, flags + 1);
    c->launches += 2;
    if (p->mode >= 3) { // :407-440
        const int stepX = vd.nBlkSizeX - vd.nOverlapX, stepY = vd.nBlkSizeY - vd.nOverlapY;
        if (p->mode <= 5) {
            MVB_CUDA_CHECK(cudaMemsetAsync(raw[0], 0, 2 * raw_bytes, st));
            k_occ_small<<<(nb + 127) / 128, 128, 0, st>>>(vec_b, 1, vd.nBlkX, vd.nBlkY, p->ml, vd.nPel, 256 - time256, stepX, stepY, raw[0]);
            k_occ_small<<<(nb + 127) / 128, 128, 0, st>>>(vec_f, 0, vd.nBlkX, vd.nBlkY, p->ml, vd.nPel, time256, stepX, stepY, raw[1]);
        } else {
            const double f = 4.0 / (p->ml * vd.nBlkSizeX * vd.nBlkSizeY);
            k_sad_small<<<(nb + 127) / 128, 128, 0, st>>>(vec_b, vd.nBlkX, vd.nBlkY, f, vd.nPel, 256 - time256, stepX, stepY, p->bits - 8, raw[0]);
            k_sad_small<<<(nb + 127) / 128, 128, 0, st>>>(vec_f, vd.nBlkX, vd.nBlkY, f, vd.nPel, time256, stepX, stepY, p->bits - 8, raw[1]);
        }
        k_pad_mask<<<(ns + 255) / 256, 256, 0, st>>>(raw[0], vd.nBlkX, vd.nBlkY, p->nBlkXP, p->nBlkYP, mask[0]);
        k_pad_mask<<<(ns + 255) / 256, 256, 0, st>>>(raw[1], vd.nBlkX, vd.nBlkY, p->nBlkXP, p->nBlkYP, mask[1]);
        k_mult_masks<<<(ns + 255) / 256, 256, 0, st>>>(mask[1], mask[0], ns, mask[2]);
        c->launches += 5;
    }
    for (int plane = 0; plane < planes; plane++) {
        const BfpsPlane P = make_bfps_plane(p, time256, plane, mask[0], mask[1], mask[2], c->bfw_dev[plane ? 1 : 0], t[plane ? 1 : 0], super_plane_of(sd, plane));
        dim3 blk(32, 8), grid((P.W + 31) / 32, (P.H + 7) / 8);
        if (sd.bps == 1)
            k_blockfps<uint8_t><<<grid, blk, 0, st>>>(P, vec_b, vec_f, right_super, left_super, dev_dst + cd.offset[plane], cd.pitch[plane], flags);
        else
            k_blockfps<uint16_t><<<grid, blk, 0, st>>>(P, vec_b, vec_f, right_super, left_super, dev_dst + cd.offset[plane], cd.pitch[plane], flags);
        c->launches++;
    }
    MVB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

static int blob_bytes_of(const uint8_t *host_blob) { return *(const int32_t *)host_blob; }

} // namespace mvb200

using namespace mvb200;

extern "C" {

int mvb200_scdetection_batch(mvb200_ctx *ctx, const mvb200_scdetection_params *p, const void *dev_vectors, int64_t vectors_stride, int n, int32_t *scene_change) {
    if (n <= 0)
        return 0;
    Ctx *c = &ctx->c;
    MVB_CUDA_CHECK(cudaSetDevice(c->device));
    int *flags = (int *)c->get_scratch(kSlotSmall, (size_t)n * sizeof(int));
    if (!flags)
        return 1;
    k_flow_scd<<<n, 256, 0, c->stream>>>((const uint8_t *)dev_vectors, vectors_stride, p->vd.nBlkX * p->vd.nBlkY, p->nSCD1, p->nSCD2, 1, flags);
    c->launches++;
    MVB_CUDA_CHECK(cudaGetLastError());
    MVB_CUDA_CHECK(cudaMemcpyAsync(scene_change, flags, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    MVB_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < n; i++)
        scene_change[i] = !scene_change[i]; // the property is "scene change", the flag "usable" (MVSCDetection.c:63)
    return 0;
}

int mvb200_scdetection_frame(mvb200_ctx *ctx, const mvb200_scdetection_params *p, const uint8_t *vectors, int32_t *scene_change) {
    Ctx *c = &ctx->c;
    MVB_CUDA_CHECK(cudaSetDevice(c->device));
    const size_t bytes = (size_t)blob_bytes_of(vectors);
    uint8_t *dev = (uint8_t *)c->get_scratch(kSlotStage, bytes + 256);
    if (!dev)
        return 1;
    MVB_CUDA_CHECK(cudaMemcpyAsync(dev, vectors, bytes, cudaMemcpyHostToDevice, c->stream));
    return mvb200_scdetection_batch(ctx, p, dev, 0, 1, scene_change);
}

int mvb200_mask_batch(mvb200_ctx *ctx, const mvb200_mask_params *p, const void *dev_src_luma, int64_t src_pitch, int64_t src_frame_bytes, const void *dev_vectors, int64_t vectors_stride, int n, void *dev_dst) {
    Ctx *c = &ctx->c;
    MVB_CUDA_CHECK(cudaSetDevice(c->device));
    mvb200_frame_layout ml;
    mvb200_mask_layout(p, &ml);
    for (int f = 0; f < n; f++)
        if (launch_mask(c, p, ml, dev_src_luma ? (const uint8_t *)dev_src_luma + (long long)f * src_frame_bytes : nullptr, src_pitch, (const uint8_t *)dev_vectors + (long long)f * vectors_stride, (uint8_t *)dev_dst + (long long)f * ml.frame_bytes))
            return 1;
    return 0;
}

int mvb200_mask_frame(mvb200_ctx *ctx, const mvb200_mask_params *p, const uint8_t *src_luma, ptrdiff_t src_stride, const uint8_t *vectors, uint8_t *const dst[3], const ptrdiff_t dst_stride[3]) {
    Ctx *c = &ctx->c;
    MVB_CUDA_CHECK(cudaSetDevice(c->device));
    mvb200_frame_layout ml;
    mvb200_mask_layout(p, &ml);
    const size_t blob = (size_t)blob_bytes_of(vectors), luma_pitch = align256((size_t)p->nWidth), luma_bytes = p->kind == 5 ? luma_pitch * p->nHeight : 0;
    uint8_t *buf = (uint8_t *)c->get_scratch(kSlotStage, (size_t)ml.frame_bytes + align256(blob) + luma_bytes + 256);
    if (!buf)
        return 1;
    uint8_t *ddst = buf, *dvec = buf + ml.frame_bytes, *dluma = dvec + align256(blob);
    MVB_CUDA_CHECK(cudaMemcpyAsync(dvec, vectors, blob, cudaMemcpyHostToDevice, c->stream));
    if (p->kind == 5) {
        if (!src_luma)
            return set_error("Mask: kind=5 needs the source clip's luma plane");
        MVB_CUDA_CHECK(cudaMemcpy2DAsync(dluma, luma_pitch, src_luma, (size_t)src_stride, (size_t)p->nWidth, (size_t)p->nHeight, cudaMemcpyHostToDevice, c->stream));
    }
    if (launch_mask(c, p, ml, p->kind == 5 ? dluma : nullptr, (long long)luma_pitch, dvec, ddst))
        return 1;
    for (int pl = 0; pl < 3; pl++)
        MVB_CUDA_CHECK(cudaMemcpy2DAsync(dst[pl], (size_t)dst_stride[pl], ddst + ml.offset[pl], (size_t)ml.pitch[pl], (size_t)ml.plane_w[pl], (size_t)ml.plane_h[pl], cudaMemcpyDeviceToHost, c->stream));
    MVB_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return 0;
}

int mvb200_finest_batch(mvb200_ctx *ctx, const mvb200_super_params *sp, const void *dev_super, int n, void *dev_dst) {
    if (sp->levels > kDevMaxLevels)
        return set_error("mvtools_b200: more than %d pyramid levels are not supported on the device", kDevMaxLevels);
    Ctx *c = &ctx->c;
    MVB_CUDA_CHECK(cudaSetDevice(c->device));
    SuperDesc sd;
    make_super_desc(sp, &sd);
    mvb200_frame_layout fl;
    mvb200_finest_layout(sp, &fl);
    for (int f = 0; f < n; f++)
        if (launch_finest(c, sp, sd, fl, (const uint8_t *)dev_super + (long long)f * sd.frame_bytes, (uint8_t *)dev_dst + (long long)f * fl.frame_bytes))
            return 1;
    return 0;
}

int mvb200_finest_frame(mvb200_ctx *ctx, const mvb200_super_params *sp, const uint8_t *const super[3], const ptrdiff_t super_stride[3], uint8_t *const dst[3], const ptrdiff_t dst_stride[3]) {
    if (sp->levels > kDevMaxLevels)
        return set_error("mvtools_b200: more than %d pyramid levels are not supported on the device", kDevMaxLevels);
    Ctx *c = &ctx->c;
    MVB_CUDA_CHECK(cudaSetDevice(c->device));
    SuperDesc sd;
    make_super_desc(sp, &sd);
    mvb200_frame_layout fl;
    mvb200_finest_layout(sp, &fl);
    uint8_t *buf = (uint8_t *)c->get_scratch(kSlotStage, (size_t)sd.frame_bytes + (size_t)fl.frame_bytes);
    if (!buf)
        return 1;
    uint8_t *dsup = buf, *ddst = buf + sd.frame_bytes;
    const bool chroma = (sp->modeyuv & (MVB200_UPLANE | MVB200_VPLANE)) != 0;
    for (int pl = 0; pl < sd.nplanes; pl++) { // only level 0 (the top pel^2 * padded_h rows of each plane) is read
        if (pl && !chroma)
            continue;
        const mvb200_level_geom &g = sd.lv[pl][0];
        MVB_CUDA_CHECK(cudaMemcpy2DAsync(dsup + sd.offset[pl], (size_t)sd.pitch[pl], super[pl], (size_t)super_stride[pl], (size_t)g.padded_w * sd.bps, (size_t)g.padded_h * sp->pel * sp->pel, cudaMemcpyHostToDevice, c->stream));
    }
    if (launch_finest(c, sp, sd, fl, dsup, ddst))
        return 1;
    for (int pl = 0; pl < sd.nplanes; pl++) {
        if (pl && !chroma)
            continue;
        const mvb200_level_geom &g = sd.lv[pl][0];
        MVB_CUDA_CHECK(cudaMemcpy2DAsync(dst[pl], (size_t)dst_stride[pl], ddst + fl.offset[pl], (size_t)fl.pitch[pl], (size_t)g.padded_w * sp->pel * sd.bps, (size_t)g.padded_h * sp->pel, cudaMemcpyDeviceToHost, c->stream));
    }
    MVB_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return 0;
}

static int flow_check(const mvb200_flow_params *fp, const mvb200_super_params *sp) {
    if (sp->levels > kDevMaxLevels)
        return set_error("mvtools_b200: more than %d pyramid levels are not supported on the device", kDevMaxLevels);
    if (fp->super_pel != sp->pel || fp->super_hpad != sp->hpad || fp->super_vpad != sp->vpad || fp->bits != sp->bits)
        return set_error("Flow: the parameters were derived for another super clip");
    return 0;
}

// dev_usable (may be null): one int per frame, 1 = compensated, 0 = the copy of the source frame
static int flow_batch(mvb200_ctx *ctx, const mvb200_flow_params *fp, const mvb200_super_params *sp, const void *dev_src, const void *dev_super, int first, int n, const void *dev_vectors, int64_t vectors_stride, void *dev_dst, int *dev_usable) {
    if (flow_check(fp, sp))
        return 1;
    if (n <= 0)
        return 0;
    if (first < 0 || first + n > fp->num_frames)
        return set_error("Flow: frames [%d, %d) are outside the clip of %d frames", first, first + n, fp->num_frames);
    if (fp->fields && !fp->tff_exists && fp->vd.nPel > 1)
        return set_error("Flow: _Field property not found in super frame. Therefore, you must pass tff argument.");
    Ctx *c = &ctx->c;
    MVB_CUDA_CHECK(cudaSetDevice(c->device));
    SuperDesc sd;
    ClipDesc cd;
    make_super_desc(sp, &sd);
    make_clip_desc(sp, &cd);
    for (int f = 0; f < n; f++) {
        const int fn = first + f, off = fp->vd.nDeltaFrame;
        int nref = off > 0 ? (fp->vd.isBackward ? fn + off : fn - off) : -off; // MVFlow.cpp:186-193
        if (nref < 0 || nref >= fp->num_frames)
            nref = -1;
        int fs = 0;
        if (nref >= 0 && fp->fields && fp->vd.nPel > 1 && (nref - fn) % 2 != 0) { // :268-299 with an explicit tff
            const int src_top = fp->tff ^ (fn % 2), ref_top = fp->tff ^ (nref % 2);
            fs = (src_top && !ref_top) ? fp->vd.nPel / 2 : ((ref_top && !src_top) ? -(fp->vd.nPel / 2) : 0);
        }
        int *usable;
        if (launch_flow(c, fp, sd, cd, nref >= 0 ? (const uint8_t *)dev_super + (long long)nref * sd.frame_bytes : nullptr, (const uint8_t *)dev_vectors + (long long)f * vectors_stride, fs,
                        dev_src ? (const uint8_t *)dev_src + (long long)fn * cd.frame_bytes : nullptr, (uint8_t *)dev_dst + (long long)f * cd.frame_bytes, &usable))
            return 1;
        if (dev_usable) // (the flag lives in per-launch scratch: keep it before the next frame's launch reuses the slot)
            MVB_CUDA_CHECK(cudaMemcpyAsync(dev_usable + f, usable, sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
    }
    return 0;
}

int mvb200_flow_batch(mvb200_ctx *ctx, const mvb200_flow_params *fp, const mvb200_super_params *sp, const void *dev_src, const void *dev_super, int first, int n, const void *dev_vectors, int64_t vectors_stride, void *dev_dst) {
    return flow_batch(ctx, fp, sp, dev_src, dev_super, first, n, dev_vectors, vectors_stride, dev_dst, nullptr);
}

int mvb200_flow_frame(mvb200_ctx *ctx, const mvb200_flow_params *fp, const mvb200_super_params *sp, const uint8_t *const ref_super[3], const ptrdiff_t ref_stride[3], const uint8_t *vectors, int field_shift, uint8_t *const dst[3], const ptrdiff_t dst_stride[3], int32_t *usable_out) {
    if (flow_check(fp, sp))
        return 1;
    Ctx *c = &ctx->c;
    MVB_CUDA_CHECK(cudaSetDevice(c->device));
    SuperDesc sd;
    ClipDesc cd;
    make_super_desc(sp, &sd);
    make_clip_desc(sp, &cd);
    const size_t blob = (size_t)blob_bytes_of(vectors);
    uint8_t *buf = (uint8_t *)c->get_scratch(kSlotStage, (size_t)sd.frame_bytes + (size_t)cd.frame_bytes + align256(blob) + 256);
    if (!buf)
        return 1;
    uint8_t *dref = buf, *ddst = buf + sd.frame_bytes, *dvec = ddst + cd.frame_bytes;
    MVB_CUDA_CHECK(cudaMemcpyAsync(dvec, vectors, blob, cudaMemcpyHostToDevice, c->stream));
    if (ref_super)
        for (int pl = 0; pl < fp->nplanes; pl++) { // level 0 only
            const mvb200_level_geom &g = sd.lv[pl][0];
            MVB_CUDA_CHECK(cudaMemcpy2DAsync(dref + sd.offset[pl], (size_t)sd.pitch[pl], ref_super[pl], (size_t)ref_stride[pl], (size_t)g.padded_w * sd.bps, (size_t)g.padded_h * sp->pel * sp->pel, cudaMemcpyHostToDevice, c->stream));
        }
    int *usable_dev = nullptr;
    if (launch_flow(c, fp, sd, cd, ref_super ? dref : nullptr, dvec, field_shift, nullptr, ddst, &usable_dev))
        return 1;
    int usable = 0;
    MVB_CUDA_CHECK(cudaMemcpyAsync(&usable, usable_dev, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    MVB_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    *usable_out = usable;
    if (usable) {
        for (int pl = 0; pl < fp->nplanes; pl++)
            MVB_CUDA_CHECK(cudaMemcpy2DAsync(dst[pl], (size_t)dst_stride[pl], ddst + cd.offset[pl], (size_t)cd.pitch[pl], (size_t)cd.w[pl] * cd.bps, (size_t)cd.h[pl], cudaMemcpyDeviceToHost, c->stream));
        MVB_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

static int flowinter_check(const mvb200_flowinter_params *p, const mvb200_super_params *sp) {
    if (sp->levels > kDevMaxLevels)
        return set_error("mvtools_b200: more than %d pyramid levels are not supported on the device", kDevMaxLevels);
    if (p->super_pel != sp->pel || p->super_hpad != sp->hpad || p->super_vpad != sp->vpad || p->bits != sp->bits)
        return set_error("FlowInter: the parameters were derived for another super clip");
    return 0;
}

// dev_flags (may be null): two ints per frame (backward / forward vectors usable); the frame is interpolated when both are
// set, else it is the blend / copy of the source frames
static int flowinter_batch(mvb200_ctx *ctx, const mvb200_flowinter_params *p, const mvb200_super_params *sp, const void *dev_src, const void *dev_super, int first, int n, const void *dev_bw, const void *dev_fw,
                           int64_t vectors_stride, void *dev_dst, int *dev_flags) {
    if (flowinter_check(p, sp))
        return 1;
    if (n <= 0)
        return 0;
    if (first < 0 || first + n > p->num_frames)
        return set_error("FlowInter: frames [%d, %d) are outside the clip of %d frames", first, first + n, p->num_frames);
    if (!dev_src)
        return set_error("FlowInter: the batch entry point needs the source clip (unusable frames are made from it)");
    Ctx *c = &ctx->c;
    MVB_CUDA_CHECK(cudaSetDevice(c->device));
    SuperDesc sd;
    ClipDesc cd;
    make_super_desc(sp, &sd);
    make_clip_desc(sp, &cd);
    const uint8_t *sup = (const uint8_t *)dev_super, *bw = (const uint8_t *)dev_bw, *fw = (const uint8_t *)dev_fw, *src = (const uint8_t *)dev_src;
    for (int f = 0; f < n; f++) {
        const int fn = first + f, off = p->vd.nDeltaFrame, fr = fn + off;
        const bool in_range = fr < p->num_frames; // MVFlowInter.c:115
        const int fref = fr < p->num_frames - 1 ? fr : p->num_frames - 1; // VSMIN(n + off, numFrames - 1), :421
        int *flags;
        if (launch_flowinter(c, p, sd, cd, p->time256, sup + (long long)fn * sd.frame_bytes, in_range ? sup + (long long)fr * sd.frame_bytes : nullptr, in_range ? bw + (long long)fn * vectors_stride : nullptr,
                             in_range ? fw + (long long)fr * vectors_stride : nullptr, in_range ? bw + (long long)fr * vectors_stride : nullptr, in_range ? fw + (long long)fn * vectors_stride : nullptr,
                             src + (long long)fn * cd.frame_bytes, src + (long long)fref * cd.frame_bytes, (uint8_t *)dev_dst + (long long)f * cd.frame_bytes, &flags))
            return 1;
        if (dev_flags) // (per-launch scratch: keep the pair before the next frame's launch reuses the slot)
            MVB_CUDA_CHECK(cudaMemcpyAsync(dev_flags + 2 * f, flags, 2 * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
    }
    return 0;
}

int mvb200_flowinter_batch(mvb200_ctx *ctx, const mvb200_flowinter_params *p, const mvb200_super_params *sp, const void *dev_src, const void *dev_super, int first, int n, const void *dev_bw, const void *dev_fw,
                           int64_t vectors_stride, void *dev_dst) {
    return flowinter_batch(ctx, p, sp, dev_src, dev_super, first, n, dev_bw, dev_fw, vectors_stride, dev_dst, nullptr);
}

int mvb200_flowinter_frame(mvb200_ctx *ctx, const mvb200_flowinter_params *p, const mvb200_super_params *sp, int time256, const uint8_t *const left_super[3], const ptrdiff_t left_stride[3],
                           const uint8_t *const right_super[3], const ptrdiff_t right_stride[3], const uint8_t *vec_b, const uint8_t *vec_f, const uint8_t *vec_bb, const uint8_t *vec_ff, uint8_t *const dst[3],
                           const ptrdiff_t dst_stride[3], int32_t *usable_out) {
    if (flowinter_check(p, sp))
        return 1;
    *usable_out = 0;
    if (!left_super || !right_super || !vec_b || !vec_f)
        return 0;
    if (time256 < 0 || time256 > 256)
        return set_error("FlowInter: time256 = %d is outside [0, 256]", time256);
    Ctx *c = &ctx->c;
    MVB_CUDA_CHECK(cudaSetDevice(c->device));
    SuperDesc sd;
    ClipDesc cd;
    make_super_desc(sp, &sd);
    make_clip_desc(sp, &cd);
    const uint8_t *host_blobs[4] = {vec_b, vec_f, vec_bb, vec_ff};
    size_t blob_off[4], total = 0;
    for (int i = 0; i < 4; i++) {
        blob_off[i] = total;
        total += host_blobs[i] ? align256((size_t)blob_bytes_of(host_blobs[i])) : 0;
    }
    uint8_t *buf = (uint8_t *)c->get_scratch(kSlotStage, 2 * (size_t)sd.frame_bytes + (size_t)cd.frame_bytes + total + 256);
    if (!buf)
        return 1;
    uint8_t *dleft = buf, *dright = buf + sd.frame_bytes, *ddst = dright + sd.frame_bytes, *dvec = ddst + cd.frame_bytes;
    const uint8_t *dev_blobs[4];
    for (int i = 0; i < 4; i++) {
        dev_blobs[i] = host_blobs[i] ? dvec + blob_off[i] : nullptr;
        if (host_blobs[i])
            MVB_CUDA_CHECK(cudaMemcpyAsync(dvec + blob_off[i], host_blobs[i], (size_t)blob_bytes_of(host_blobs[i]), cudaMemcpyHostToDevice, c->stream));
    }
    for (int pl = 0; pl < p->nplanes; pl++) { // level 0 only
        const mvb200_level_geom &g = sd.lv[pl][0];
        MVB_CUDA_CHECK(cudaMemcpy2DAsync(dleft + sd.offset[pl], (size_t)sd.pitch[pl], left_super[pl], (size_t)left_stride[pl], (size_t)g.padded_w * sd.bps, (size_t)g.padded_h * sp->pel * sp->pel, cudaMemcpyHostToDevice, c->stream));
        MVB_CUDA_CHECK(cudaMemcpy2DAsync(dright + sd.offset[pl], (size_t)sd.pitch[pl], right_super[pl], (size_t)right_stride[pl], (size_t)g.padded_w * sd.bps, (size_t)g.padded_h * sp->pel * sp->pel, cudaMemcpyHostToDevice, c->stream));
    }
    int *flags_dev = nullptr;
    if (launch_flowinter(c, p, sd, cd, time256, dleft, dright, dev_blobs[0], dev_blobs[1], dev_blobs[2], dev_blobs[3], nullptr, nullptr, ddst, &flags_dev))
        return 1;
    int flags[4] = {0, 0, 0, 0};
    MVB_CUDA_CHECK(cudaMemcpyAsync(flags, flags_dev, sizeof(flags), cudaMemcpyDeviceToHost, c->stream));
    MVB_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    *usable_out = flags[0] && flags[1];
    if (*usable_out) {
        for (int pl = 0; pl < p->nplanes; pl++)
            MVB_CUDA_CHECK(cudaMemcpy2DAsync(dst[pl], (size_t)dst_stride[pl], ddst + cd.offset[pl], (size_t)cd.pitch[pl], (size_t)cd.w[pl] * cd.bps, (size_t)cd.h[pl], cudaMemcpyDeviceToHost, c->stream));
        MVB_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

int mvb200_blend_frame(mvb200_ctx *ctx, const mvb200_super_params *sp, int time256, const uint8_t *const src[3], const ptrdiff_t src_stride[3], const uint8_t *const ref[3], const ptrdiff_t ref_stride[3], uint8_t *const dst[3],
                       const ptrdiff_t dst_stride[3]) {
    Ctx *c = &ctx->c;
    MVB_CUDA_CHECK(cudaSetDevice(c->device));
    ClipDesc cd;
    make_clip_desc(sp, &cd);
    uint8_t *buf = (uint8_t *)c->get_scratch(kSlotStage, 3 * (size_t)cd.frame_bytes);
    if (!buf)
        return 1;
    uint8_t *dsrc = buf, *dref = buf + cd.frame_bytes, *ddst = dref + cd.frame_bytes;
    for (int pl = 0; pl < cd.nplanes; pl++) {
        MVB_CUDA_CHECK(cudaMemcpy2DAsync(dsrc + cd.offset[pl], (size_t)cd.pitch[pl], src[pl], (size_t)src_stride[pl], (size_t)cd.w[pl] * cd.bps, (size_t)cd.h[pl], cudaMemcpyHostToDevice, c->stream));
        MVB_CUDA_CHECK(cudaMemcpy2DAsync(dref + cd.offset[pl], (size_t)cd.pitch[pl], ref[pl], (size_t)ref_stride[pl], (size_t)cd.w[pl] * cd.bps, (size_t)cd.h[pl], cudaMemcpyHostToDevice, c->stream));
        dim3 blk(32, 8), grid((cd.w[pl] + 31) / 32, (cd.h[pl] + 7) / 8);
        if (cd.bps == 1)
            k_blend<uint8_t><<<grid, blk, 0, c->stream>>>(cd.w[pl], cd.h[pl], time256, dsrc + cd.offset[pl], dref + cd.offset[pl], ddst + cd.offset[pl], cd.pitch[pl]);
        else
            k_blend<uint16_t><<<grid, blk, 0, c->stream>>>(cd.w[pl], cd.h[pl], time256, dsrc + cd.offset[pl], dref + cd.offset[pl], ddst + cd.offset[pl], cd.pitch[pl]);
        c->launches++;
        MVB_CUDA_CHECK(cudaMemcpy2DAsync(dst[pl], (size_t)dst_stride[pl], ddst + cd.offset[pl], (size_t)cd.pitch[pl], (size_t)cd.w[pl] * cd.bps, (size_t)cd.h[pl], cudaMemcpyDeviceToHost, c->stream));
    }
    MVB_CUDA_CHECK(cudaGetLastError());
    MVB_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return 0;
}

static int flowblur_check(const mvb200_flowblur_params *p, const mvb200_super_params *sp) {
    if (sp->levels > kDevMaxLevels)
        return set_error("mvtools_b200: more than %d pyramid levels are not supported on the device", kDevMaxLevels);
    if (p->super_pel != sp->pel || p->super_hpad != sp->hpad || p->super_vpad != sp->vpad || p->bits != sp->bits)
        return set_error("FlowBlur: the parameters were derived for another super clip");
    return 0;
}

int mvb200_flowblur_batch(mvb200_ctx *ctx, const mvb200_flowblur_params *p, const mvb200_super_params *sp, const void *dev_src, const void *dev_super, int first, int n, const void *dev_bw, const void *dev_fw, int64_t vectors_stride,
                          void *dev_dst) {
    if (flowblur_check(p, sp))
        return 1;
    if (n <= 0)
        return 0;
    if (first < 0 || first + n > p->num_frames)
        return set_error("FlowBlur: frames [%d, %d) are outside the clip of %d frames", first, first + n, p->num_frames);
    if (!dev_src)
        return set_error("FlowBlur: the batch entry point needs the source clip (unusable frames are copies of it)");
    Ctx *c = &ctx->c;
    MVB_CUDA_CHECK(cudaSetDevice(c->device));
    SuperDesc sd;
    ClipDesc cd;
    make_super_desc(sp, &sd);
    make_clip_desc(sp, &cd);
    for (int f = 0; f < n; f++) {
        const int fn = first + f, off = p->vd.nDeltaFrame;
        const bool in_range = fn - off >= 0 && fn + off < p->num_frames; // MVFlowBlur.c:173
        int *flags;
        if (launch_flowblur(c, p, sd, cd, (const uint8_t *)dev_super + (long long)fn * sd.frame_bytes, in_range ? (const uint8_t *)dev_bw + (long long)(fn - off) * vectors_stride : nullptr,
                            in_range ? (const uint8_t *)dev_fw + (long long)(fn + off) * vectors_stride : nullptr, (const uint8_t *)dev_src + (long long)fn * cd.frame_bytes, (uint8_t *)dev_dst + (long long)f * cd.frame_bytes, &flags))
            return 1;
    }
    return 0;
}

int mvb200_flowblur_frame(mvb200_ctx *ctx, const mvb200_flowblur_params *p, const mvb200_super_params *sp, const uint8_t *const super[3], const ptrdiff_t super_stride[3], const uint8_t *vec_b, const uint8_t *vec_f, uint8_t *const dst[3],
                          const ptrdiff_t dst_stride[3], int32_t *usable_out) {
    if (flowblur_check(p, sp))
        return 1;
    *usable_out = 0;
    if (!super || !vec_b || !vec_f)
        return 0;
    Ctx *c = &ctx->c;
    MVB_CUDA_CHECK(cudaSetDevice(c->device));
    SuperDesc sd;
    ClipDesc cd;
    make_super_desc(sp, &sd);
    make_clip_desc(sp, &cd);
    const size_t bb = (size_t)blob_bytes_of(vec_b), bf = (size_t)blob_bytes_of(vec_f);
    uint8_t *buf = (uint8_t *)c->get_scratch(kSlotStage, (size_t)sd.frame_bytes + (size_t)cd.frame_bytes + align256(bb) + align256(bf) + 256);
    if (!buf)
        return 1;
    uint8_t *dsup = buf, *ddst = buf + sd.frame_bytes, *dvb = ddst + cd.frame_bytes, *dvf = dvb + align256(bb);
    MVB_CUDA_CHECK(cudaMemcpyAsync(dvb, vec_b, bb, cudaMemcpyHostToDevice, c->stream));
    MVB_CUDA_CHECK(cudaMemcpyAsync(dvf, vec_f, bf, cudaMemcpyHostToDevice, c->stream));
    for (int pl = 0; pl < p->nplanes; pl++) { // level 0 only
        const mvb200_level_geom &g = sd.lv[pl][0];
        MVB_CUDA_CHECK(cudaMemcpy2DAsync(dsup + sd.offset[pl], (size_t)sd.pitch[pl], super[pl], (size_t)super_stride[pl], (size_t)g.padded_w * sd.bps, (size_t)g.padded_h * sp->pel * sp->pel, cudaMemcpyHostToDevice, c->stream));
    }
    int *flags_dev = nullptr;
    if (launch_flowblur(c, p, sd, cd, dsup, dvb, dvf, nullptr, ddst, &flags_dev))
        return 1;
    int flags[2] = {0, 0};
    MVB_CUDA_CHECK(cudaMemcpyAsync(flags, flags_dev, sizeof(flags), cudaMemcpyDeviceToHost, c->stream));
    MVB_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    *usable_out = flags[0] && flags[1];
    if (*usable_out) {
        for (int pl = 0; pl < p->nplanes; pl++)
            MVB_CUDA_CHECK(cudaMemcpy2DAsync(dst[pl], (size_t)dst_stride[pl], ddst + cd.offset[pl], (size_t)cd.pitch[pl], (size_t)cd.w[pl] * cd.bps, (size_t)cd.h[pl], cudaMemcpyDeviceToHost, c->stream));
        MVB_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

int mvb200_blockfps_frame(mvb200_ctx *ctx, const mvb200_blockfps_params *p, const mvb200_super_params *sp, int time256, const uint8_t *const left_super[3], const ptrdiff_t left_stride[3], const uint8_t *const right_super[3],
                          const ptrdiff_t right_stride[3], const uint8_t *vec_b, const uint8_t *vec_f, uint8_t *const dst[3], const ptrdiff_t dst_stride[3], int32_t *usable_out) {
    if (sp->levels > kDevMaxLevels)
        return set_error("mvtools_b200: more than %d pyramid levels are not supported on the device", kDevMaxLevels);
    if (p->super_pel != sp->pel || p->super_hpad != sp->hpad || p->super_vpad != sp->vpad || p->bits != sp->bits)
        return set_error("BlockFPS: the parameters were derived for another super clip");
    *usable_out = 0;
    if (!left_super || !right_super || !vec_b || !vec_f)
        return 0;
    if (time256 < 0 || time256 > 256)
        return set_error("BlockFPS: time256 = %d is outside [0, 256]", time256);
    Ctx *c = &ctx->c;
    MVB_CUDA_CHECK(cudaSetDevice(c->device));
    SuperDesc sd;
    ClipDesc cd;
    make_super_desc(sp, &sd);
    make_clip_desc(sp, &cd);
    const int planes = (p->super_modeyuv & (MVB200_UPLANE | MVB200_VPLANE)) ? 3 : 1;
    const size_t bb = (size_t)blob_bytes_of(vec_b), bf = (size_t)blob_bytes_of(vec_f);
    uint8_t *buf = (uint8_t *)c->get_scratch(kSlotStage, 2 * (size_t)sd.frame_bytes + (size_t)cd.frame_bytes + align256(bb) + align256(bf) + 256);
    if (!buf)
        return 1;
    uint8_t *dleft = buf, *dright = buf + sd.frame_bytes, *ddst = dright + sd.frame_bytes, *dvb = ddst + cd.frame_bytes, *dvf = dvb + align256(bb);
    MVB_CUDA_CHECK(cudaMemcpyAsync(dvb, vec_b, bb, cudaMemcpyHostToDevice, c->stream));
    MVB_CUDA_CHECK(cudaMemcpyAsync(dvf, vec_f, bf, cudaMemcpyHostToDevice, c->stream));
    for (int pl = 0; pl < planes; pl++) { // level 0 only
        const mvb200_level_geom &g = sd.lv[pl][0];
        MVB_CUDA_CHECK(cudaMemcpy2DAsync(dleft + sd.offset[pl], (size_t)sd.pitch[pl], left_super[pl], (size_t)left_stride[pl], (size_t)g.padded_w * sd.bps, (size_t)g.padded_h * sp->pel * sp->pel, cudaMemcpyHostToDevice, c->stream));
        MVB_CUDA_CHECK(cudaMemcpy2DAsync(dright + sd.offset[pl], (size_t)sd.pitch[pl], right_super[pl], (size_t)right_stride[pl], (size_t)g.padded_w * sd.bps, (size_t)g.padded_h * sp->pel * sp->pel, cudaMemcpyHostToDevice, c->stream));
    }
    int *flags_dev = nullptr;
    if (launch_blockfps(c, p, sd, cd, time256, dleft, dright, dvb, dvf, ddst, &flags_dev))
        return 1;
    int flags[2] = {0, 0};
    MVB_CUDA_CHECK(cudaMemcpyAsync(flags, flags_dev, sizeof(flags), cudaMemcpyDeviceToHost, c->stream));
    MVB_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    *usable_out = flags[0] && flags[1];
    if (*usable_out) {
        for (int pl = 0; pl < planes; pl++)
            MVB_CUDA_CHECK(cudaMemcpy2DAsync(dst[pl], (size_t)dst_stride[pl], ddst + cd.offset[pl], (size_t)cd.pitch[pl], (size_t)cd.w[pl] * cd.bps, (size_t)cd.h[pl], cudaMemcpyDeviceToHost, c->stream));
        MVB_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

int mvb200_flowfps_clip_host(mvb200_ctx *ctx, const mvb200_super_params *sp, const mvb200_analyse_params *ap_bw, const mvb200_analyse_params *ap_fw, const mvb200_flowfps_params *p, const void *host_src, int src_first, int n,
                             int out_first, int out_count, void *host_dst, int32_t *usable_out) {
    const mvb200_flowinter_params *q = &p->inter;
    if (flowinter_check(q, sp))
        return 1;
    if (n <= 0 || out_count <= 0)
        return 0;
    if (src_first < 0 || src_first + n > p->num_frames_in || out_first < 0 || out_first + out_count > p->num_frames_out)
        return set_error("FlowFPS: the window is outside the clip");
    Ctx *c = &ctx->c;
    MVB_CUDA_CHECK(cudaSetDevice(c->device));
    SuperDesc sd;
    ClipDesc cd;
    make_super_desc(sp, &sd);
    make_clip_desc(sp, &cd);
    // the segment is the clip as far as the searches are concerned (references outside it: default vectors, only met by halo frames)
    mvb200_analyse_params aps[2] = {*ap_bw, *ap_fw};
    aps[0].num_frames = aps[1].num_frames = n;
    const mvb200_analyse_params *app[2] = {&aps[0], &aps[1]};
    const int64_t vstride = ((int64_t)(aps[0].vectors_size > aps[1].vectors_size ? aps[0].vectors_size : aps[1].vectors_size) + 255) / 256 * 256;
    uint8_t *buf = (uint8_t *)c->get_scratch(kSlotStage, (size_t)cd.frame_bytes * (n + out_count) + (size_t)sd.frame_bytes * n + (size_t)vstride * n * 2 + align256(2 * sizeof(int) * (size_t)out_count) + 256);
    if (!buf)
        return 1;
    uint8_t *dsrc = buf, *ddst = dsrc + (size_t)cd.frame_bytes * n, *dsup = ddst + (size_t)cd.frame_bytes * out_count, *dbw = dsup + (size_t)sd.frame_bytes * n, *dfw = dbw + (size_t)vstride * n;
    int *dflags = (int *)(dfw + (size_t)vstride * n); // two per output frame: both set = interpolated from the vectors
    if (usable_out)
        MVB_CUDA_CHECK(cudaMemsetAsync(dflags, 0, 2 * sizeof(int) * (size_t)out_count, c->stream));
    MVB_CUDA_CHECK(cudaMemcpyAsync(dsrc, host_src, (size_t)cd.frame_bytes * n, cudaMemcpyHostToDevice, c->stream));
    if (mvb200_super_batch(ctx, sp, dsrc, dsup, n))
        return 1;
    void *outs[2] = {dbw, dfw};
    if (mvb200_analyse_multi_batch(ctx, app, 2, sp, dsup, 0, n, outs, vstride))
        return 1;
    const int off = q->vd.nDeltaFrame, N = p->num_frames_in;
    for (int k = 0; k < out_count; k++) {
        int32_t nleft, time256;
        mvb200_flowfps_position(p, out_first + k, &nleft, &time256);
        const int nright = nleft + off;
        uint8_t *dst = ddst + (size_t)cd.frame_bytes * k;
        auto local = [&](int f) { return f - src_first; };                 // position of source frame f in the segment
        auto have = [&](int f) { return f >= src_first && f < src_first + n; };
        const int fl = nleft < N - 1 ? nleft : N - 1, fr = nright < N - 1 ? nright : N - 1; // VSMIN(.., numFrames - 1), MVFlowFPS.c:137-141,474-486
        if (time256 == 0 || time256 == 256) { // the input frame itself
            const int f = time256 == 0 ? fl : fr;
            if (!have(f))
                return set_error("FlowFPS: source frame %d is not in the segment", f);
            MVB_CUDA_CHECK(cudaMemcpyAsync(dst, dsrc + (size_t)cd.frame_bytes * local(f), (size_t)cd.frame_bytes, cudaMemcpyDeviceToDevice, c->stream));
            continue;
        }
        if (!have(fl) || !have(fr))
            return set_error("FlowFPS: source frames %d / %d are not in the segment", fl, fr);
        const bool in_range = nleft < N && nright < N; // :156
        const bool extra = p->maskmode == 2 && in_range;
        if (in_range && !have(nright))
            return set_error("FlowFPS: source frame %d is not in the segment", nright);
        // the extra pair -- mvbw of the right frame, mvfw of the left one -- was searched against frames nright + off and
        // nleft - off: inside the clip they must be inside the segment too, or the segment's searches gave default vectors
        // where the whole clip's would not
        if (extra && ((nright + off < N && !have(nright + off)) || (nleft - off >= 0 && !have(nleft - off))))
            return set_error("FlowFPS: the segment [%d, %d) does not hold frames %d / %d the extra vectors of frames %d / %d refer to", src_first, src_first + n, nleft - off, nright + off, nleft, nright);
        int *flags;
        if (launch_flowinter(c, q, sd, cd, time256, in_range ? dsup + (size_t)sd.frame_bytes * local(nleft) : nullptr, in_range ? dsup + (size_t)sd.frame_bytes * local(nright) : nullptr,
                             in_range ? dbw + (size_t)vstride * local(nleft) : nullptr, in_range ? dfw + (size_t)vstride * local(nright) : nullptr, extra ? dbw + (size_t)vstride * local(nright) : nullptr,
                             extra ? dfw + (size_t)vstride * local(nleft) : nullptr, dsrc + (size_t)cd.frame_bytes * local(fl), dsrc + (size_t)cd.frame_bytes * local(fr), dst, &flags))
            return 1;
        if (usable_out) // (per-launch scratch: keep the pair before the next frame's launch reuses the slot)
            MVB_CUDA_CHECK(cudaMemcpyAsync(dflags + 2 * k, flags, 2 * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
    }
    std::vector<int> pairs(usable_out ? 2 * (size_t)out_count : 0);
    if (usable_out)
        MVB_CUDA_CHECK(cudaMemcpyAsync(pairs.data(), dflags, sizeof(int) * pairs.size(), cudaMemcpyDeviceToHost, c->stream));
    MVB_CUDA_CHECK(cudaMemcpyAsync(host_dst, ddst, (size_t)cd.frame_bytes * out_count, cudaMemcpyDeviceToHost, c->stream));
    MVB_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    for (int k = 0; usable_out && k < out_count; k++)
        usable_out[k] = pairs[2 * k] && pairs[2 * k + 1];
    return 0;
}

// Common front of the *_clip_host entry points: the segment's source frames to the device, their super frames, and the
// blobs of nap (1 or 2) mv.Analyse instances for EVERY frame of the segment (num_frames = n: the segment is the clip).
struct ChainBuffers {
    uint8_t *dsrc, *ddst, *dsup, *dvec[2];
    int *dflags; // room for two per output frame
    int64_t vstride;
    SuperDesc sd;
    ClipDesc cd;
};
// (out_buffers: output frames to make room for when they are not frames [out_first, out_first + out_count) of the source clip -- a rate conversion)
// reach_back / reach_fwd: how far before / after an output frame the vectors it is made of refer (the frames they were searched
// against).  The segment's searches see the segment as the whole clip, so inside the real clip those frames must be inside
// the segment -- else a vector clip silently holds default vectors where the whole clip's search would not.
static int chain_front(mvb200_ctx *ctx, const mvb200_super_params *sp, const mvb200_analyse_params *const *ap, int nap, const void *host_src, int src_first, int n, int out_first, int out_count, const char *filter, ChainBuffers *B,
                       int reach_back, int reach_fwd, int out_buffers = -1) {
    if (sp->levels > kDevMaxLevels)
        return set_error("mvtools_b200: more than %d pyramid levels are not supported on the device", kDevMaxLevels);
    if (out_first < src_first || out_first + out_count > src_first + n)
        return set_error("%s: output frames [%d, %d) are outside the segment [%d, %d)", filter, out_first, out_first + out_count, src_first, src_first + n);
    if (out_count > 0) {
        const int realN = ap[0]->num_frames;
        const int need0 = out_first - reach_back > 0 ? out_first - reach_back : 0, need1 = out_first + out_count + reach_fwd < realN ? out_first + out_count + reach_fwd : realN;
        if (need0 < src_first || need1 > src_first + n)
            return set_error("%s: the segment [%d, %d) does not hold the frames [%d, %d) that the vectors of output frames [%d, %d) refer to", filter, src_first, src_first + n, need0, need1, out_first, out_first + out_count);
    }
    mvb200_analyse_params aps[2];
    const mvb200_analyse_params *app[2];
    B->vstride = 0;
    for (int i = 0; i < nap; i++) {
        if (ap[i]->ad.nDeltaFrame <= 0 || ap[i]->fields)
            return set_error("%s: the clip pipeline handles relative references without field shifts only", filter);
        aps[i] = *ap[i];
        aps[i].num_frames = n;
        app[i] = &aps[i];
        const int64_t v = ((int64_t)aps[i].vectors_size + 255) / 256 * 256;
        B->vstride = v > B->vstride ? v : B->vstride;
    }
    if (out_buffers >= 0)
        out_count = out_buffers;
    Ctx *c = &ctx->c;
    MVB_CUDA_CHECK(cudaSetDevice(c->device));
    make_super_desc(sp, &B->sd);
    make_clip_desc(sp, &B->cd);
    uint8_t *buf = (uint8_t *)c->get_scratch(kSlotStage, (size_t)B->cd.frame_bytes * (n + out_count) + (size_t)B->sd.frame_bytes * n + (size_t)B->vstride * n * nap + align256(2 * sizeof(int) * (size_t)out_count) + 256);
    if (!buf)
        return 1;
    B->dsrc = buf;
    B->ddst = B->dsrc + (size_t)B->cd.frame_bytes * n;
    B->dsup = B->ddst + (size_t)B->cd.frame_bytes * out_count;
    B->dvec[0] = B->dsup + (size_t)B->sd.frame_bytes * n;
    B->dvec[1] = B->dvec[0] + (size_t)B->vstride * n;
    B->dflags = (int *)(B->dvec[0] + (size_t)B->vstride * n * nap);
    MVB_CUDA_CHECK(cudaMemcpyAsync(B->dsrc, host_src, (size_t)B->cd.frame_bytes * n, cudaMemcpyHostToDevice, c->stream));
    void *outs[2] = {B->dvec[0], B->dvec[1]};
    return mvb200_super_batch(ctx, sp, B->dsrc, B->dsup, n) || mvb200_analyse_multi_batch(ctx, app, nap, sp, B->dsup, 0, n, outs, B->vstride);
}
static int chain_back(mvb200_ctx *ctx, const ChainBuffers &B, int out_count, void *host_dst) {
    MVB_CUDA_CHECK(cudaMemcpyAsync(host_dst, B.ddst, (size_t)B.cd.frame_bytes * out_count, cudaMemcpyDeviceToHost, ctx->c.stream));
    MVB_CUDA_CHECK(cudaStreamSynchronize(ctx->c.stream));
    return 0;
}

int mvb200_flow_clip_host(mvb200_ctx *ctx, const mvb200_super_params *sp, const mvb200_analyse_params *ap, const mvb200_flow_params *fp, const void *host_src, int src_first, int n, int out_first, int out_count, void *host_dst,
                          int32_t *usable_out) {
    if (n <= 0 || out_count <= 0)
        return 0;
    if (fp->fields)
        return set_error("Flow: the clip pipeline handles relative references without field shifts only");
    ChainBuffers B;
    const int d = ap->ad.nDeltaFrame;
    if (chain_front(ctx, sp, &ap, 1, host_src, src_first, n, out_first, out_count, "Flow", &B, ap->ad.isBackward ? 0 : d, ap->ad.isBackward ? d : 0))
        return 1;
    mvb200_flow_params q = *fp;
    q.num_frames = n;
    const int first = out_first - src_first;
    if (flow_batch(ctx, &q, sp, B.dsrc, B.dsup, first, out_count, B.dvec[0] + (size_t)B.vstride * first, B.vstride, B.ddst, usable_out ? B.dflags : nullptr))
        return 1;
    static_assert(sizeof(int) == sizeof(int32_t), "flags are copied out as they are");
    if (usable_out)
        MVB_CUDA_CHECK(cudaMemcpyAsync(usable_out, B.dflags, sizeof(int) * (size_t)out_count, cudaMemcpyDeviceToHost, ctx->c.stream));
    return chain_back(ctx, B, out_count, host_dst);
}

int mvb200_flowinter_clip_host(mvb200_ctx *ctx, const mvb200_super_params *sp, const mvb200_analyse_params *ap_bw, const mvb200_analyse_params *ap_fw, const mvb200_flowinter_params *p, const void *host_src, int src_first, int n,
                               int out_first, int out_count, void *host_dst, int32_t *usable_out) {
    if (n <= 0 || out_count <= 0)
        return 0;
    ChainBuffers B;
    const mvb200_analyse_params *ap[2] = {ap_bw, ap_fw};
    // frames n and n + off; their extra vectors (mvfw of n, mvbw of n + off) were searched against n - off and n + 2 off
    if (chain_front(ctx, sp, ap, 2, host_src, src_first, n, out_first, out_count, "FlowInter", &B, p->vd.nDeltaFrame, 2 * p->vd.nDeltaFrame))
        return 1;
    mvb200_flowinter_params q = *p;
    q.num_frames = n;
    if (flowinter_batch(ctx, &q, sp, B.dsrc, B.dsup, out_first - src_first, out_count, B.dvec[0], B.dvec[1], B.vstride, B.ddst, usable_out ? B.dflags : nullptr))
        return 1;
    std::vector<int> pairs(usable_out ? 2 * (size_t)out_count : 0);
    if (usable_out)
        MVB_CUDA_CHECK(cudaMemcpyAsync(pairs.data(), B.dflags, sizeof(int) * pairs.size(), cudaMemcpyDeviceToHost, ctx->c.stream));
    if (chain_back(ctx, B, out_count, host_dst)) // (synchronises the stream)
        return 1;
    for (int k = 0; usable_out && k < out_count; k++)
        usable_out[k] = pairs[2 * k] && pairs[2 * k + 1];
    return 0;
}

int mvb200_flowblur_clip_host(mvb200_ctx *ctx, const mvb200_super_params *sp, const mvb200_analyse_params *ap_bw, const mvb200_analyse_params *ap_fw, const mvb200_flowblur_params *p, const void *host_src, int src_first, int n,
                              int out_first, int out_count, void *host_dst) {
    if (n <= 0 || out_count <= 0)
        return 0;
    ChainBuffers B;
    const mvb200_analyse_params *ap[2] = {ap_bw, ap_fw};
    if (chain_front(ctx, sp, ap, 2, host_src, src_first, n, out_first, out_count, "FlowBlur", &B, p->vd.nDeltaFrame, p->vd.nDeltaFrame))
        return 1;
    mvb200_flowblur_params q = *p;
    q.num_frames = n;
    if (mvb200_flowblur_batch(ctx, &q, sp, B.dsrc, B.dsup, out_first - src_first, out_count, B.dvec[0], B.dvec[1], B.vstride, B.ddst))
        return 1;
    return chain_back(ctx, B, out_count, host_dst);
}

int mvb200_blockfps_clip_host(mvb200_ctx *ctx, const mvb200_super_params *sp, const mvb200_analyse_params *ap_bw, const mvb200_analyse_params *ap_fw, const mvb200_blockfps_params *p, const void *host_src, int src_first, int n,
                              int out_first, int out_count, void *host_dst, int32_t *usable_out) {
    if (p->super_pel != sp->pel || p->super_hpad != sp->hpad || p->super_vpad != sp->vpad || p->bits != sp->bits)
        return set_error("BlockFPS: the parameters were derived for another super clip");
    if (n <= 0 || out_count <= 0)
        return 0;
    if (src_first < 0 || src_first + n > p->num_frames_in || out_first < 0 || out_first + out_count > p->num_frames_out)
        return set_error("BlockFPS: the window is outside the clip");
    ChainBuffers B;
    const mvb200_analyse_params *ap[2] = {ap_bw, ap_fw};
    // (chain_front's own window test compares output with source frame numbers; they are different clips here)
    if (chain_front(ctx, sp, ap, 2, host_src, src_first, n, src_first, 0, "BlockFPS", &B, 0, 0, out_count))
        return 1;
    Ctx *c = &ctx->c;
    const ClipDesc &cd = B.cd;
    const int off = p->vd.nDeltaFrame, N = p->num_frames_in;
    const int planes = (p->super_modeyuv & (MVB200_UPLANE | MVB200_VPLANE)) ? 3 : 1; // the planes the reference writes (:652)
    auto local = [&](int f) { return f - src_first; }; // position of source frame f in the segment
    auto have = [&](int f) { return f >= src_first && f < src_first + n; };
    for (int k = 0; k < out_count; k++) {
        int32_t nleft, time256;
        mvb200_blockfps_position(p, out_first + k, &nleft, &time256);
        const int nright = nleft + off;
        uint8_t *dst = B.ddst + (size_t)cd.frame_bytes * k;
        const int fl = nleft < N - 1 ? nleft : N - 1, fr = nright < N - 1 ? nright : N - 1; // VSMIN(.., numFrames - 1), MVBlockFPS.c:246-272
        if (time256 == 0 || time256 == 256) { // the input frame itself
            const int f = time256 == 0 ? fl : fr;
            if (!have(f))
                return set_error("BlockFPS: source frame %d is not in the segment", f);
            MVB_CUDA_CHECK(cudaMemcpyAsync(dst, B.dsrc + (size_t)cd.frame_bytes * local(f), (size_t)cd.frame_bytes, cudaMemcpyDeviceToDevice, c->stream));
            MVB_CUDA_CHECK(cudaMemsetAsync(B.dflags + 2 * k, 0, 2 * sizeof(int), c->stream));
            continue;
        }
        if (!have(fl) || !have(fr))
            return set_error("BlockFPS: source frames %d / %d are not in the segment", fl, fr);
        const uint8_t *src = B.dsrc + (size_t)cd.frame_bytes * local(fl), *ref = B.dsrc + (size_t)cd.frame_bytes * local(fr);
        const bool in_range = nleft < N && nright < N; // :275
        int *flags = nullptr;
        if (in_range) {
            if (launch_blockfps(c, p, B.sd, cd, time256, B.dsup + (size_t)B.sd.frame_bytes * local(nleft), B.dsup + (size_t)B.sd.frame_bytes * local(nright), B.dvec[0] + (size_t)B.vstride * local(nleft),
                                B.dvec[1] + (size_t)B.vstride * local(nright), dst, &flags))
                return 1;
            MVB_CUDA_CHECK(cudaMemcpyAsync(B.dflags + 2 * k, flags, 2 * sizeof(int), cudaMemcpyDeviceToDevice, c->stream)); // (per-launch scratch)
        } else {
            MVB_CUDA_CHECK(cudaMemsetAsync(B.dflags + 2 * k, 0, 2 * sizeof(int), c->stream));
        }
        // frames the vectors cannot make (scene change, clip end): the blend of the source frames, or the left one
        if (planes < cd.nplanes) // (planes the reference leaves uninitialised: the left frame's here)
            MVB_CUDA_CHECK(cudaMemcpyAsync(dst + cd.offset[planes], src + cd.offset[planes], (size_t)(cd.frame_bytes - cd.offset[planes]), cudaMemcpyDeviceToDevice, c->stream));
        for (int pl = 0; pl < planes && pl < cd.nplanes; pl++) {
            dim3 blk(32, 8), grid((cd.w[pl] + 31) / 32, (cd.h[pl] + 7) / 8);
            if (cd.bps == 1)
                k_fallback<uint8_t><<<grid, blk, 0, c->stream>>>(cd.w[pl], cd.h[pl], time256, p->blend, src + cd.offset[pl], ref + cd.offset[pl], dst + cd.offset[pl], cd.pitch[pl], flags);
            else
                k_fallback<uint16_t><<<grid, blk, 0, c->stream>>>(cd.w[pl], cd.h[pl], time256, p->blend, src + cd.offset[pl], ref + cd.offset[pl], dst + cd.offset[pl], cd.pitch[pl], flags);
            c->launches++;
        }
        MVB_CUDA_CHECK(cudaGetLastError());
    }
    std::vector<int> pairs(usable_out ? 2 * (size_t)out_count : 0);
    if (usable_out)
        MVB_CUDA_CHECK(cudaMemcpyAsync(pairs.data(), B.dflags, sizeof(int) * pairs.size(), cudaMemcpyDeviceToHost, c->stream));
    if (chain_back(ctx, B, out_count, host_dst)) // (synchronises the stream)
        return 1;
    for (int k = 0; usable_out && k < out_count; k++)
        usable_out[k] = pairs[2 * k] && pairs[2 * k + 1];
    return 0;
}

} // extern "C"
