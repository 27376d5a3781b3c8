// mv.Compensate on the GPU (mvcompensateGetFrame, MVCompensate.c:73-374): per block either the
// reference block at mv*time/256 (SAD < thSAD) or the source block at the grid position, blitted or
// overlap-added with the 9 raised-cosine windows, then ToPixels; uncovered strips and the scene-change
// case are plain copies out of the super frame's level-0 interior.  As for Degrain, the block loop of
// the reference becomes one gather kernel per plane: each thread owns one output sample and visits the
// <= 4 blocks covering it.
#include "engine.h"

#include <algorithm>
#include <cmath>
#include <vector>

namespace mvb200 {

void overlap_windows_host(int nx, int ny, int ox, int oy, std::vector<short> &out); // degrain.cu

namespace {

struct VEC4 { // VECTOR with 32-bit fields only (blob arrays are 4-byte aligned)
    int x, y, sad_lo, sad_hi;
    __device__ __forceinline__ long long sad() const { return (long long)(((unsigned long long)(unsigned)sad_hi << 32) | (unsigned)sad_lo); }
};

struct CompDev {
    int bits, nBlkX, nBlkY, pel, logPel, time256, fieldShift, scBehavior, haveRef;
    long long thSAD, nSCD1;
    int nSCD2;
    const uint8_t *vectors;
};

struct CompPlane {
    int plane, W, H, WB, HB, bw, bh, ox, oy, xSub, ySub, stepX0, stepY0, hpad, vpad, copyRight, copyBottom;
    long long dst_pitch, dst_off, sup_pitch, sup_off, sup_sub;
    const short *win;
};

__device__ __forceinline__ const VEC4 *level0(const uint8_t *blob, int nBlk) { return (const VEC4 *)(blob + __ldg((const int *)blob) - 16 * nBlk); }

// fgopIsUsable (Fakery.c:52-58,144-146), first half: the number of blocks whose SAD exceeds thscd1, counted by many CTAs
// (grid = (chunks, frames); one CTA scanning a 4K frame's 130 000 vectors alone took 300 us, longer than the compensation
// itself).  count[f] must be zero before the launch.
__global__ void k_comp_count(int nBlk, long long nSCD1, const uint8_t *__restrict__ vectors, long long vstride, int *__restrict__ count) {
    const uint8_t *blob = vectors + (long long)blockIdx.y * vstride;
    const VEC4 *v = level0(blob, nBlk);
    int cnt = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nBlk; i += gridDim.x * blockDim.x)
        cnt += v[i].sad() > nSCD1 ? 1 : 0;
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0 && cnt)
        atomicAdd(&count[blockIdx.y], cnt);
}

// ... second half, by every consumer thread: usable = valid blob && count <= thscd2 && a reference frame exists
__device__ __forceinline__ bool comp_usable(const CompDev &D, const int *count) { return ((const int *)D.vectors)[1] == 1 && !(*count > D.nSCD2) && D.haveRef; }

// A frame of a batch (blockIdx.z): which frames it reads, MVCompensate.c:80-87,188-223
struct CompBatch {
    const uint8_t *sup;
    uint8_t *dst;
    const uint8_t *vectors;
    long long sup_frame_bytes, dst_frame_bytes, vstride;
    int first, num_frames, delta, backward;
};

template <typename T>
__device__ __forceinline__ void compensate_sample(const CompDev &D, const CompPlane &B, const uint8_t *__restrict__ src_super, const uint8_t *__restrict__ ref_super, uint8_t *__restrict__ dst, const int *__restrict__ count) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= B.W || y >= B.H)
        return;
    T *drow = (T *)(dst + B.dst_off + (long long)y * B.dst_pitch);
    const long long interior = B.sup_off + (long long)(y + B.vpad) * B.sup_pitch + (long long)(x + B.hpad) * (int)sizeof(T);
    if (!comp_usable(D, count)) { // MVCompensate.c:348-364: whole-frame copy of the source (or reference) super frame's interior
        const uint8_t *s = (!D.scBehavior && D.haveRef) ? ref_super : src_super;
        drow[x] = *(const T *)(s + interior);
        return;
    }
    if (x >= B.WB || y >= B.HB) { // :330-342
        const uint8_t *s = D.scBehavior ? src_super : ref_super;
        drow[x] = *(const T *)(s + interior);
        return;
    }
    const VEC4 *vec = level0(D.vectors, D.nBlkX * D.nBlkY);
    const int pixelMax = (1 << D.bits) - 1;
    const int sx = B.bw - B.ox, sy = B.bh - B.oy;
    int bx1, by1, bx0, by0;
    if (!B.win) {
        bx0 = bx1 = x / B.bw;
        by0 = by1 = y / B.bh;
    } else {
        bx1 = min(x / sx, D.nBlkX - 1);
        by1 = min(y / sy, D.nBlkY - 1);
        bx0 = (bx1 >= 1 && x - bx1 * sx < B.ox) ? bx1 - 1 : bx1;
        by0 = (by1 >= 1 && y - by1 * sy < B.oy) ? by1 - 1 : by1;
    }
    unsigned acc = 0;
    int out = 0;
    for (int by = by0; by <= by1; by++)
        for (int bx = bx0; bx <= bx1; bx++) {
            const VEC4 v = vec[by * D.nBlkX + bx];
            int blx, bly;
            const uint8_t *frame;
            if (v.sad() < D.thSAD) { // :238-242 / :286-290 (C division: truncation toward zero)
                blx = (bx * B.stepX0) * D.pel + v.x * D.time256 / 256;
                bly = (by * B.stepY0) * D.pel + v.y * D.time256 / 256 + D.fieldShift;
                frame = ref_super;
            } else {
                blx = (bx * B.stepX0) * D.pel;
                bly = (by * B.stepY0) * D.pel + D.fieldShift;
                frame = src_super;
            }
            const int ax = (B.plane ? blx >> B.xSub : blx) + B.hpad * D.pel;
            const int ay = (B.plane ? bly >> B.ySub : bly) + B.vpad * D.pel;
            const int m = D.pel - 1;
            const int idx = (ax & m) | ((ay & m) << D.logPel);
            const int ix = B.win ? x - bx * sx : x - bx * B.bw, iy = B.win ? y - by * sy : y - by * B.bh;
            const int px = (int)*(const T *)(frame + B.sup_off + idx * B.sup_sub + (long long)((ay >> D.logPel) + iy) * B.sup_pitch + (long long)((ax >> D.logPel) + ix) * (int)sizeof(T));
            if (!B.win) {
                out = px;
            } else {
                const int wby = ((by + D.nBlkY - 3) / (D.nBlkY - 2)) * 3;
                const int wbx = bx == D.nBlkX - 1 ? 2 : (bx == 0 ? 0 : 1);
                acc += (unsigned)((px * (int)B.win[(wby + wbx) * B.bw * B.bh + iy * B.bw + ix]) >> 6);
                if (sizeof(T) == 1)
                    acc &= 0xffffu;
            }
        }
    if (B.win)
        out = min(pixelMax, (int)((acc + 16) >> 5));
    drow[x] = (T)out;
}

template <typename T>
__global__ void k_compensate(CompDev D, CompPlane B, const uint8_t *__restrict__ src_super, const uint8_t *__restrict__ ref_super, uint8_t *__restrict__ dst, const int *__restrict__ count) {
    compensate_sample<T>(D, B, src_super, ref_super, dst, count);
}

// the frames of a batch in one launch (grid.z = frame; relative or static reference, no field shift)
template <typename T>
__global__ void k_compensate_batch(CompDev D, CompPlane B, CompBatch Q, const int *__restrict__ count) {
    const int f = blockIdx.z, fn = Q.first + f;
    int nref = Q.delta > 0 ? fn + (Q.backward ? Q.delta : -Q.delta) : -Q.delta;
    if (nref < 0 || nref >= Q.num_frames)
        nref = -1;
    const uint8_t *src = Q.sup + (long long)fn * Q.sup_frame_bytes;
    D.haveRef = nref >= 0;
    D.vectors = Q.vectors + (long long)f * Q.vstride;
    compensate_sample<T>(D, B, src, nref >= 0 ? Q.sup + (long long)nref * Q.sup_frame_bytes : src, Q.dst + (long long)f * Q.dst_frame_bytes, count + f);
}

} // namespace

// dev_ref_super == nullptr and batch == nullptr: a frame without a reference.  batch != nullptr: n frames [first, first + n)
// of the resident super clip `dev_src_super` in one launch per plane (dev_vectors / dev_dst: the first frame's).
static int compensate_launch(Ctx *ctx, const mvb200_compensate_params *cp, const SuperDesc &sd, const ClipDesc &cd, const uint8_t *dev_src_super, const uint8_t *dev_ref_super, const uint8_t *dev_vectors, int field_shift,
                             uint8_t *dev_dst, bool batch, int first, int n, long long vstride) {
    cudaStream_t st = ctx->stream;
    const mvb200_analysis_data &vd = cp->vd;
    CompDev D;
    memset(&D, 0, sizeof(D));
    D.bits = cp->bits;
    D.nBlkX = vd.nBlkX;
    D.nBlkY = vd.nBlkY;
    D.pel = vd.nPel;
    D.logPel = vd.nPel == 4 ? 2 : (vd.nPel == 2 ? 1 : 0);
    D.time256 = cp->time256;
    D.fieldShift = field_shift;
    D.scBehavior = cp->scBehavior;
    D.haveRef = dev_ref_super != nullptr;
    D.thSAD = cp->thSAD;
    D.nSCD1 = cp->nSCD1;
    D.nSCD2 = cp->nSCD2;
    D.vectors = dev_vectors;
    int *count = (int *)ctx->get_scratch(2, sizeof(int) * (size_t)n);
    if (!count)
        return 1;
    MVB_CUDA_CHECK(cudaMemsetAsync(count, 0, sizeof(int) * (size_t)n, st));
    const int nBlk = vd.nBlkX * vd.nBlkY;
    k_comp_count<<<dim3((unsigned)std::min(64, (nBlk + 255) / 256), (unsigned)n), 256, 0, st>>>(nBlk, cp->nSCD1, dev_vectors, vstride, count);
    ctx->launches++;
    CompBatch Q;
    memset(&Q, 0, sizeof(Q));
    Q.sup = dev_src_super;
    Q.dst = dev_dst;
    Q.vectors = dev_vectors;
    Q.sup_frame_bytes = sd.frame_bytes;
    Q.dst_frame_bytes = cd.frame_bytes;
    Q.vstride = vstride;
    Q.first = first;
    Q.num_frames = cp->num_frames;
    Q.delta = vd.nDeltaFrame;
    Q.backward = vd.isBackward;
    const bool overlap = cp->nOverlapX[0] > 0 || cp->nOverlapY[0] > 0;
    const int num_planes = (cp->super_modeyuv & (MVB200_UPLANE | MVB200_VPLANE)) ? 3 : 1;
    short *dev_win[2] = {nullptr, nullptr};
    if (overlap) { // window tables (built on the host, libm cosf): resident on the device until the configuration changes
        const int key[8] = {cp->nBlkSizeX[0], cp->nBlkSizeY[0], cp->nOverlapX[0], cp->nOverlapY[0], num_planes > 1 ? cp->nBlkSizeX[1] : 0, num_planes > 1 ? cp->nBlkSizeY[1] : 0,
                            num_planes > 1 ? cp->nOverlapX[1] : 0, num_planes > 1 ? cp->nOverlapY[1] : 0};
        if (!ctx->cwin_valid || memcmp(key, ctx->cwin_key, sizeof(key)) != 0) {
            std::vector<short> wl, wc;
            overlap_windows_host(cp->nBlkSizeX[0], cp->nBlkSizeY[0], cp->nOverlapX[0], cp->nOverlapY[0], wl);
            if (num_planes > 1)
                overlap_windows_host(cp->nBlkSizeX[1], cp->nBlkSizeY[1], cp->nOverlapX[1], cp->nOverlapY[1], wc);
            size_t bl = wl.size() * sizeof(short), bc = wc.size() * sizeof(short), offc = (bl + 255) & ~(size_t)255;
            // first use of this configuration (or a change): drain, rebuild the tables, upload synchronously
            uint8_t *dev = (uint8_t *)ctx->get_scratch(6, offc + bc + 256);
            if (!dev)
                return 1;
            MVB_CUDA_CHECK(cudaStreamSynchronize(st)); // (also past the clearing of a fresh scratch buffer)
            MVB_CUDA_CHECK(upload_table(ctx, dev, wl.data(), bl));
            if (bc)
                MVB_CUDA_CHECK(upload_table(ctx, dev + offc, wc.data(), bc));
            ctx->cwin_dev[0] = (short *)dev;
            ctx->cwin_dev[1] = (short *)(dev + offc);
            memcpy(ctx->cwin_key, key, sizeof(key));
            ctx->cwin_valid = true;
        }
        dev_win[0] = ctx->cwin_dev[0];
        dev_win[1] = ctx->cwin_dev[1];
    }
    for (int plane = 0; plane < num_planes; plane++) {
        CompPlane B;
        memset(&B, 0, sizeof(B));
        B.plane = plane;
        B.W = cp->nWidth[plane];
        B.H = cp->nHeight[plane];
        B.WB = cp->nWidth_B[0] < cp->nWidth[0] ? cp->nWidth_B[plane] : B.W;
        B.HB = cp->nHeight_B[0] < cp->nHeight[0] ? cp->nHeight_B[plane] : B.H;
        B.bw = cp->nBlkSizeX[plane];
        B.bh = cp->nBlkSizeY[plane];
        B.ox = cp->nOverlapX[plane];
        B.oy = cp->nOverlapY[plane];
        B.xSub = plane ? cp->ssw : 0;
        B.ySub = plane ? cp->ssh : 0;
        B.stepX0 = cp->nBlkSizeX[0] - cp->nOverlapX[0];
        B.stepY0 = cp->nBlkSizeY[0] - cp->nOverlapY[0];
        const mvb200_level_geom &g = sd.lv[plane][0];
        B.hpad = g.hpad;
        B.vpad = g.vpad;
        B.dst_pitch = cd.pitch[plane];
        B.dst_off = cd.offset[plane];
        B.sup_pitch = sd.pitch[plane];
        B.sup_off = sd.offset[plane];
        B.sup_sub = sd.pitch[plane] * g.padded_h;
        B.win = overlap ? dev_win[plane ? 1 : 0] : nullptr;
        dim3 blk(32, 8), grid((B.W + 31) / 32, (B.H + 7) / 8, batch ? (unsigned)n : 1u);
        if (batch) {
            if (sd.bps == 1)
                k_compensate_batch<uint8_t><<<grid, blk, 0, st>>>(D, B, Q, count);
            else
                k_compensate_batch<uint16_t><<<grid, blk, 0, st>>>(D, B, Q, count);
        } else if (sd.bps == 1) {
            k_compensate<uint8_t><<<grid, blk, 0, st>>>(D, B, dev_src_super, dev_ref_super ? dev_ref_super : dev_src_super, dev_dst, count);
        } else {
            k_compensate<uint16_t><<<grid, blk, 0, st>>>(D, B, dev_src_super, dev_ref_super ? dev_ref_super : dev_src_super, dev_dst, count);
        }
        ctx->launches++;
    }
    MVB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int launch_compensate(Ctx *ctx, const mvb200_compensate_params *cp, const SuperDesc &sd, const ClipDesc &cd, const uint8_t *dev_src_super, const uint8_t *dev_ref_super, const uint8_t *dev_vectors, int field_shift, uint8_t *dev_dst) {
    return compensate_launch(ctx, cp, sd, cd, dev_src_super, dev_ref_super, dev_vectors, field_shift, dev_dst, false, 0, 1, 0);
}

// Frames [first, first + n) of a resident super clip (no field handling: the caller checks cp->fields): one launch per plane
int launch_compensate_batch(Ctx *ctx, const mvb200_compensate_params *cp, const SuperDesc &sd, const ClipDesc &cd, const uint8_t *dev_super, int first, int n, const uint8_t *dev_vectors, long long vectors_stride,
                            uint8_t *dev_dst) {
    return compensate_launch(ctx, cp, sd, cd, dev_super, dev_super, dev_vectors, 0, dev_dst, true, first, n, vectors_stride);
}

} // namespace mvb200
