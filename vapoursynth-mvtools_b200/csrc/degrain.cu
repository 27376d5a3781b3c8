// mv.Degrain1..6 on the GPU (mvdegrainGetFrame<radius>, MVDegrains.cpp:85-330).
//
// The reference walks blocks: per block it blends src + 2r motion-compensated
// reference blocks (Degrain_C, MVDegrains.h:31-53), multiplies by one of 9 overlap
// windows and accumulates into a 16/32-bit plane (overlaps_c, Overlap.cpp:144-158),
// then rounds (ToPixels, Overlap.cpp:335-356) and limits (LimitChanges_C,
// MVDegrains.h:168-181).  All of that is order-independent integer arithmetic, so
// here it is ONE gather kernel per plane: each thread owns one output sample, visits
// the <=4 blocks covering it, and writes the final pixel once (coalesced store) --
// no accumulator plane, no atomics, no second pass.
//
// Three launches per call, all batched over frames:
//   k_scd     usable[r] = valid && !sceneChange     (fgopIsUsable, Fakery.c:52-58,144-146)
//   k_weights per block: DegrainWeight + normaliseWeights (MVDegrains.h:184-223) for the
//             luma and chroma thresholds -- the only FP64 on this path -- stored as int16
//   k_blend   the fused gather described above, per plane
#include "engine.h"

#include <cmath>
#include <vector>

namespace mvb200 {

namespace {

constexpr int kMaxR2 = 2 * MVB200_MAX_RADIUS;

struct DegrainDev {
    int R2, bps, bits, nplanes, logPel, pel;
    int nBlkX, nBlkY, nLvCount;
    long long thSAD[2]; // luma, chroma
    long long nSCD1;
    int nSCD2;
    int delta[kMaxR2], backward[kMaxR2];
    int num_frames, first;
    long long vectors_stride;
    const uint8_t *vectors[kMaxR2];
};

struct BlockW { // normalised weights of one block: [0] luma planes, [1] chroma planes
    short wsrc[2];
    short w[2][kMaxR2];
    short pad[2];
};

struct VEC { // the blob's VECTOR arrays are only 4-byte aligned: 32-bit fields only
    int x, y;
    int sad_lo, sad_hi;
    __device__ __forceinline__ long long sad() const { return (long long)(((unsigned long long)(unsigned)sad_hi << 32) | (unsigned)sad_lo); }
};

// The finest level (the one Degrain reads: fgopGetBlock(fgop, 0, i), Fakery.c:110-133) is stored LAST in
// the blob (coarsest first, GroupOfPlanes.c:76-85), also for `divide` clips: it is the final 16*nBlk bytes.
__device__ __forceinline__ const VEC *blob_level0(const uint8_t *blob, int nBlk) {
    return (const VEC *)(blob + __ldg((const int *)blob) - 16 * nBlk);
}

// one CTA per (r, frame)
__global__ void k_scd(DegrainDev D, int *__restrict__ usable) {
    int r = blockIdx.x, f = blockIdx.y;
    const uint8_t *blob = D.vectors[r] + (long long)f * D.vectors_stride;
    int valid = ((const int *)blob)[1] == 1;
    const VEC *v = blob_level0(blob, D.nBlkX * D.nBlkY);
    int n = D.nBlkX * D.nBlkY;
    int cnt = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        cnt += v[i].sad() > D.nSCD1 ? 1 : 0;
    __shared__ int s;
    if (threadIdx.x == 0)
        s = 0;
    __syncthreads();
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0)
        atomicAdd(&s, cnt);
    __syncthreads();
    if (threadIdx.x == 0) {
        int n_ref = D.first + f + (D.backward[r] ? D.delta[r] : -D.delta[r]);
        int ok = valid && !(s > D.nSCD2) && n_ref >= 0 && n_ref < D.num_frames;
        usable[f * D.R2 + r] = ok;
    }
}

// DegrainWeight, MVDegrains.h:184-189
__device__ __forceinline__ int degrain_weight(long long thSAD, long long blockSAD) {
    if (blockSAD >= thSAD)
        return 0;
    return __double2int_rz(__ddiv_rn((double)((thSAD - blockSAD) * (thSAD + blockSAD) * 256), (double)(thSAD * thSAD + blockSAD * blockSAD)));
}

// thread per (block, frame)
__global__ void k_weights(DegrainDev D, const int *__restrict__ usable, BlockW *__restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int f = blockIdx.y;
    int n = D.nBlkX * D.nBlkY;
    if (i >= n)
        return;
    long long sad[kMaxR2];
    for (int r = 0; r < D.R2; r++) {
        if (usable[f * D.R2 + r]) {
            const VEC *v = blob_level0(D.vectors[r] + (long long)f * D.vectors_stride, n);
            sad[r] = v[i].sad();
        } else {
            sad[r] = -1;
        }
    }
    BlockW bw;
    for (int c = 0; c < 2; c++) {
        int W[kMaxR2];
        int WSum = 256 + 1;
        for (int r = 0; r < D.R2; r++) {
            W[r] = sad[r] >= 0 ? degrain_weight(D.thSAD[c], sad[r]) : 0;
            WSum += W[r];
        }
        // normaliseWeights, MVDegrains.h:209-223
        double scale = __ddiv_rn(256.0, (double)WSum);
        int WSrc = 256;
        for (int r = 0; r < D.R2; r++) {
            W[r] = __double2int_rz(__dmul_rn((double)W[r], scale));
            WSrc -= W[r];
            bw.w[c][r] = (short)W[r];
        }
        for (int r = D.R2; r < kMaxR2; r++)
            bw.w[c][r] = 0;
        bw.wsrc[c] = (short)WSrc;
    }
    bw.pad[0] = bw.pad[1] = 0;
    out[(long long)f * n + i] = bw;
}

struct BlendPlane {
    int plane, cls;       // cls: 0 luma, 1 chroma weights
    int W, H, WB, HB;     // plane size, block-covered size
    int bw, bh, ox, oy;   // block / overlap size in this plane
    int xSub, ySub;
    int limit, copyRight, copyBottom, process;
    long long src_pitch, src_off, dst_pitch, dst_off, sup_pitch, sup_off, sup_sub;
    int hpad_pel, vpad_pel;
    int stepX0, stepY0;   // luma block step (for the luma-unit block origin)
    const short *win;     // 9 windows (device), or nullptr without overlap
};

template <typename T>
__device__ __forceinline__ int blend_block(const DegrainDev &D, const BlendPlane &B, const BlockW &w, const uint8_t *sup_frames, long long sup_frame_bytes, int f,
                                           const int *usable, int bx, int by, int ix, int iy, int srcv) {
    // Degrain_C, MVDegrains.h:31-53 for the sample (ix, iy) of block (bx, by)
    int sum = 128 + srcv * w.wsrc[B.cls];
    int bi = by * D.nBlkX + bx;
    for (int r = 0; r < D.R2; r++) {
        int wr = w.w[B.cls][r];
        if (wr == 0)
            continue; // unusable reference or zero weight: contributes nothing
        const VEC *v = blob_level0(D.vectors[r] + (long long)f * D.vectors_stride, D.nBlkX * D.nBlkY);
        int vx = v[bi].x, vy = v[bi].y;
        // useBlock, MVDegrains.h:192-206 + mvpGetPointer, MVFrame.cpp:1733-1755
        int blx = ((bx * B.stepX0) << D.logPel) + vx;
        int bly = ((by * B.stepY0) << D.logPel) + vy;
        int ax = (B.plane ? blx >> B.xSub : blx) + B.hpad_pel;
        int ay = (B.plane ? bly >> B.ySub : bly) + B.vpad_pel;
        int m = D.pel - 1;
        int idx = (ax & m) | ((ay & m) << D.logPel);
        int n_ref = D.first + f + (D.backward[r] ? D.delta[r] : -D.delta[r]);
        const T *p = (const T *)(sup_frames + (long long)n_ref * sup_frame_bytes + B.sup_off + idx * B.sup_sub + (long long)((ay >> D.logPel) + iy) * B.sup_pitch) + (ax >> D.logPel) + ix;
        sum += (int)__ldg(p) * wr;
    }
    (void)usable;
    return (int)(T)(sum >> 8);
}

template <typename T>
__global__ void k_blend(DegrainDev D, BlendPlane B, const uint8_t *__restrict__ src_frames, long long src_frame_bytes, const uint8_t *__restrict__ sup_frames,
                        long long sup_frame_bytes, uint8_t *__restrict__ dst_frames, long long dst_frame_bytes, const int *__restrict__ usable,
                        const BlockW *__restrict__ weights) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    int f = blockIdx.z;
    if (x >= B.W || y >= B.H)
        return;
    const T *srow = (const T *)(src_frames + (long long)(D.first + f) * src_frame_bytes + B.src_off + (long long)y * B.src_pitch);
    T *drow = (T *)(dst_frames + (long long)f * dst_frame_bytes + B.dst_off + (long long)y * B.dst_pitch);
    int srcv = srow[x];
    if (!B.process || x >= B.WB || y >= B.HB) {
        // unprocessed plane (:211-214) or the uncovered right/bottom strips (:238-249, :290-298)
        drow[x] = (T)srcv;
        return;
    }
    const BlockW *wf = weights + (long long)f * D.nBlkX * D.nBlkY;
    int pixelMax = (1 << D.bits) - 1;
    int out;
    if (!B.win) {
        int bx = x / B.bw, by = y / B.bh;
        out = blend_block<T>(D, B, wf[by * D.nBlkX + bx], sup_frames, sup_frame_bytes, f, usable, bx, by, x - bx * B.bw, y - by * B.bh, srcv);
    } else {
        int sx = B.bw - B.ox, sy = B.bh - B.oy;
        int bx1 = min(x / sx, D.nBlkX - 1), by1 = min(y / sy, D.nBlkY - 1);
        int bx0 = (bx1 >= 1 && x - bx1 * sx < B.ox) ? bx1 - 1 : bx1;
        int by0 = (by1 >= 1 && y - by1 * sy < B.oy) ? by1 - 1 : by1;
        unsigned acc = 0;
        for (int by = by0; by <= by1; by++) {
            int wby = ((by + D.nBlkY - 3) / (D.nBlkY - 2)) * 3; // MVDegrains.cpp:256
            int iy = y - by * sy;
            for (int bx = bx0; bx <= bx1; bx++) {
                int wbx = bx == D.nBlkX - 1 ? 2 : (bx == 0 ? 0 : 1); // :257-262
                int ix = x - bx * sx;
                int v = blend_block<T>(D, B, wf[by * D.nBlkX + bx], sup_frames, sup_frame_bytes, f, usable, bx, by, ix, iy, srcv);
                int wv = B.win[(wby + wbx) * B.bw * B.bh + iy * B.bw + ix];
                acc += (unsigned)((v * wv) >> 6); // overlaps_c, Overlap.cpp:144-158
                if (sizeof(T) == 1)
                    acc &= 0xffffu; // the reference accumulates 8-bit video in uint16_t
            }
        }
        out = min(pixelMax, (int)((acc + 16) >> 5)); // ToPixels, Overlap.cpp:335-356
    }
    if (B.limit < pixelMax) // LimitChanges_C, MVDegrains.h:168-181
        out = min(max(out, srcv - B.limit), srcv + B.limit);
    drow[x] = (T)out;
}


// ---- four samples per thread ------------------------------------------------------------
// When block size, step and overlap are multiples of 4 samples, the 4 samples x..x+3 (x % 4 == 0)
// are covered by the same blocks, so the per-block work (weights, vectors, sub-plane addressing) is
// done once per 4 outputs and the pixels move as 32/64-bit words.
template <typename T>
struct Quad {
    int v[4];
};

template <typename T>
__device__ __forceinline__ void load4(const uint8_t *p, int (&o)[4]) {
    // p: address of 4 consecutive samples with sample alignment only
    unsigned long long a = (unsigned long long)p;
    if (sizeof(T) == 1) {
        const unsigned *q = (const unsigned *)(a & ~3ull);
        unsigned w = __funnelshift_r(__ldg(q), __ldg(q + 1), ((unsigned)a & 3u) * 8u);
        o[0] = w & 255;
        o[1] = (w >> 8) & 255;
        o[2] = (w >> 16) & 255;
        o[3] = w >> 24;
    } else {
        const unsigned *q = (const unsigned *)(a & ~3ull);
        unsigned sh = ((unsigned)a & 2u) * 8u;
        unsigned w0 = __ldg(q), w1 = __ldg(q + 1), w2 = sh ? __ldg(q + 2) : 0u;
        unsigned lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
        o[0] = lo & 0xffff;
        o[1] = lo >> 16;
        o[2] = hi & 0xffff;
        o[3] = hi >> 16;
    }
}

template <typename T>
__device__ __forceinline__ void blend_block4(const DegrainDev &D, const BlendPlane &B, const BlockW &w, const uint8_t *sup_frames, long long sup_frame_bytes, int f, int bx,
                                             int by, int ix, int iy, const int (&srcv)[4], int (&out)[4]) {
    int sum[4];
#pragma unroll
    for (int i = 0; i < 4; i++)
        sum[i] = 128 + srcv[i] * w.wsrc[B.cls];
    const int bi = by * D.nBlkX + bx;
    for (int r = 0; r < D.R2; r++) {
        const int wr = w.w[B.cls][r];
        if (wr == 0)
            continue;
        const VEC *v = blob_level0(D.vectors[r] + (long long)f * D.vectors_stride, D.nBlkX * D.nBlkY);
        const int vx = v[bi].x, vy = v[bi].y;
        const int blx = ((bx * B.stepX0) << D.logPel) + vx;
        const int bly = ((by * B.stepY0) << D.logPel) + vy;
        const int ax = (B.plane ? blx >> B.xSub : blx) + B.hpad_pel;
        const int ay = (B.plane ? bly >> B.ySub : bly) + B.vpad_pel;
        const int m = D.pel - 1;
        const int idx = (ax & m) | ((ay & m) << D.logPel);
        const int n_ref = D.first + f + (D.backward[r] ? D.delta[r] : -D.delta[r]);
        const uint8_t *p = sup_frames + (long long)n_ref * sup_frame_bytes + B.sup_off + idx * B.sup_sub + (long long)((ay >> D.logPel) + iy) * B.sup_pitch +
                           (long long)((ax >> D.logPel) + ix) * (int)sizeof(T);
        int rv[4];
        load4<T>(p, rv);
#pragma unroll
        for (int i = 0; i < 4; i++)
            sum[i] += rv[i] * wr;
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
        out[i] = (int)(T)(sum[i] >> 8);
}

template <typename T>
__global__ void k_blend4(DegrainDev D, BlendPlane B, const uint8_t *__restrict__ src_frames, long long src_frame_bytes, const uint8_t *__restrict__ sup_frames,
                         long long sup_frame_bytes, uint8_t *__restrict__ dst_frames, long long dst_frame_bytes, const BlockW *__restrict__ weights) {
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int f = blockIdx.z;
    if (x >= B.W || y >= B.H)
        return;
    const uint8_t *srow = src_frames + (long long)(D.first + f) * src_frame_bytes + B.src_off + (long long)y * B.src_pitch;
    uint8_t *drow = dst_frames + (long long)f * dst_frame_bytes + B.dst_off + (long long)y * B.dst_pitch;
    int srcv[4], out[4];
    load4<T>(srow + (long long)x * (int)sizeof(T), srcv);
    const int pixelMax = (1 << D.bits) - 1;
    const int nvalid = min(4, B.W - x);
    // a quad never straddles the covered / uncovered boundary unless WB % 4 != 0: handle per sample below
    const bool covered = B.process && x + 3 < B.WB && y < B.HB;
    if (covered) {
        const BlockW *wf = weights + (long long)f * D.nBlkX * D.nBlkY;
        if (!B.win) {
            const int bx = x / B.bw, by = y / B.bh;
            blend_block4<T>(D, B, wf[by * D.nBlkX + bx], sup_frames, sup_frame_bytes, f, bx, by, x - bx * B.bw, y - by * B.bh, srcv, out);
        } else {
            const int sx = B.bw - B.ox, sy = B.bh - B.oy;
            const int bx1 = min(x / sx, D.nBlkX - 1), by1 = min(y / sy, D.nBlkY - 1);
            const int bx0 = (bx1 >= 1 && x - bx1 * sx < B.ox) ? bx1 - 1 : bx1;
            const int by0 = (by1 >= 1 && y - by1 * sy < B.oy) ? by1 - 1 : by1;
            unsigned acc[4] = {0, 0, 0, 0};
            for (int by = by0; by <= by1; by++) {
                const int wby = ((by + D.nBlkY - 3) / (D.nBlkY - 2)) * 3;
                const int iy = y - by * sy;
                for (int bx = bx0; bx <= bx1; bx++) {
                    const int wbx = bx == D.nBlkX - 1 ? 2 : (bx == 0 ? 0 : 1);
                    const int ix = x - bx * sx;
                    int v[4];
                    blend_block4<T>(D, B, wf[by * D.nBlkX + bx], sup_frames, sup_frame_bytes, f, bx, by, ix, iy, srcv, v);
                    const short4 wv = *(const short4 *)(B.win + (wby + wbx) * B.bw * B.bh + iy * B.bw + ix);
                    const int ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        acc[i] += (unsigned)((v[i] * ww[i]) >> 6);
                        if (sizeof(T) == 1)
                            acc[i] &= 0xffffu;
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; i++)
                out[i] = min(pixelMax, (int)((acc[i] + 16) >> 5));
        }
        if (B.limit < pixelMax) {
#pragma unroll
            for (int i = 0; i < 4; i++)
                out[i] = min(max(out[i], srcv[i] - B.limit), srcv[i] + B.limit);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; i++)
            out[i] = srcv[i];
    }
    T *d = (T *)drow + x;
    if (nvalid == 4) {
        if (sizeof(T) == 1)
            *(unsigned *)d = (unsigned)out[0] | ((unsigned)out[1] << 8) | ((unsigned)out[2] << 16) | ((unsigned)out[3] << 24);
        else
            *(uint2 *)d = make_uint2((unsigned)out[0] | ((unsigned)out[1] << 16), (unsigned)out[2] | ((unsigned)out[3] << 16));
    } else {
        for (int i = 0; i < nvalid; i++)
            d[i] = (T)out[i];
    }
}

// ---- tiled variant of the quad kernel -----------------------------------------------------
// A CTA owns a TW x TH tile of the plane.  Phase 1: the CTA's threads resolve, ONCE per (block, reference)
// pair that can touch the tile, the weight and the address of the motion-compensated block (vector load,
// sub-plane selection, padding) into shared memory.  Phase 2: every thread blends one quad, walking the
// <= 4 covering blocks x 2r references with one LDS.128 + two loads + a funnel shift per reference; for
// 8-bit clips two samples share a register (the weighted sum fits 16 bits, MVDegrains.h:31-53).
constexpr int kTileW = 64, kTileH = 8, kTileMaxEntries = 1024, kTileMaxBlocks = 256;

template <typename T>
__global__ void __launch_bounds__(kTileW / 4 * kTileH) k_blend_tile(DegrainDev D, BlendPlane B, const uint8_t *__restrict__ src_frames, long long src_frame_bytes,
                                                                  const uint8_t *__restrict__ sup_frames, long long sup_frame_bytes, uint8_t *__restrict__ dst_frames,
                                                                  long long dst_frame_bytes, const BlockW *__restrict__ weights) {
    __shared__ int4 ent[kTileMaxEntries]; // {offset lo, offset hi, weight, -}
    __shared__ int wsrc_s[kTileMaxBlocks];
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthreads = blockDim.x * blockDim.y;
    const int tx0 = blockIdx.x * kTileW, ty0 = blockIdx.y * kTileH;
    const int f = blockIdx.z;
    const int sx = B.bw - B.ox, sy = B.bh - B.oy;
    const int bxlo = max(0, tx0 / sx - 1), bxhi = min(D.nBlkX - 1, (tx0 + kTileW - 1) / sx);
    const int bylo = max(0, ty0 / sy - 1), byhi = min(D.nBlkY - 1, (ty0 + kTileH - 1) / sy);
    const int nbx = bxhi - bxlo + 1, nby = byhi - bylo + 1;
    const int nblk = D.nBlkX * D.nBlkY;
    if (B.process) {
        const BlockW *wf = weights + (long long)f * nblk;
        for (int e = tid; e < nbx * nby * D.R2; e += nthreads) {
            const int b = e / D.R2, r = e - b * D.R2;
            const int lby = b / nbx, lbx = b - lby * nbx;
            const int bx = bxlo + lbx, by = bylo + lby, bi = by * D.nBlkX + bx;
            const int wr = wf[bi].w[B.cls][r];
            long long off = 0;
            if (wr != 0) {
                const VEC *v = blob_level0(D.vectors[r] + (long long)f * D.vectors_stride, nblk);
                const int vx = v[bi].x, vy = v[bi].y;
                // useBlock, MVDegrains.h:192-206 + mvpGetPointer, MVFrame.cpp:1733-1755
                const int blx = ((bx * B.stepX0) << D.logPel) + vx;
                const int bly = ((by * B.stepY0) << D.logPel) + vy;
                const int ax = (B.plane ? blx >> B.xSub : blx) + B.hpad_pel;
                const int ay = (B.plane ? bly >> B.ySub : bly) + B.vpad_pel;
                const int m = D.pel - 1;
                const int idx = (ax & m) | ((ay & m) << D.logPel);
                const int n_ref = D.first + f + (D.backward[r] ? D.delta[r] : -D.delta[r]);
                off = (long long)n_ref * sup_frame_bytes + B.sup_off + idx * B.sup_sub + (long long)(ay >> D.logPel) * B.sup_pitch + (long long)(ax >> D.logPel) * (int)sizeof(T);
            }
            ent[e] = make_int4((int)(unsigned)(unsigned long long)off, (int)(unsigned)((unsigned long long)off >> 32), wr, 0);
            if (r == 0)
                wsrc_s[b] = wf[bi].wsrc[B.cls];
        }
    }
    __syncthreads();
    const int x = tx0 + threadIdx.x * 4, y = ty0 + threadIdx.y;
    if (x >= B.W || y >= B.H)
        return;
    const uint8_t *srow = src_frames + (long long)(D.first + f) * src_frame_bytes + B.src_off + (long long)y * B.src_pitch;
    uint8_t *drow = dst_frames + (long long)f * dst_frame_bytes + B.dst_off + (long long)y * B.dst_pitch;
    int srcv[4], out[4];
    load4<T>(srow + (long long)x * (int)sizeof(T), srcv);
    const int pixelMax = (1 << D.bits) - 1;
    const int nvalid = min(4, B.W - x);
    const bool covered = B.process && x + 3 < B.WB && y < B.HB;
    if (covered) {
        int bx0, bx1, by0, by1;
        if (!B.win) {
            bx0 = bx1 = x / B.bw;
            by0 = by1 = y / B.bh;
        } else {
            bx1 = min(x / sx, D.nBlkX - 1);
            by1 = min(y / sy, D.nBlkY - 1);
            bx0 = (bx1 >= 1 && x - bx1 * sx < B.ox) ? bx1 - 1 : bx1;
            by0 = (by1 >= 1 && y - by1 * sy < B.oy) ? by1 - 1 : by1;
        }
        unsigned acc[4] = {0, 0, 0, 0};
        for (int by = by0; by <= by1; by++) {
            const int iy = y - by * sy;
            for (int bx = bx0; bx <= bx1; bx++) {
                const int ix = x - bx * sx;
                const int lb = (by - bylo) * nbx + (bx - bxlo);
                const int ws = wsrc_s[lb];
                const long long rowoff = (long long)iy * B.sup_pitch + ix * (int)sizeof(T);
                int v[4];
                if (sizeof(T) == 1) {
                    // two samples per register: even samples in sumA, odd in sumB; each half stays below 2^16
                    const unsigned sw = (unsigned)srcv[0] | ((unsigned)srcv[2] << 16), so = (unsigned)srcv[1] | ((unsigned)srcv[3] << 16);
                    unsigned sumA = 0x00800080u + sw * (unsigned)ws, sumB = 0x00800080u + so * (unsigned)ws;
                    for (int r = 0; r < D.R2; r++) {
                        const int4 e = ent[lb * D.R2 + r];
                        if (e.z == 0)
                            continue; // unusable reference or zero weight: contributes nothing
                        const uint8_t *p = sup_frames + (((long long)e.y << 32) | (unsigned)e.x) + rowoff;
                        const unsigned long long a = (unsigned long long)p;
                        const unsigned *q = (const unsigned *)(a & ~3ull);
                        const unsigned w = __funnelshift_r(__ldg(q), __ldg(q + 1), ((unsigned)a & 3u) * 8u);
                        sumA += (w & 0x00ff00ffu) * (unsigned)e.z;
                        sumB += ((w >> 8) & 0x00ff00ffu) * (unsigned)e.z;
                    }
                    v[0] = (sumA >> 8) & 0xff;
                    v[2] = sumA >> 24;
                    v[1] = (sumB >> 8) & 0xff;
                    v[3] = sumB >> 24;
                } else {
                    int sum[4];
#pragma unroll
                    for (int i = 0; i < 4; i++)
                        sum[i] = 128 + srcv[i] * ws;
                    for (int r = 0; r < D.R2; r++) {
                        const int4 e = ent[lb * D.R2 + r];
                        if (e.z == 0)
                            continue;
                        int rv[4];
                        load4<T>(sup_frames + (((long long)e.y << 32) | (unsigned)e.x) + rowoff, rv);
#pragma unroll
                        for (int i = 0; i < 4; i++)
                            sum[i] += rv[i] * e.z;
                    }
#pragma unroll
                    for (int i = 0; i < 4; i++)
                        v[i] = (int)(T)(sum[i] >> 8);
                }
                if (!B.win) {
#pragma unroll
                    for (int i = 0; i < 4; i++)
                        out[i] = v[i];
                } else {
                    const int wby = ((by + D.nBlkY - 3) / (D.nBlkY - 2)) * 3;       // MVDegrains.cpp:256
                    const int wbx = bx == D.nBlkX - 1 ? 2 : (bx == 0 ? 0 : 1);       // :257-262
                    const short4 wv = *(const short4 *)(B.win + (wby + wbx) * B.bw * B.bh + iy * B.bw + ix);
                    const int ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        acc[i] += (unsigned)((v[i] * ww[i]) >> 6); // overlaps_c, Overlap.cpp:144-158
                        if (sizeof(T) == 1)
                            acc[i] &= 0xffffu; // the reference accumulates 8-bit video in uint16_t
                    }
                }
            }
        }
        if (B.win) {
#pragma unroll
            for (int i = 0; i < 4; i++)
                out[i] = min(pixelMax, (int)((acc[i] + 16) >> 5)); // ToPixels, Overlap.cpp:335-356
        }
        if (B.limit < pixelMax) { // LimitChanges_C, MVDegrains.h:168-181
#pragma unroll
            for (int i = 0; i < 4; i++)
                out[i] = min(max(out[i], srcv[i] - B.limit), srcv[i] + B.limit);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; i++)
            out[i] = srcv[i];
    }
    T *d = (T *)drow + x;
    if (nvalid == 4) {
        if (sizeof(T) == 1)
            *(unsigned *)d = (unsigned)out[0] | ((unsigned)out[1] << 8) | ((unsigned)out[2] << 16) | ((unsigned)out[3] << 24);
        else
            *(uint2 *)d = make_uint2((unsigned)out[0] | ((unsigned)out[1] << 16), (unsigned)out[2] | ((unsigned)out[3] << 16));
    } else {
        for (int i = 0; i < nvalid; i++)
            d[i] = (T)out[i];
    }
}

// overInit, Overlap.cpp:40-125 -- evaluated on the HOST with the same float/double
// expression (cosf of a double argument), never on the GPU.
void overlap_windows(int nx, int ny, int ox, int oy, std::vector<short> &out) {
    std::vector<float> wx(3 * nx), wy(3 * ny);
    float *wxm = wx.data(), *wxf = wxm + nx, *wxl = wxm + 2 * nx;
    float *wym = wy.data(), *wyf = wym + ny, *wyl = wym + 2 * ny;
    for (int i = 0; i < ox; i++) {
        wxm[i] = cosf(M_PI * (i - ox + 0.5f) / (ox * 2));
        wxm[i] = wxm[i] * wxm[i];
        wxf[i] = 1;
        wxl[i] = wxm[i];
    }
    for (int i = ox; i < nx - ox; i++)
        wxm[i] = wxf[i] = wxl[i] = 1;
    for (int i = nx - ox; i < nx; i++) {
        wxm[i] = cosf(M_PI * (i - nx + ox + 0.5f) / (ox * 2));
        wxm[i] = wxm[i] * wxm[i];
        wxf[i] = wxm[i];
        wxl[i] = 1;
    }
    for (int i = 0; i < oy; i++) {
        wym[i] = cosf(M_PI * (i - oy + 0.5f) / (oy * 2));
        wym[i] = wym[i] * wym[i];
        wyf[i] = 1;
        wyl[i] = wym[i];
    }
    for (int i = oy; i < ny - oy; i++)
        wym[i] = wyf[i] = wyl[i] = 1;
    for (int i = ny - oy; i < ny; i++) {
        wym[i] = cosf(M_PI * (i - ny + oy + 0.5f) / (oy * 2));
        wym[i] = wym[i] * wym[i];
        wyf[i] = wym[i];
        wyl[i] = 1;
    }
    const float *ys[3] = {wyf, wym, wyl};
    const float *xs[3] = {wxf, wxm, wxl};
    out.resize((size_t)9 * nx * ny);
    for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++)
            for (int j = 0; j < ny; j++)
                for (int i = 0; i < nx; i++)
                    out[((size_t)(a * 3 + b) * ny + j) * nx + i] = (short)(int)(ys[a][j] * xs[b][i] * 2048 + 0.5f);
}

} // namespace

void overlap_windows_host(int nx, int ny, int ox, int oy, std::vector<short> &out) { overlap_windows(nx, ny, ox, oy, out); }

} // namespace mvb200

extern "C" void mvb200_overlap_windows(int nx, int ny, int ox, int oy, int16_t *out9) {
    std::vector<short> w;
    mvb200::overlap_windows_host(nx, ny, ox, oy, w);
    memcpy(out9, w.data(), w.size() * sizeof(short));
}

namespace mvb200 {

int launch_degrain(Ctx *ctx, const mvb200_degrain_params *dp, const SuperDesc &sd, const ClipDesc &cd, const uint8_t *dev_src, const uint8_t *dev_super, int num_frames,
                   int first, int n, const uint8_t *const *dev_vectors, long long vectors_stride, uint8_t *dev_dst) {
    if (n <= 0)
        return 0;
    cudaStream_t st = ctx->stream;
    DegrainDev D;
    memset(&D, 0, sizeof(D));
    const mvb200_analysis_data &v0 = dp->vd[0];
    D.R2 = dp->radius * 2;
    D.bps = sd.bps;
    D.bits = dp->bits;
    D.nplanes = dp->nplanes;
    D.pel = v0.nPel;
    D.logPel = v0.nPel == 4 ? 2 : (v0.nPel == 2 ? 1 : 0);
    D.nBlkX = v0.nBlkX;
    D.nBlkY = v0.nBlkY;
    D.nLvCount = v0.nLvCount;
    D.thSAD[0] = dp->thSAD[0];
    D.thSAD[1] = dp->thSAD[1];
    D.nSCD1 = dp->nSCD1;
    D.nSCD2 = dp->nSCD2;
    D.num_frames = num_frames;
    D.first = first;
    D.vectors_stride = vectors_stride;
    for (int r = 0; r < D.R2; r++) {
        D.delta[r] = dp->vd[r].nDeltaFrame;
        D.backward[r] = dp->vd[r].isBackward;
        D.vectors[r] = dev_vectors[r];
    }
    const int nblk = D.nBlkX * D.nBlkY;
    int *usable = (int *)ctx->get_scratch(2, sizeof(int) * (size_t)n * D.R2);
    BlockW *weights = (BlockW *)ctx->get_scratch(3, sizeof(BlockW) * (size_t)n * nblk);
    if (!usable || !weights)
        return 1;
    k_scd<<<dim3(D.R2, n), 256, 0, st>>>(D, usable);
    k_weights<<<dim3((nblk + 127) / 128, n), 128, 0, st>>>(D, usable, weights);
    ctx->launches += 2;

    const bool overlap = dp->nOverlapX[0] > 0 || dp->nOverlapY[0] > 0;
    // overlap windows: luma and chroma tables, built on the host (libm cosf), uploaded once per call
    short *dev_win[2] = {nullptr, nullptr};
    if (overlap) {
        const int key[8] = {dp->nBlkSizeX[0], dp->nBlkSizeY[0], dp->nOverlapX[0], dp->nOverlapY[0], dp->nplanes > 1 ? dp->nBlkSizeX[1] : 0, dp->nplanes > 1 ? dp->nBlkSizeY[1] : 0,
                            dp->nplanes > 1 ? dp->nOverlapX[1] : 0, dp->nplanes > 1 ? dp->nOverlapY[1] : 0};
        if (!ctx->win_valid || memcmp(key, ctx->win_key, sizeof(key)) != 0) {
            std::vector<short> wl, wc;
            overlap_windows(dp->nBlkSizeX[0], dp->nBlkSizeY[0], dp->nOverlapX[0], dp->nOverlapY[0], wl);
            if (dp->nplanes > 1)
                overlap_windows(dp->nBlkSizeX[1], dp->nBlkSizeY[1], dp->nOverlapX[1], dp->nOverlapY[1], wc);
            size_t bl = wl.size() * sizeof(short), bc = wc.size() * sizeof(short);
            size_t offc = (bl + 255) & ~(size_t)255;
            // first use of this configuration (or a change): drain, rebuild the tables, upload synchronously
            uint8_t *dev = (uint8_t *)ctx->get_scratch(4, offc + bc + 256);
            if (!dev)
                return 1;
            MVB_CUDA_CHECK(cudaStreamSynchronize(st)); // (also past the clearing of a fresh scratch buffer)
            MVB_CUDA_CHECK(upload_table(ctx, dev, wl.data(), bl));
            if (bc)
                MVB_CUDA_CHECK(upload_table(ctx, dev + offc, wc.data(), bc));
            ctx->win_dev[0] = (short *)dev;
            ctx->win_dev[1] = (short *)(dev + offc);
            memcpy(ctx->win_key, key, sizeof(key));
            ctx->win_valid = true;
        }
        dev_win[0] = ctx->win_dev[0];
        dev_win[1] = ctx->win_dev[1];
    }
    for (int plane = 0; plane < dp->nplanes; plane++) {
        BlendPlane B;
        memset(&B, 0, sizeof(B));
        B.plane = plane;
        B.cls = plane ? 1 : 0;
        B.W = dp->nWidth[plane];
        B.H = dp->nHeight[plane];
        B.WB = dp->nWidth_B[plane];
        B.HB = dp->nHeight_B[plane];
        B.bw = dp->nBlkSizeX[plane];
        B.bh = dp->nBlkSizeY[plane];
        B.ox = dp->nOverlapX[plane];
        B.oy = dp->nOverlapY[plane];
        B.xSub = plane ? dp->ssw : 0;
        B.ySub = plane ? dp->ssh : 0;
        B.limit = dp->nLimit[plane];
        B.process = dp->process[plane];
        B.src_pitch = cd.pitch[plane];
        B.src_off = cd.offset[plane];
        B.dst_pitch = cd.pitch[plane];
        B.dst_off = cd.offset[plane];
        B.sup_pitch = sd.pitch[plane];
        const mvb200_level_geom &g = sd.lv[plane][0];
        B.sup_off = sd.offset[plane];
        B.sup_sub = sd.pitch[plane] * g.padded_h;
        B.hpad_pel = g.hpad * D.pel;
        B.vpad_pel = g.vpad * D.pel;
        B.stepX0 = dp->nBlkSizeX[0] - dp->nOverlapX[0];
        B.stepY0 = dp->nBlkSizeY[0] - dp->nOverlapY[0];
        B.win = overlap ? dev_win[B.cls] : nullptr;
        // the reference copies the strips when the LUMA coverage is short (:238,:290 test index 0)
        if (!(dp->nWidth_B[0] < dp->nWidth[0]))
            B.WB = B.W;
        if (!(dp->nHeight_B[0] < dp->nHeight[0]))
            B.HB = B.H;
        // 4 samples per thread when every block boundary is a multiple of 4 samples (and rows start 4-aligned)
        const bool quad = B.process && (B.bw % 4 == 0) && (B.ox % 4 == 0) && (B.WB % 4 == 0) && (!overlap || ((B.bw - B.ox) % 4 == 0));
        const int tsx = B.bw - B.ox, tsy = B.bh - B.oy;
        const int tile_blocks = (kTileW / tsx + 2) * (kTileH / tsy + 2);
        if (quad && tile_blocks <= kTileMaxBlocks && tile_blocks * D.R2 <= kTileMaxEntries) {
            dim3 blkt(kTileW / 4, kTileH);
            dim3 gridt((B.W + kTileW - 1) / kTileW, (B.H + kTileH - 1) / kTileH, n);
            if (sd.bps == 1)
                k_blend_tile<uint8_t><<<gridt, blkt, 0, st>>>(D, B, dev_src, cd.frame_bytes, dev_super, sd.frame_bytes, dev_dst, cd.frame_bytes, weights);
            else
                k_blend_tile<uint16_t><<<gridt, blkt, 0, st>>>(D, B, dev_src, cd.frame_bytes, dev_super, sd.frame_bytes, dev_dst, cd.frame_bytes, weights);
            ctx->launches++;
            continue;
        }
        if (quad) {
            dim3 blk4(32, 8);
            dim3 grid4(((B.W + 3) / 4 + blk4.x - 1) / blk4.x, (B.H + blk4.y - 1) / blk4.y, n);
            if (sd.bps == 1)
                k_blend4<uint8_t><<<grid4, blk4, 0, st>>>(D, B, dev_src, cd.frame_bytes, dev_super, sd.frame_bytes, dev_dst, cd.frame_bytes, weights);
            else
                k_blend4<uint16_t><<<grid4, blk4, 0, st>>>(D, B, dev_src, cd.frame_bytes, dev_super, sd.frame_bytes, dev_dst, cd.frame_bytes, weights);
            ctx->launches++;
            continue;
        }
        dim3 blk(32, 8);
        dim3 grid((B.W + blk.x - 1) / blk.x, (B.H + blk.y - 1) / blk.y, n);
        if (sd.bps == 1)
            k_blend<uint8_t><<<grid, blk, 0, st>>>(D, B, dev_src, cd.frame_bytes, dev_super, sd.frame_bytes, dev_dst, cd.frame_bytes, usable, weights);
        else
            k_blend<uint16_t><<<grid, blk, 0, st>>>(D, B, dev_src, cd.frame_bytes, dev_super, sd.frame_bytes, dev_dst, cd.frame_bytes, usable, weights);
        ctx->launches++;
    }
    MVB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

} // namespace mvb200
