// mv.Analyse on the GPU: the hierarchical predictor-driven block-matching search
// (gopSearchMVs GroupOfPlanes.c:69-125, doPobSearchMVs PlaneOfBlocks.cpp:971-1131,
// pobPseudoEPZSearch :819-968 and everything below it).
//
// Execution model.  One search (analyse instance x frame) is a strict serial chain:
// block k's predictors are the results of block k-1 (meander scan) and of the row
// above, `badcount` is a running counter and every accept is a strict '<' against
// the running minimum.  Bit-exactness forbids reordering, so
//   * ONE WARP walks one chain from the coarsest level to the finest, keeping the
//     chain state (best vector, running minimum cost, lambda, badcount, global MV)
//     in per-lane registers -- every lane computes the same scalars, so there is
//     no cross-lane protocol for control flow;
//   * the 32 lanes split the PIXELS of each candidate block: unaligned 32-bit
//     loads of source and reference words, VABSDIFF4.U8.ACC (`__vsadu4`) /
//     2x16-bit absolute differences, one REDUX.SUM per plane class;
//   * between levels the lanes split the BLOCKS: global-MV histogram
//     (pobEstimateGlobalMVDoubled :1559-1636) and the 2x vector-field upsampling
//     (pobInterpolatePrediction :1447-1514) are block-parallel;
//   * throughput comes from running thousands of chains at once: a batch of F
//     frames x 2*radius vector clips gives F*2*radius independent warps.
// The MVTools_vectors blob itself is the working memory: before a level is
// searched its VECTOR array holds the interpolated predictors, and each block's
// result overwrites its slot (exactly the role of pob->vectors in the reference).
//
// Host-identical scalar arithmetic: the few FP64 operations (lambda scaling
// :461-462, overlap-aware interpolation :1457,1500-1502) use IEEE double mul/div
// with round-to-nearest and C truncation; the file is compiled with -fmad=false.
#include "search_common.cuh"

#include <cmath>
#include <vector>

#include <cstdlib>

namespace mvb200 {

namespace {

struct PlaneView {
    const uint8_t *base; // sub-plane 0, padded origin
    long long pitch;     // bytes
    long long sub;       // bytes between pel sub-planes = pitch * padded_h
    int hpad, vpad, pw, ph;
};

struct State {
    // per level
    int nBlkX, nBlkY, blkW, blkH, ovX, ovY, logPel, logScale, logxR, logyR, bps, chroma, smallest;
    int cw, ch;
    PlaneView src[3], ref[3];
    VEC *vectors;
    int searchType, searchParam;
    int penaltyZero, pglobal, penaltyNew, badrange, tryMany, dctmode, dctweight16;
    long long badSAD, LSAD, verybigSAD;
    VEC globalMV, zeroMVfs;
    // per block
    int x[3], y[3], blkx, blky, blkIdx, scanDir;
    int dxMin, dxMax, dyMin, dyMax;
    long long nLambda, nMinCost;
    VEC bestMV, predictor, predictors[4];
    const uint8_t *pSrc[3];
    int srcLuma, badcount;
    int lane;
    // dct 1..4
    const float *dcx, *dcy; // dcx[k * blkW + j] = (float)(2 cos(pi (2j + 1) k / (2 blkW)))
    float *dtmp;            // blkW * blkH row-pass results
    int *dsrc;              // the source block's DCT image
    int dctshift, bits;
};

// ------------------------------------------------------------------ pixels

__device__ __forceinline__ unsigned ldu32(const uint8_t *p) {
    unsigned long long a = (unsigned long long)p;
    const unsigned *q = (const unsigned *)(a & ~3ull);
    unsigned sh = ((unsigned)a & 3u) * 8u;
    unsigned lo = __ldg(q);
    if (sh == 0)
        return lo;
    unsigned hi = __ldg(q + 1);
    return __funnelshift_r(lo, hi, sh);
}

__device__ __forceinline__ unsigned absdiff_word(unsigned a, unsigned b, int bps) {
    return bps == 1 ? __vsadu4(a, b) : __vsadu2(a, b);
}

// per-lane partial SAD over a (rowbytes x h) block; caller reduces.   sad_c, SADFunctions.cpp:354-367
__device__ __forceinline__ unsigned lane_sad(const uint8_t *a, long long ap, const uint8_t *b, long long bp, int rowbytes, int h, int bps, int lane) {
    // rows shorter than a word (2-pixel 8-bit chroma rows of 4x4 blocks): mask the word
    const unsigned mask = rowbytes >= 4 ? 0xffffffffu : ((1u << (8 * rowbytes)) - 1u);
    int wpr = max(rowbytes >> 2, 1);
    int lw = 31 - __clz(wpr);
    int total = wpr * h;
    unsigned acc = 0;
    for (int i = lane; i < total; i += 32) {
        int row = i >> lw, col = (i & (wpr - 1)) << 2;
        acc += absdiff_word(ldu32(a + row * ap + col) & mask, ldu32(b + row * bp + col) & mask, bps);
    }
    return acc;
}

__device__ __forceinline__ unsigned lane_luma(const uint8_t *a, long long ap, int rowbytes, int h, int bps, int lane) {
    const unsigned mask = rowbytes >= 4 ? 0xffffffffu : ((1u << (8 * rowbytes)) - 1u);
    int wpr = max(rowbytes >> 2, 1);
    int lw = 31 - __clz(wpr);
    int total = wpr * h;
    unsigned acc = 0;
    for (int i = lane; i < total; i += 32) {
        int row = i >> lw, col = (i & (wpr - 1)) << 2;
        unsigned v = ldu32(a + row * ap + col) & mask;
        acc += bps == 1 ? __dp4a(v, 0x01010101u, 0u) : ((v & 0xffffu) + (v >> 16));
    }
    return acc;
}

__device__ __forceinline__ void hadamard4(int &d0, int &d1, int &d2, int &d3, int s0, int s1, int s2, int s3) {
    int t0 = s0 + s1, t1 = s0 - s1, t2 = s2 + s3, t3 = s2 - s3;
    d0 = t0 + t2;
    d2 = t0 - t2;
    d1 = t1 + t3;
    d3 = t1 - t3;
}

// per-lane partial SATD, NOT yet halved.  Satd_C, SADFunctions.cpp:605-710: every 8x4 (or lone 4x4) tile's sum of absolute
// Hadamard coefficients is halved; each 4x4 half is even, so the block's SATD is (sum over all tiles) >> 1 -- the callers halve the
// warp sum.  One lane per tile ROW (four adjacent lanes per 4x4 tile: 16 lanes busy on an 8x8 block, where one lane per tile kept
// 4 busy): a lane transforms its row horizontally, the vertical transform is two butterfly exchanges between the tile's four lanes.
// Called by all 32 lanes (full-mask shuffles).
__device__ unsigned lane_satd(const uint8_t *a, long long ap, const uint8_t *b, long long bp, int w, int h, int bps, int lane) {
    const int tx = w >> 2, units = tx * (h >> 2) * 4;
    unsigned acc = 0;
    for (int u0 = 0; u0 < units; u0 += 32) {
        const int u = u0 + lane;
        const bool live = u < units;
        const int t = (live ? u : 0) >> 2, r = u & 3;
        const int ox = (t % tx) * 4, oy = (t / tx) * 4 + r;
        int m0 = 0, m1 = 0, m2 = 0, m3 = 0;
        if (live) {
            const uint8_t *pa = a + oy * ap + ox * bps, *pb = b + oy * bp + ox * bps;
            int d0, d1, d2, d3;
            if (bps == 1) {
                unsigned va = ldu32(pa), vb = ldu32(pb);
                d0 = (int)(va & 255) - (int)(vb & 255);
                d1 = (int)((va >> 8) & 255) - (int)((vb >> 8) & 255);
                d2 = (int)((va >> 16) & 255) - (int)((vb >> 16) & 255);
                d3 = (int)(va >> 24) - (int)(vb >> 24);
            } else {
                unsigned va = ldu32(pa), vb = ldu32(pb), va2 = ldu32(pa + 4), vb2 = ldu32(pb + 4);
                d0 = (int)(va & 0xffff) - (int)(vb & 0xffff);
                d1 = (int)(va >> 16) - (int)(vb >> 16);
                d2 = (int)(va2 & 0xffff) - (int)(vb2 & 0xffff);
                d3 = (int)(va2 >> 16) - (int)(vb2 >> 16);
            }
            hadamard4(m0, m1, m2, m3, d0, d1, d2, d3);
        }
        // vertical HADAMARD4 over the tile's rows r = 0..3 (lanes 4k .. 4k + 3): rows (0,1) and (2,3) first, then the pairs
        int v[4] = {m0, m1, m2, m3};
#pragma unroll
        for (int c = 0; c < 4; c++) {
            int o = __shfl_xor_sync(0xffffffffu, v[c], 1);
            v[c] = (r & 1) ? o - v[c] : v[c] + o; // even row: t0 + t1 (t2 + t3), odd row: t0 - t1 (t2 - t3)
            o = __shfl_xor_sync(0xffffffffu, v[c], 2);
            v[c] = (r & 2) ? o - v[c] : v[c] + o; // rows 0, 1: sums, rows 2, 3: differences of the pairs
            acc += (unsigned)abs(v[c]);
        }
    }
    return acc;
}

__device__ __forceinline__ unsigned wsum(unsigned v) { return __reduce_add_sync(0xffffffffu, v); }

// mvpGetAbsolutePointerPel1/2/4, MVFrame.cpp:1706-1730
__device__ __forceinline__ const uint8_t *pel_ptr(const PlaneView &v, int bps, int logpel, int x, int y) {
    int m = (1 << logpel) - 1;
    int idx = (x & m) | ((y & m) << logpel);
    return v.base + idx * v.sub + (long long)(y >> logpel) * v.pitch + (long long)(x >> logpel) * bps;
}

__device__ __forceinline__ const uint8_t *ref_luma(const State &S, int vx, int vy) { // pobGetRefBlock :35-53
    return pel_ptr(S.ref[0], S.bps, S.logPel, (S.x[0] << S.logPel) + vx, (S.y[0] << S.logPel) + vy);
}

__device__ __forceinline__ const uint8_t *ref_chroma(const State &S, int p, int vx, int vy) { // pobGetRefBlockU/V :56-101
    int xbias = (vx < 0) * ((1 << S.logxR) - 1);
    int ybias = (vy < 0) * ((1 << S.logyR) - 1);
    return pel_ptr(S.ref[p], S.bps, S.logPel, (S.x[p] << S.logPel) + ((vx + xbias) >> S.logxR), (S.y[p] << S.logPel) + ((vy + ybias) >> S.logyR));
}

// dctBytes2D, DCTFFTW.cpp:200-208, by one warp: Pixels2Float, FFTW's REDFT10 x REDFT10 evaluated from its definition
// (unnormalised DCT-II, rows then columns, float dot products in index order, no FMA -- the same arithmetic as the
// oracle and as the FFTW stand-in the reference is built with here; against the real FFTW: float rounding, "parity
// unpinned"), Float2Pixels_C (:31-54).  is_src: the DCT image is stored (S.dsrc); otherwise it is compared
// with the stored one on the fly: returns this lane's part of SAD(dctSrc, dctRef), and |DC difference| in dc (lane 0).
__device__ unsigned dct_block(State &S, const uint8_t *p, long long pitch, bool is_src, unsigned &dc) {
    const int nx = S.blkW, ny = S.blkH, n = nx * ny, lx = 31 - __clz(nx), ly = 31 - __clz(ny);
    for (int idx = S.lane; idx < n; idx += 32) { // rows: tmp[i][k] = sum_j x[i][j] cx[k][j]
        const int i = idx >> lx, k = idx & (nx - 1);
        const uint8_t *row = p + i * pitch;
        const float *c = S.dcx + k * nx;
        float acc = 0.f;
        for (int j = 0; j < nx; j++) {
            const float x = (float)(S.bps == 1 ? (int)row[j] : (int)((const uint16_t *)row)[j]);
            acc = __fadd_rn(acc, __fmul_rn(x, c[j]));
        }
        S.dtmp[idx] = acc;
    }
    __syncwarp();
    const int pixelMax = (1 << S.bits) - 1, pixelHalf = 1 << (S.bits - 1);
    unsigned sad = 0;
    dc = 0;
    for (int idx = S.lane; idx < n; idx += 32) { // columns: y[k][j] = sum_i tmp[i][j] cy[k][i]
        const int k = idx >> lx, j = idx & (nx - 1);
        const float *c = S.dcy + k * ny;
        float acc = 0.f;
        for (int i = 0; i < ny; i++)
            acc = __fadd_rn(acc, __fmul_rn(S.dtmp[(i << lx) + j], c[i]));
        int v;
        if (idx == 0) // DC (:50-53)
            v = min(pixelMax, max(0, (__float2int_rn(__fmul_rn(acc, 0.5f)) >> (S.dctshift + 2)) + pixelHalf));
        else
            v = min(pixelMax, max(0, (__float2int_rn(__fmul_rn(acc, 0.70710678118654752440084436210485f)) >> S.dctshift) + pixelHalf));
        if (is_src) {
            S.dsrc[idx] = v;
        } else {
            const unsigned d = (unsigned)abs(S.dsrc[idx] - v);
            sad += d;
            if (idx == 0)
                dc = d;
        }
    }
    (void)ly;
    __syncwarp();
    return sad;
}

// pobLumaSAD<dctmode>, PlaneOfBlocks.cpp:117-203
__device__ long long luma_sad(State &S, const uint8_t *pRef) {
    long long sp = S.src[0].pitch, rp = S.ref[0].pitch;
    int d = S.dctmode;
    if (d == 5)
        return wsum(lane_satd(S.pSrc[0], sp, pRef, rp, S.blkW, S.blkH, S.bps, S.lane)) >> 1;
    long long sad = wsum(lane_sad(S.pSrc[0], sp, pRef, rp, S.blkW * S.bps, S.blkH, S.bps, S.lane));
    if (d == 0)
        return sad;
    if (d <= 4) { // the FFTW DCT modes (:123-163)
        if (d == 1 || (d == 2 && S.dctweight16 > 0)) {
            unsigned dc;
            unsigned part = dct_block(S, pRef, rp, false, dc);
            const unsigned tot = wsum(part) + 3u * wsum(dc);
            const long long dctsad = (long long)tot * (long long)S.blkW / 2;
            if (d == 1)
                return dctsad;
            return (sad * (16 - S.dctweight16) + dctsad * S.dctweight16) / 16;
        }
        if (d == 2)
            return sad;
        int refLuma = (int)wsum(lane_luma(pRef, rp, S.blkW * S.bps, S.blkH, S.bps, S.lane));
        if (abs(S.srcLuma - refLuma) > ((S.srcLuma + refLuma) >> 5)) {
            unsigned dc;
            const long long dctsad = (long long)wsum(dct_block(S, pRef, rp, false, dc)) * (long long)S.blkW / 2;
            if (d == 3)
                sad = sad / 2 + dctsad / 2;
            else
                sad = sad / 4 + dctsad / 2 + dctsad / 4;
        }
        return sad;
    }
    if (d == 6) {
        if (S.dctweight16 > 0) {
            long long dctsad = wsum(lane_satd(S.pSrc[0], sp, pRef, rp, S.blkW, S.blkH, S.bps, S.lane)) >> 1;
            sad = (sad * (16 - S.dctweight16) + dctsad * S.dctweight16) / 16;
        }
        return sad;
    }
    if (d == 9) {
        if (S.dctweight16 > 1) {
            int half = S.dctweight16 / 2;
            long long dctsad = wsum(lane_satd(S.pSrc[0], sp, pRef, rp, S.blkW, S.blkH, S.bps, S.lane)) >> 1;
            sad = (sad * (16 - half) + dctsad * half) / 16;
        }
        return sad;
    }
    int refLuma = (int)wsum(lane_luma(pRef, rp, S.blkW * S.bps, S.blkH, S.bps, S.lane));
    int shift = d == 10 ? 4 : 5;
    if (abs(S.srcLuma - refLuma) > ((S.srcLuma + refLuma) >> shift)) {
        long long dctsad = wsum(lane_satd(S.pSrc[0], sp, pRef, rp, S.blkW, S.blkH, S.bps, S.lane)) >> 1;
        if (d == 7)
            sad = sad / 2 + dctsad / 2;
        else if (d == 8)
            sad = sad / 4 + dctsad / 2 + dctsad / 4;
        else
            sad = sad / 2 + dctsad / 4 + sad / 4;
    }
    return sad;
}

__device__ __forceinline__ long long chroma_sad(const State &S, int vx, int vy) {
    unsigned acc = lane_sad(S.pSrc[1], S.src[1].pitch, ref_chroma(S, 1, vx, vy), S.ref[1].pitch, S.cw * S.bps, S.ch, S.bps, S.lane);
    acc += lane_sad(S.pSrc[2], S.src[2].pitch, ref_chroma(S, 2, vx, vy), S.ref[2].pitch, S.cw * S.bps, S.ch, S.bps, S.lane);
    return wsum(acc);
}

// ------------------------------------------------------------------ search

#define F_PENALTYNEW 2
#define F_UPDATEDIR 4
#define F_UPDATEBEST 8

// pobCheckMV_Template, PlaneOfBlocks.cpp:219-261 (early outs kept: they are warp-uniform)
__device__ __noinline__ void check_mv(State &S, int vx, int vy, int flags, int *dir, int val) {
    if (!(vx >= S.dxMin && vy >= S.dyMin && vx < S.dxMax && vy < S.dyMax))
        return;
    int ddx = S.predictor.x - vx, ddy = S.predictor.y - vy;
    int dist = (int)((unsigned)(ddx * ddx) + (unsigned)(ddy * ddy));
    long long cost = (int)((S.nLambda * dist) >> 8); // pobMotionDistorsion :111-114
    if (cost >= S.nMinCost)
        return;
    long long sad = luma_sad(S, ref_luma(S, vx, vy));
    cost += sad + ((flags & F_PENALTYNEW) ? ((S.penaltyNew * sad) >> 8) : 0);
    if (cost >= S.nMinCost)
        return;
    long long saduv = 0;
    if (S.chroma) {
        saduv = chroma_sad(S, vx, vy);
        cost += saduv + ((flags & F_PENALTYNEW) ? ((S.penaltyNew * saduv) >> 8) : 0);
        if (cost >= S.nMinCost)
            return;
    }
    if (flags & F_UPDATEBEST) {
        S.bestMV.x = vx;
        S.bestMV.y = vy;
    }
    S.nMinCost = cost;
    S.bestMV.sad = sad + saduv;
    if (flags & F_UPDATEDIR)
        *dir = val;
}

__device__ __forceinline__ void CheckMV0(State &S, int vx, int vy) { check_mv(S, vx, vy, F_UPDATEBEST, nullptr, 0); }
__device__ __forceinline__ void CheckMV(State &S, int vx, int vy) { check_mv(S, vx, vy, F_PENALTYNEW | F_UPDATEBEST, nullptr, 0); }
__device__ __forceinline__ void CheckMV2(State &S, int vx, int vy, int *dir, int val) { check_mv(S, vx, vy, F_PENALTYNEW | F_UPDATEDIR | F_UPDATEBEST, dir, val); }
__device__ __forceinline__ void CheckMVdir(State &S, int vx, int vy, int *dir, int val) { check_mv(S, vx, vy, F_PENALTYNEW | F_UPDATEDIR, dir, val); }

__device__ __forceinline__ int clipx(const State &S, int v) { return min(max(v, S.dxMin), S.dxMax - 1); }
__device__ __forceinline__ int clipy(const State &S, int v) { return min(max(v, S.dyMin), S.dyMax - 1); }
__device__ __forceinline__ VEC clipmv(const State &S, VEC v) { // pobClipMV :293-311
    VEC r;
    r.x = clipx(S, v.x);
    r.y = clipy(S, v.y);
    r.sad = v.sad;
    return r;
}
__device__ __forceinline__ int median3(int a, int b, int c) { return max(min(a, b), min(max(a, b), c)); }
__device__ __forceinline__ bool vec_ok(const State &S, int vx, int vy) { return vx >= S.dxMin && vy >= S.dyMin && vx < S.dxMax && vy < S.dyMax; }

// pobFetchPredictors, PlaneOfBlocks.cpp:419-463
__device__ void fetch_predictors(State &S) {
    bool hasNext = (S.scanDir == 1 && S.blkx < S.nBlkX - 1) || (S.scanDir == -1 && S.blkx > 0);
    if ((S.scanDir == 1 && S.blkx > 0) || (S.scanDir == -1 && S.blkx < S.nBlkX - 1))
        S.predictors[1] = clipmv(S, ldvec(&S.vectors[S.blkIdx - S.scanDir]));
    else
        S.predictors[1] = clipmv(S, S.zeroMVfs);
    if (S.blky > 0)
        S.predictors[2] = clipmv(S, ldvec(&S.vectors[S.blkIdx - S.nBlkX]));
    else
        S.predictors[2] = clipmv(S, S.zeroMVfs);
    if (S.blky < S.nBlkY - 1 && hasNext)
        S.predictors[3] = clipmv(S, ldvec(&S.vectors[S.blkIdx + S.nBlkX + S.scanDir]));
    else if (S.blky > 0 && hasNext)
        S.predictors[3] = clipmv(S, ldvec(&S.vectors[S.blkIdx - S.nBlkX + S.scanDir]));
    else
        S.predictors[3] = clipmv(S, S.zeroMVfs);
    if (S.blky > 0) {
        S.predictors[0].x = median3(S.predictors[1].x, S.predictors[2].x, S.predictors[3].x);
        S.predictors[0].y = median3(S.predictors[1].y, S.predictors[2].y, S.predictors[3].y);
        S.predictors[0].sad = max(S.predictors[1].sad, max(S.predictors[2].sad, S.predictors[3].sad));
    } else {
        S.predictors[0] = S.predictors[1];
    }
    if (S.smallest)
        S.predictor = S.predictors[0];
    double scale = __ddiv_rn((double)S.LSAD, (double)(S.LSAD + (S.predictor.sad >> 1)));
    S.nLambda = __double2ll_rz(__dmul_rn(__dmul_rn((double)S.nLambda, scale), scale));
}

__device__ void nstep_search(State &S, int stp) { // :466-485
    for (int length = stp; length > 0; length--) {
        int dx = S.bestMV.x, dy = S.bestMV.y;
        CheckMV(S, dx + length, dy + length);
        CheckMV(S, dx + length, dy);
        CheckMV(S, dx + length, dy - length);
        CheckMV(S, dx, dy - length);
        CheckMV(S, dx, dy + length);
        CheckMV(S, dx - length, dy + length);
        CheckMV(S, dx - length, dy);
        CheckMV(S, dx - length, dy - length);
    }
}

__device__ void onetime_search(State &S, int length) { // :488-527
    int direction = 0;
    int dx = S.bestMV.x, dy = S.bestMV.y;
    CheckMV2(S, dx - length, dy, &direction, 2);
    CheckMV2(S, dx + length, dy, &direction, 1);
    if (direction == 1) {
        while (direction) {
            direction = 0;
            dx += length;
            CheckMV2(S, dx + length, dy, &direction, 1);
        }
    } else if (direction == 2) {
        while (direction) {
            direction = 0;
            dx -= length;
            CheckMV2(S, dx - length, dy, &direction, 1);
        }
    }
    CheckMV2(S, dx, dy - length, &direction, 2);
    CheckMV2(S, dx, dy + length, &direction, 1);
    if (direction == 1) {
        while (direction) {
            direction = 0;
            dy += length;
            CheckMV2(S, dx, dy + length, &direction, 1);
        }
    } else if (direction == 2) {
        while (direction) {
            direction = 0;
            dy -= length;
            CheckMV2(S, dx, dy - length, &direction, 1);
        }
    }
}

__device__ void diamond_search(State &S, int length) { // :530-632
    enum { R = 1, L = 2, D = 4, U = 8 };
    int direction = 15, last;
    while (direction > 0) {
        int dx = S.bestMV.x, dy = S.bestMV.y;
        last = direction;
        direction = 0;
        if (last & R) CheckMV2(S, dx + length, dy, &direction, R);
        if (last & L) CheckMV2(S, dx - length, dy, &direction, L);
        if (last & D) CheckMV2(S, dx, dy + length, &direction, D);
        if (last & U) CheckMV2(S, dx, dy - length, &direction, U);
        if (direction) {
            last = direction;
            dx = S.bestMV.x;
            dy = S.bestMV.y;
            if (last & (R + L)) {
                CheckMV2(S, dx, dy + length, &direction, D);
                CheckMV2(S, dx, dy - length, &direction, U);
            } else {
                CheckMV2(S, dx + length, dy, &direction, R);
                CheckMV2(S, dx - length, dy, &direction, L);
            }
        } else {
            // the diagonal probes of every case, as (sx, sy, value) triples in the reference's order
            int sx[4], sy[4], vv[4], cnt = 0;
#define DG(a, b, c) (sx[cnt] = (a), sy[cnt] = (b), vv[cnt] = (c), cnt++)
            switch (last) {
            case R: DG(1, 1, R + D); DG(1, -1, R + U); break;
            case L: DG(-1, 1, L + D); DG(-1, -1, L + U); break;
            case D: DG(1, 1, R + D); DG(-1, 1, L + D); break;
            case U: DG(1, -1, R + U); DG(-1, -1, L + U); break;
            case R + D: DG(1, 1, R + D); DG(-1, 1, L + D); DG(1, -1, R + U); break;
            case L + D: DG(1, 1, R + D); DG(-1, 1, L + D); DG(-1, -1, L + U); break;
            case R + U: DG(1, 1, R + D); DG(-1, -1, L + U); DG(1, -1, R + U); break;
            case L + U: DG(-1, -1, L + U); DG(-1, 1, L + D); DG(1, -1, R + U); break;
            default: DG(1, 1, R + D); DG(-1, 1, L + D); DG(1, -1, R + U); DG(-1, -1, L + U); break;
            }
#undef DG
            for (int k = 0; k < cnt; k++)
                CheckMV2(S, dx + sx[k] * length, dy + sy[k] * length, &direction, vv[k]);
        }
    }
}

__device__ void expanding_search(State &S, int r, int s, int mvx, int mvy) { // :635-658
    for (int i = -r + s; i < r; i += s) {
        CheckMV(S, mvx + i, mvy - r);
        CheckMV(S, mvx + i, mvy + r);
    }
    for (int j = -r + s; j < r; j += s) {
        CheckMV(S, mvx - r, mvy + j);
        CheckMV(S, mvx + r, mvy + j);
    }
    CheckMV(S, mvx - r, mvy - r);
    CheckMV(S, mvx - r, mvy + r);
    CheckMV(S, mvx + r, mvy - r);
    CheckMV(S, mvx + r, mvy + r);
}

__constant__ int c_mod6m1[8] = {5, 0, 1, 2, 3, 4, 5, 0};
__constant__ int c_hex2[8][2] = {{-1, -2}, {-2, 0}, {-1, 2}, {1, 2}, {2, 0}, {1, -2}, {-1, -2}, {-2, 0}};
__constant__ int c_hex4[16][2] = {{-4, 2}, {-4, 1}, {-4, 0}, {-4, -1}, {-4, -2}, {4, -2}, {4, -1}, {4, 0},
                                  {4, 1},  {4, 2},  {2, 3},  {0, 4},   {-2, 3},  {-2, -3}, {0, -4}, {2, -3}};

__device__ void hex2_search(State &S, int range) { // :661-724
    int dir = -2;
    int bmx = S.bestMV.x, bmy = S.bestMV.y;
    if (range > 1) {
        CheckMVdir(S, bmx - 2, bmy, &dir, 0);
        CheckMVdir(S, bmx - 1, bmy + 2, &dir, 1);
        CheckMVdir(S, bmx + 1, bmy + 2, &dir, 2);
        CheckMVdir(S, bmx + 2, bmy, &dir, 3);
        CheckMVdir(S, bmx + 1, bmy - 2, &dir, 4);
        CheckMVdir(S, bmx - 1, bmy - 2, &dir, 5);
        if (dir != -2) {
            bmx += c_hex2[dir + 1][0];
            bmy += c_hex2[dir + 1][1];
            for (int i = 1; i < range / 2 && vec_ok(S, bmx, bmy); i++) {
                const int odir = c_mod6m1[dir + 1];
                dir = -2;
                CheckMVdir(S, bmx + c_hex2[odir + 0][0], bmy + c_hex2[odir + 0][1], &dir, odir - 1);
                CheckMVdir(S, bmx + c_hex2[odir + 1][0], bmy + c_hex2[odir + 1][1], &dir, odir);
                CheckMVdir(S, bmx + c_hex2[odir + 2][0], bmy + c_hex2[odir + 2][1], &dir, odir + 1);
                if (dir == -2)
                    break;
                bmx += c_hex2[dir + 1][0];
                bmy += c_hex2[dir + 1][1];
            }
        }
        S.bestMV.x = bmx;
        S.bestMV.y = bmy;
    }
    expanding_search(S, 1, 1, bmx, bmy);
}

__device__ void umh_search(State &S, int range, int omx, int omy) { // :727-769
    for (int i = 1; i < range; i += 2) {
        CheckMV(S, omx - i, omy);
        CheckMV(S, omx + i, omy);
    }
    for (int j = 1; j < range; j += 2) {
        CheckMV(S, omx, omy - j);
        CheckMV(S, omx, omy + j);
    }
    int i = 1;
    do {
        for (int j = 0; j < 16; j++)
            CheckMV(S, omx + c_hex4[j][0] * i, omy + c_hex4[j][1] * i);
    } while (++i <= range / 4);
    hex2_search(S, range);
}

__device__ void refine(State &S) { // pobRefine :772-816
    switch (S.searchType) {
    case 0:
        for (int i = S.searchParam; i > 0; i /= 2)
            onetime_search(S, i);
        break;
    case 1:
        nstep_search(S, S.searchParam);
        break;
    case 2:
        for (int i = S.searchParam; i > 0; i /= 2)
            diamond_search(S, i);
        break;
    case 3: {
        int mvx = S.bestMV.x, mvy = S.bestMV.y;
        for (int i = 1; i <= S.searchParam; i++)
            expanding_search(S, i, 1, mvx, mvy);
        break;
    }
    case 4:
        hex2_search(S, S.searchParam);
        break;
    case 5:
        umh_search(S, S.searchParam, S.bestMV.x, S.bestMV.y);
        break;
    case 6: {
        int mvx = S.bestMV.x, mvy = S.bestMV.y;
        for (int i = 1; i <= S.searchParam; i++) {
            CheckMV(S, mvx - i, mvy);
            CheckMV(S, mvx + i, mvy);
        }
        break;
    }
    case 7: {
        int mvx = S.bestMV.x, mvy = S.bestMV.y;
        for (int i = 1; i <= S.searchParam; i++) {
            CheckMV(S, mvx, mvy - i);
            CheckMV(S, mvx, mvy + i);
        }
        break;
    }
    }
}

__device__ long long full_sad_at(State &S, int lx, int ly, int cx, int cy) {
    long long sad = luma_sad(S, ref_luma(S, lx, ly));
    if (S.chroma)
        sad += chroma_sad(S, cx, cy);
    return sad;
}

// pobPseudoEPZSearch, PlaneOfBlocks.cpp:819-968
__device__ void epz_search(State &S) {
    fetch_predictors(S);
    if (S.dctmode >= 1 && S.dctmode <= 4) { // make dct of source block (:824-828 / :1336-1338)
        unsigned dc;
        dct_block(S, S.pSrc[0], S.src[0].pitch, true, dc);
    }
    if (S.dctmode >= 3)
        S.srcLuma = (int)wsum(lane_luma(S.pSrc[0], S.src[0].pitch, S.blkW * S.bps, S.blkH, S.bps, S.lane));

    S.bestMV.x = S.zeroMVfs.x;
    S.bestMV.y = S.zeroMVfs.y;
    long long sad = full_sad_at(S, 0, S.zeroMVfs.y, 0, 0);
    S.bestMV.sad = sad;
    S.nMinCost = sad + ((S.penaltyZero * sad) >> 8);

    VEC bestMany[7];
    long long costMany[7];
    if (S.tryMany) {
        for (int i = 0; i < 7; i++)
            costMany[i] = 0;
        refine(S);
        bestMany[0] = S.bestMV;
        costMany[0] = S.nMinCost;
    }

    S.globalMV = clipmv(S, S.globalMV);
    sad = full_sad_at(S, S.globalMV.x, S.globalMV.y, S.globalMV.x, S.globalMV.y);
    long long cost = sad + ((S.pglobal * sad) >> 8);
    if (cost < S.nMinCost || S.tryMany) {
        S.bestMV.x = S.globalMV.x;
        S.bestMV.y = S.globalMV.y;
        S.bestMV.sad = sad;
        S.nMinCost = cost;
    }
    if (S.tryMany) {
        refine(S);
        bestMany[1] = S.bestMV;
        costMany[1] = S.nMinCost;
    }

    sad = full_sad_at(S, S.predictor.x, S.predictor.y, S.predictor.x, S.predictor.y);
    cost = sad;
    if (cost < S.nMinCost || S.tryMany) {
        S.bestMV.x = S.predictor.x;
        S.bestMV.y = S.predictor.y;
        S.bestMV.sad = sad;
        S.nMinCost = cost;
    }
    if (S.tryMany) {
        refine(S);
        bestMany[2] = S.bestMV;
        costMany[2] = S.nMinCost;
    }

    for (int i = 0; i < 4; i++) {
        if (S.tryMany)
            S.nMinCost = S.verybigSAD + 1;
        CheckMV0(S, S.predictors[i].x, S.predictors[i].y);
        if (S.tryMany) {
            refine(S);
            bestMany[i + 3] = S.bestMV;
            costMany[i + 3] = S.nMinCost;
        }
    }

    if (S.tryMany) {
        S.nMinCost = S.verybigSAD + 1;
        for (int i = 0; i < 7; i++)
            if (costMany[i] < S.nMinCost) {
                S.bestMV = bestMany[i];
                S.nMinCost = costMany[i];
            }
    } else {
        refine(S);
    }

    long long foundSAD = S.bestMV.sad;
    if (S.blkIdx > 1 && foundSAD > (S.badSAD + S.badSAD * S.badcount / 16)) {
        S.badcount++;
        if (S.badrange > 0) {
            umh_search(S, S.badrange * (1 << S.logPel), 0, 0);
        } else if (S.badrange < 0) {
            for (int i = 1; i < -S.badrange * (1 << S.logPel); i += (1 << S.logPel)) {
                expanding_search(S, i, 1 << S.logPel, 0, 0);
                if (S.bestMV.sad < foundSAD / 4)
                    break;
            }
        }
        int mvx = S.bestMV.x, mvy = S.bestMV.y;
        for (int i = 1; i < (1 << S.logPel); i++)
            expanding_search(S, i, 1, mvx, mvy);
    }
}

__device__ void bind_planes(PlaneView v[3], const uint8_t *frame, const SuperDesc &sd, int level, int nplanes_used) {
    for (int p = 0; p < 3; p++) {
        v[p].base = nullptr;
        if (p < nplanes_used) {
            const mvb200_level_geom &g = sd.lv[p][level];
            v[p].pitch = sd.pitch[p];
            v[p].base = frame + sd.offset[p] + (long long)g.row0 * sd.pitch[p];
            v[p].sub = sd.pitch[p] * g.padded_h;
            v[p].hpad = g.hpad;
            v[p].vpad = g.vpad;
            v[p].pw = g.padded_w;
            v[p].ph = g.padded_h;
        }
    }
}

// doPobSearchMVs, PlaneOfBlocks.cpp:971-1131
__device__ void search_level(State &S, const SearchParams &P, int st, int stp, const VEC &globalMVec, int fieldShift, int *pmeanLumaChange, int pglobal, int tryMany) {
    S.dctweight16 = min(16, abs(*pmeanLumaChange) / (S.blkW * S.blkH));
    S.badSAD = P.badSAD;
    S.badrange = P.badrange;
    S.zeroMVfs.x = 0;
    S.zeroMVfs.y = fieldShift;
    S.zeroMVfs.sad = 0;
    S.globalMV.x = (1 << S.logPel) * globalMVec.x;
    S.globalMV.y = (1 << S.logPel) * globalMVec.y + fieldShift;
    S.globalMV.sad = globalMVec.sad;
    S.searchType = st;
    S.searchParam = stp;
    int nScale = 1 << S.logScale;
    int nLambdaLevel = P.nLambda / ((1 << S.logPel) * (1 << S.logPel));
    if (P.plevel == 1)
        nLambdaLevel = nLambdaLevel * nScale;
    else if (P.plevel == 2)
        nLambdaLevel = nLambdaLevel * nScale * nScale;
    S.penaltyZero = P.pzero;
    S.pglobal = pglobal;
    S.badcount = 0;
    S.tryMany = tryMany;
    S.penaltyNew = P.pnew;
    S.LSAD = P.lsad;
    int sumLumaChange = 0;
    const int stepX = S.blkW - S.ovX, stepY = S.blkH - S.ovY;
    const PlaneView &g0 = S.src[0];
    const int hps = g0.hpad >> S.logScale, vps = g0.vpad >> S.logScale;
    for (int p = 0; p < 3; p++)
        S.y[p] = S.src[p].base ? S.src[p].vpad : 0;
    for (S.blky = 0; S.blky < S.nBlkY; S.blky++) {
        S.scanDir = (S.blky % 2 == 0 || P.meander == 0) ? 1 : -1;
        int blkxStart = (S.scanDir == 1) ? 0 : S.nBlkX - 1;
        for (int iblkx = 0; iblkx < S.nBlkX; iblkx++) {
            S.blkx = blkxStart + iblkx * S.scanDir;
            S.blkIdx = S.blky * S.nBlkX + S.blkx;
            S.x[0] = g0.hpad + stepX * S.blkx;
            if (S.chroma) {
                S.x[1] = S.src[1].hpad + (stepX >> S.logxR) * S.blkx;
                S.x[2] = S.src[2].hpad + (stepX >> S.logxR) * S.blkx;
            }
            for (int p = 0; p < (S.chroma ? 3 : 1); p++)
                S.pSrc[p] = S.src[p].base + (long long)S.y[p] * S.src[p].pitch + (long long)S.x[p] * S.bps;
            S.nLambda = S.blky == 0 ? 0 : nLambdaLevel;
            S.dxMax = (g0.pw - S.x[0] - S.blkW - g0.hpad + hps) << S.logPel;
            S.dyMax = (g0.ph - S.y[0] - S.blkH - g0.vpad + vps) << S.logPel;
            S.dxMin = -((S.x[0] - g0.hpad + hps) << S.logPel);
            S.dyMin = -((S.y[0] - g0.vpad + vps) << S.logPel);
            S.predictor = clipmv(S, ldvec(&S.vectors[S.blkIdx]));
            epz_search(S);
            __syncwarp();
            if (S.lane == 0)
                stvec(&S.vectors[S.blkIdx], S.bestMV);
            __syncwarp();
            if (S.smallest)
                sumLumaChange += (int)wsum(lane_luma(ref_luma(S, 0, 0), S.ref[0].pitch, S.blkW * S.bps, S.blkH, S.bps, S.lane)) -
                                 (int)wsum(lane_luma(S.pSrc[0], S.src[0].pitch, S.blkW * S.bps, S.blkH, S.bps, S.lane));
        }
        S.y[0] += stepY;
        if (S.src[1].base) S.y[1] += stepY >> S.logyR;
        if (S.src[2].base) S.y[2] += stepY >> S.logyR;
    }
    if (S.smallest)
        *pmeanLumaChange = sumLumaChange / (S.nBlkX * S.nBlkY);
}

__global__ void __launch_bounds__(32) k_search(const SuperDesc *__restrict__ sdp, const SearchParams *__restrict__ params, const SearchJob *__restrict__ jobs, int njobs, int *__restrict__ hist_all, int hist_ints) {
    int chain = blockIdx.x;
    if (chain >= njobs)
        return;
    const int lane = threadIdx.x;
    const SearchJob job = jobs[chain];
    const SearchParams &P = params[job.param];
    const SuperDesc &sd = *sdp;
    const int L = P.levels;
    chain_prologue(job, P, lane);
    if (!job.ref)
        return;

    int *hist = hist_all + (long long)chain * hist_ints;
    const int nplanes_used = P.chroma ? 3 : 1;
    VEC globalMV;
    globalMV.x = 0;
    globalMV.y = 0;
    globalMV.sad = -1;
    const int pglobal = P.global ? P.pglobal : P.pzero;
    int meanLumaChange = 0;

    State S;
    S.lane = lane;
    S.blkW = P.blkW;
    S.blkH = P.blkH;
    S.ovX = P.ovX;
    S.ovY = P.ovY;
    S.logxR = sd.logxR;
    S.logyR = sd.logyR;
    S.bps = P.bps;
    S.chroma = P.chroma;
    S.cw = P.blkW >> sd.logxR;
    S.ch = P.blkH >> sd.logyR;
    S.dctmode = P.dctmode;
    S.verybigSAD = P.verybigSAD;
    S.srcLuma = 0;
    S.bits = P.bits;
    S.dcx = P.dct_cos;
    S.dcy = P.dct_cos ? P.dct_cos + P.blkW * P.blkW : nullptr;
    S.dtmp = (float *)(P.dct_scratch + (long long)chain * P.dct_scratch_stride);
    S.dsrc = (int *)(S.dtmp + P.blkW * P.blkH);
    S.dctshift = 0;
    while ((1 << S.dctshift) < P.blkW * P.blkH) // dctInit, DCTFFTW.cpp:142-149
        S.dctshift++;

    for (int l = L - 1; l >= 0; l--) {
        S.nBlkX = P.nBlkX[l];
        S.nBlkY = P.nBlkY[l];
        S.logPel = l == 0 ? P.logPel : 0;
        S.logScale = l;
        S.smallest = (l == L - 1);
        S.vectors = (VEC *)(job.out + P.level_off[l] + 4);
        bind_planes(S.src, job.src, sd, l, nplanes_used);
        bind_planes(S.ref, job.ref, sd, l, nplanes_used);
        int st, stp, fs, tm;
        if (l == L - 1) {
            st = (L == 1 || P.search == 6 || P.search == 7) ? P.search : P.search_coarse;
            stp = (L == 1) ? P.nPelSearch : P.nSearchParam;
            fs = (L - 1 == 0) ? job.field_shift : 0;
            tm = P.tryMany && L > 1;
        } else {
            st = (l == 0 || P.search == 6 || P.search == 7) ? P.search : P.search_coarse;
            stp = (l == 0) ? P.nPelSearch : P.nSearchParam;
            const VEC *coarse = (const VEC *)(job.out + P.level_off[l + 1] + 4);
            if (P.global)
                estimate_global_mv(coarse, P.nBlkX[l + 1] * P.nBlkY[l + 1], 8192 * 1 * 2, hist, globalMV, lane);
            interpolate_prediction(S.vectors, S.nBlkX, S.nBlkY, S.logPel, coarse, P.nBlkX[l + 1], P.nBlkY[l + 1], S.blkW, S.blkH, S.ovX, S.ovY, lane);
            __syncwarp();
            fs = (l == 0) ? job.field_shift : 0;
            tm = P.tryMany && l > 0;
        }
        search_level(S, P, st, stp, globalMV, fs, &meanLumaChange, pglobal, tm);
        __syncwarp();
    }
    chain_divide(job, P, lane);
}


// ------------------------------------------------------------------ Recalculate
// doPobRecalculateMVs, PlaneOfBlocks.cpp:1158-1424: the predictor of every block comes from the OLD
// vector field, there is no neighbour predictor and no badcount, so blocks are independent: one warp per
// block, grid = blocks x frames.  The refinement itself is the generic pattern code above.
__global__ void __launch_bounds__(32) k_recalc(const SuperDesc *__restrict__ sdp, const RecalcParams *__restrict__ pp, const RecalcJob *__restrict__ jobs) {
    const SuperDesc &sd = *sdp;
    const RecalcParams &P = *pp;
    const RecalcJob job = jobs[blockIdx.y];
    const int lane = threadIdx.x;
    const int blkIdx = blockIdx.x;
    const int n = P.nBlkX * P.nBlkY;
    int *hdr = (int *)job.out;
    VEC *res = (VEC *)(job.out + 12);
    const int old_valid = __ldg((const int *)job.old + 1) == 1;
    const bool live = job.ref != nullptr && old_valid;
    if (blkIdx == 0 && lane == 0) {
        hdr[0] = P.vectors_size;
        hdr[1] = live ? 1 : 0;
        hdr[2] = 4 + n * 16;
        if (P.divide)
            *(int *)(job.out + 12 + n * 16) = 4 + n * 16 * 4;
    }
    if (!live) { // gopWriteDefaultToArray, GroupOfPlanes.c:150-164
        VEC d;
        d.x = 0;
        d.y = 0;
        d.sad = P.verybigSAD;
        if (lane == 0)
            stvec(&res[blkIdx], d);
        if (P.divide && lane < 4) {
            VEC *q = (VEC *)(job.out + 12 + n * 16 + 4);
            const int by = blkIdx / P.nBlkX, bx = blkIdx - by * P.nBlkX;
            stvec(&q[(by * 2 + (lane >> 1)) * (P.nBlkX * 2) + bx * 2 + (lane & 1)], d);
        }
        return;
    }
    State S;
    S.lane = lane;
    S.nBlkX = P.nBlkX;
    S.nBlkY = P.nBlkY;
    S.blkW = P.blkW;
    S.blkH = P.blkH;
    S.ovX = P.ovX;
    S.ovY = P.ovY;
    S.logPel = P.logPel;
    S.logScale = 0;
    S.logxR = sd.logxR;
    S.logyR = sd.logyR;
    S.bps = P.bps;
    S.chroma = P.chroma;
    S.smallest = 1;
    S.cw = P.blkW >> sd.logxR;
    S.ch = P.blkH >> sd.logyR;
    S.dctmode = P.dctmode;
    S.dctweight16 = 8; // :1166
    S.bits = P.bits;
    S.dcx = P.dct_cos;
    S.dcy = P.dct_cos ? P.dct_cos + P.blkW * P.blkW : nullptr;
    S.dtmp = (float *)(P.dct_scratch + ((long long)blockIdx.y * gridDim.x + blockIdx.x) * P.dct_scratch_stride);
    S.dsrc = (int *)(S.dtmp + P.blkW * P.blkH);
    S.dctshift = 0;
    while ((1 << S.dctshift) < P.blkW * P.blkH)
        S.dctshift++;
    S.verybigSAD = P.verybigSAD;
    S.vectors = res;
    S.searchType = P.search;
    S.searchParam = P.nSearchParam;
    S.penaltyNew = P.pnew;
    S.srcLuma = 0;
    bind_planes(S.src, job.src, sd, 0, P.chroma ? 3 : 1);
    bind_planes(S.ref, job.ref, sd, 0, P.chroma ? 3 : 1);
    S.blky = blkIdx / P.nBlkX;
    S.blkx = blkIdx - S.blky * P.nBlkX;
    S.blkIdx = blkIdx;
    const int stepX = P.blkW - P.ovX, stepY = P.blkH - P.ovY;
    S.x[0] = S.src[0].hpad + stepX * S.blkx;
    S.y[0] = S.src[0].vpad + stepY * S.blky;
    if (P.chroma) {
        S.x[1] = S.x[2] = S.src[1].hpad + (stepX >> sd.logxR) * S.blkx;
        S.y[1] = S.y[2] = S.src[1].vpad + (stepY >> sd.logyR) * S.blky;
    }
    for (int p = 0; p < (P.chroma ? 3 : 1); p++)
        S.pSrc[p] = S.src[p].base + (long long)S.y[p] * S.src[p].pitch + (long long)S.x[p] * S.bps;
    S.nLambda = S.blky == 0 ? 0 : P.nLambda / ((1 << P.logPel) * (1 << P.logPel));
    S.dxMax = (S.src[0].pw - S.x[0] - P.blkW) << P.logPel; // :1274-1277, the whole padded plane
    S.dyMax = (S.src[0].ph - S.y[0] - P.blkH) << P.logPel;
    S.dxMin = -(S.x[0] << P.logPel);
    S.dyMin = -(S.y[0] << P.logPel);
    // predictor from the old field (:1279-1330)
    const VEC *oldv = (const VEC *)(job.old + __ldg((const int *)job.old) - 16 * P.oBlkX * P.oBlkY);
    const int nStepXold = P.oBlkW - P.oOvX, nStepYold = P.oBlkH - P.oOvY;
    const int centerX = P.blkW / 2 + stepX * S.blkx;
    const int blkxold = (centerX - P.oBlkW / 2) / nStepXold;
    const int centerY = P.blkH / 2 + stepY * S.blky;
    const int blkyold = (centerY - P.oBlkH / 2) / nStepYold;
    const int deltaX = max(0, centerX - (P.oBlkW / 2 + nStepXold * blkxold));
    const int deltaY = max(0, centerY - (P.oBlkH / 2 + nStepYold * blkyold));
    const int bx1 = min(P.oBlkX - 1, max(0, blkxold)), bx2 = min(P.oBlkX - 1, max(0, blkxold + 1));
    const int by1 = min(P.oBlkY - 1, max(0, blkyold)), by2 = min(P.oBlkY - 1, max(0, blkyold + 1));
    VEC vo;
    if (P.smooth == 1) {
        const VEC v1 = ldvec(&oldv[bx1 + by1 * P.oBlkX]), v2 = ldvec(&oldv[bx2 + by1 * P.oBlkX]), v3 = ldvec(&oldv[bx1 + by2 * P.oBlkX]), v4 = ldvec(&oldv[bx2 + by2 * P.oBlkX]);
        const int a_x = v1.x * nStepXold + deltaX * (v2.x - v1.x);
        const int a_y = v1.y * nStepXold + deltaX * (v2.y - v1.y);
        const long long a_s = v1.sad * nStepXold + deltaX * (v2.sad - v1.sad);
        const int b_x = v3.x * nStepXold + deltaX * (v4.x - v3.x);
        const int b_y = v3.y * nStepXold + deltaX * (v4.y - v3.y);
        const long long b_s = v3.sad * nStepXold + deltaX * (v4.sad - v3.sad);
        vo.x = (a_x + deltaY * (b_x - a_x) / nStepYold) / nStepXold;
        vo.y = (a_y + deltaY * (b_y - a_y) / nStepYold) / nStepXold;
        vo.sad = (a_s + deltaY * (b_s - a_s) / nStepYold) / nStepXold;
    } else {
        const bool rx = deltaX * 2 >= nStepXold, ry = deltaY * 2 >= nStepYold;
        vo = ldvec(&oldv[(rx ? bx2 : bx1) + (ry ? by2 : by1) * P.oBlkX]);
    }
    vo.x = (vo.x << P.logPel) >> P.oLogPel;
    vo.y = (vo.y << P.logPel) >> P.oLogPel;
    S.predictor = clipmv(S, vo);
    S.predictor.sad = vo.sad * (P.blkW * P.blkH) / (P.oBlkW * P.oBlkH);
    S.bestMV = S.predictor;
    if (S.dctmode >= 1 && S.dctmode <= 4) { // make dct of source block (:824-828 / :1336-1338)
        unsigned dc;
        dct_block(S, S.pSrc[0], S.src[0].pitch, true, dc);
    }
    if (S.dctmode >= 3)
        S.srcLuma = (int)wsum(lane_luma(S.pSrc[0], S.src[0].pitch, S.blkW * S.bps, S.blkH, S.bps, S.lane));
    const long long sad = full_sad_at(S, S.predictor.x, S.predictor.y, S.predictor.x, S.predictor.y);
    S.bestMV.sad = sad;
    S.nMinCost = sad;
    if (S.bestMV.sad > P.thSAD)
        refine(S);
    if (lane == 0)
        stvec(&res[blkIdx], S.bestMV);
}

// gopExtraDivide (GroupOfPlanes.c:206-302) for Recalculate: thread per block, after k_recalc
__global__ void k_recalc_divide(const RecalcParams *__restrict__ pp, const RecalcJob *__restrict__ jobs) {
    const RecalcParams &P = *pp;
    const RecalcJob job = jobs[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int nBlkX = P.nBlkX, nBlkY = P.nBlkY, n = nBlkX * nBlkY;
    if (i >= n || ((const int *)job.out)[1] != 1)
        return;
    const VEC *in = (const VEC *)(job.out + 12);
    VEC *out = (VEC *)(job.out + 12 + n * 16 + 4);
    const int by = i / nBlkX, bx = i - by * nBlkX;
    VEC b = ldvec(&in[i]);
    b.sad >>= 2;
    VEC o[4] = {b, b, b, b};
    if (P.divide > 1 && by >= 1 && by < nBlkY - 1 && bx >= 1 && bx < nBlkX - 1) {
        VEC c = ldvec(&in[i]), lf = ldvec(&in[i - 1]), rt = ldvec(&in[i + 1]), up = ldvec(&in[i - nBlkX]), dn = ldvec(&in[i + nBlkX]);
        get_median(o[0].x, o[0].y, c.x, c.y, lf.x, lf.y, up.x, up.y);
        get_median(o[1].x, o[1].y, c.x, c.y, rt.x, rt.y, up.x, up.y);
        get_median(o[2].x, o[2].y, c.x, c.y, lf.x, lf.y, dn.x, dn.y);
        get_median(o[3].x, o[3].y, c.x, c.y, rt.x, rt.y, dn.x, dn.y);
    }
    VEC *q = out + (by * 2) * (nBlkX * 2) + bx * 2;
    stvec(&q[0], o[0]);
    stvec(&q[1], o[1]);
    stvec(&q[nBlkX * 2], o[2]);
    stvec(&q[nBlkX * 2 + 1], o[3]);
}

} // namespace

// cosine tables of the DCT modes: cx (w*w) then cy (h*h), t[k * n + i] = (float)(2 cos(pi (2i + 1) k / (2n))) evaluated in
// double on the host, exactly like the oracle and the FFTW stand-in (oracle/shim/fftw_shim.c)
static void dct_tables(int w, int h, std::vector<float> &t) {
    t.resize((size_t)w * w + (size_t)h * h);
    for (int pass = 0; pass < 2; pass++) {
        const int n = pass ? h : w;
        float *o = t.data() + (pass ? (size_t)w * w : 0);
        for (int k = 0; k < n; k++)
            for (int i = 0; i < n; i++)
                o[k * n + i] = (float)(2.0 * cos(M_PI * (2.0 * i + 1.0) * k / (2.0 * n)));
    }
}

int launch_search(Ctx *ctx, const SuperDesc &sd, const mvb200_super_geom &geom, const uint8_t *frames_base, int nframes, const SearchParams *host_params_in, int nparams, const SearchJob *host_jobs, int njobs) {
    if (njobs <= 0)
        return 0;
    const int hist_ints = 8192 * 2;
    std::vector<SearchParams> params_copy(host_params_in, host_params_in + nparams);
    SearchParams *host_params = params_copy.data();
    bool any_dct = false;
    for (int i = 0; i < nparams; i++)
        any_dct = any_dct || (host_params[i].dctmode >= 1 && host_params[i].dctmode <= 4);
    std::vector<float> cos_tab;
    size_t bytes_cos = 0;
    if (any_dct) { // all instances of a launch share the block size (checked by the callers' similarity rules; enforce)
        for (int i = 1; i < nparams; i++)
            if (host_params[i].blkW != host_params[0].blkW || host_params[i].blkH != host_params[0].blkH)
                return set_error("mvtools_b200: dct 1..4 instances of one launch must share the block size");
        dct_tables(host_params[0].blkW, host_params[0].blkH, cos_tab);
        bytes_cos = cos_tab.size() * sizeof(float);
    }
    size_t bytes_sd = sizeof(SuperDesc), bytes_p = sizeof(SearchParams) * (size_t)nparams, bytes_j = sizeof(SearchJob) * (size_t)njobs;
    size_t off_p = (bytes_sd + 255) & ~(size_t)255, off_j = (off_p + bytes_p + 255) & ~(size_t)255, off_c = (off_j + bytes_j + 255) & ~(size_t)255, total = off_c + bytes_cos;
    uint8_t *host = (uint8_t *)ctx->get_pinned(total);
    uint8_t *dev = (uint8_t *)ctx->get_scratch(0, total);
    int *hist = (int *)ctx->get_scratch(1, sizeof(int) * (size_t)hist_ints * (size_t)njobs);
    if (!host || !dev || !hist)
        return 1;
    if (any_dct) {
        const long long stride = (long long)host_params[0].blkW * host_params[0].blkH * 8;
        uint8_t *scr = (uint8_t *)ctx->get_scratch(7, (size_t)stride * njobs);
        if (!scr)
            return 1;
        for (int i = 0; i < nparams; i++) {
            host_params[i].dct_cos = (const float *)(dev + off_c);
            host_params[i].dct_scratch = scr;
            host_params[i].dct_scratch_stride = stride;
        }
    }
    // the pinned staging buffer is reused from call to call: wait until the previous upload out of it is done
    if (ctx->staging_wait())
        return 1;
    memcpy(host, &sd, bytes_sd);
    memcpy(host + off_p, host_params, bytes_p);
    memcpy(host + off_j, host_jobs, bytes_j);
    if (bytes_cos)
        memcpy(host + off_c, cos_tab.data(), bytes_cos);
    MVB_CUDA_CHECK(cudaMemcpyAsync(dev, host, total, cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->staging_posted())
        return 1;
    // specialised kernel when every parameter set qualifies (all mv.Analyse defaults do); the
    // generic kernel otherwise, or when MVB200_FORCE_GENERIC=1 (used by the tests to cover both)
    bool handled = false;
    const char *force = getenv("MVB200_FORCE_GENERIC");
    const char *notma = getenv("MVB200_NO_TMA");
    if (!(force && force[0] == '1') && !(notma && notma[0] == '1'))
        if (launch_search_tma(ctx, sd, geom, frames_base, nframes, host_params, nparams, (const SuperDesc *)dev, (const SearchParams *)(dev + off_p), (const SearchJob *)(dev + off_j), njobs, hist, hist_ints, &handled))
            return 1;
    if (handled)
        return 0;
    if (!(force && force[0] == '1'))
        if (launch_search_fast(ctx, sd, host_params, nparams, (const SuperDesc *)dev, (const SearchParams *)(dev + off_p), (const SearchJob *)(dev + off_j), njobs, hist, hist_ints, &handled))
            return 1;
    if (handled)
        return 0;
    k_search<<<njobs, 32, 0, ctx->stream>>>((const SuperDesc *)dev, (const SearchParams *)(dev + off_p), (const SearchJob *)(dev + off_j), njobs, hist, hist_ints);
    ctx->launches++;
    MVB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int launch_recalc(Ctx *ctx, const SuperDesc &sd, const RecalcParams &P_in, const RecalcJob *host_jobs, int njobs) {
    if (njobs <= 0)
        return 0;
    RecalcParams P = P_in;
    const bool dct = P.dctmode >= 1 && P.dctmode <= 4;
    std::vector<float> cos_tab;
    if (dct)
        dct_tables(P.blkW, P.blkH, cos_tab);
    const size_t bytes_cos = cos_tab.size() * sizeof(float);
    size_t bytes_sd = sizeof(SuperDesc), bytes_p = sizeof(RecalcParams), bytes_j = sizeof(RecalcJob) * (size_t)njobs;
    size_t off_p = (bytes_sd + 255) & ~(size_t)255, off_j = (off_p + bytes_p + 255) & ~(size_t)255, off_c = (off_j + bytes_j + 255) & ~(size_t)255, total = off_c + bytes_cos;
    uint8_t *host = (uint8_t *)ctx->get_pinned(total);
    uint8_t *dev = (uint8_t *)ctx->get_scratch(0, total);
    if (!host || !dev)
        return 1;
    if (dct) {
        const long long stride = (long long)P.blkW * P.blkH * 8;
        uint8_t *scr = (uint8_t *)ctx->get_scratch(7, (size_t)stride * P.nBlkX * P.nBlkY * njobs);
        if (!scr)
            return 1;
        P.dct_cos = (const float *)(dev + off_c);
        P.dct_scratch = scr;
        P.dct_scratch_stride = stride;
    }
    if (ctx->staging_wait())
        return 1;
    memcpy(host, &sd, bytes_sd);
    memcpy(host + off_p, &P, bytes_p);
    memcpy(host + off_j, host_jobs, bytes_j);
    if (bytes_cos)
        memcpy(host + off_c, cos_tab.data(), bytes_cos);
    MVB_CUDA_CHECK(cudaMemcpyAsync(dev, host, total, cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->staging_posted())
        return 1;
    const int n = P.nBlkX * P.nBlkY;
    // the specialised kernel when the new grid is one of its compiled configurations; MVB200_FORCE_GENERIC=1 / MVB200_NO_TMA=1
    // keep the generic one (the tests cover both)
    bool handled = false;
    const char *force = getenv("MVB200_FORCE_GENERIC"), *notma = getenv("MVB200_NO_TMA");
    if (!(force && force[0] == '1') && !(notma && notma[0] == '1'))
        if (launch_recalc_tma(ctx, sd, P, (const SuperDesc *)dev, (const RecalcParams *)(dev + off_p), (const RecalcJob *)(dev + off_j), njobs, &handled))
            return 1;
    if (!handled) {
        k_recalc<<<dim3(n, njobs), 32, 0, ctx->stream>>>((const SuperDesc *)dev, (const RecalcParams *)(dev + off_p), (const RecalcJob *)(dev + off_j));
        ctx->launches++;
    }
    if (P.divide) {
        k_recalc_divide<<<dim3((n + 127) / 128, njobs), 128, 0, ctx->stream>>>((const RecalcParams *)(dev + off_p), (const RecalcJob *)(dev + off_j));
        ctx->launches++;
    }
    MVB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

} // namespace mvb200
