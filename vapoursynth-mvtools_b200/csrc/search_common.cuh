// Device helpers shared by the generic search kernel (search.cu) and the specialised
// fast kernel (search_fast.cu): VECTOR access, the block-parallel between-level passes.
#pragma once
#include "engine.h"

namespace mvb200 {
namespace {

struct VEC {
    int x, y;
    long long sad;
};

__device__ __forceinline__ VEC ldvec(const VEC *p) {
    // vectors are written by this warp earlier in the chain: L2-coherent (.cg) loads.  The VECTOR
    // arrays of the blob are only 4-byte aligned (8-byte blob header + 4-byte level headers).
    const int *q = (const int *)p;
    int4 v;
    v.x = __ldcg(q);
    v.y = __ldcg(q + 1);
    v.z = __ldcg(q + 2);
    v.w = __ldcg(q + 3);
    VEC r;
    r.x = v.x;
    r.y = v.y;
    r.sad = (long long)(((unsigned long long)(unsigned)v.w << 32) | (unsigned)v.z);
    return r;
}

__device__ __forceinline__ void stvec(VEC *p, const VEC &v) {
    int *q = (int *)p;
    __stcg(q, v.x);
    __stcg(q + 1, v.y);
    __stcg(q + 2, (int)(unsigned)(unsigned long long)v.sad);
    __stcg(q + 3, (int)(unsigned)((unsigned long long)v.sad >> 32));
}

// pobEstimateGlobalMVDoubled, PlaneOfBlocks.cpp:1559-1636, lanes split the blocks.
// hist: this chain's zero-initialised scratch of `freqSize` ints; left zeroed again.
__device__ void estimate_global_mv(const VEC *v, int n, int freqSize, int *hist, VEC &g, int lane) {
    int med[2];
    for (int c = 0; c < 2; c++) {
        int indmin = freqSize - 1, indmax = 0;
        for (int i = lane; i < n; i += 32) {
            VEC t = ldvec(&v[i]);
            int ind = (freqSize >> 1) + (c ? t.y : t.x);
            if (ind >= 0 && ind < freqSize) {
                atomicAdd(&hist[ind], 1);
                indmax = max(indmax, ind);
                indmin = min(indmin, ind);
            }
        }
        indmin = __reduce_min_sync(0xffffffffu, indmin);
        indmax = __reduce_max_sync(0xffffffffu, indmax);
        __syncwarp();
        // first maximum in [indmin, indmax]: highest count, ties to the lowest index
        int bestc = -1, besti = 0x7fffffff;
        for (int i = indmin + lane; i <= indmax; i += 32) {
            int cnt = *(volatile int *)&hist[i];
            if (cnt > bestc) {
                bestc = cnt;
                besti = i;
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            int oc = __shfl_xor_sync(0xffffffffu, bestc, o), oi = __shfl_xor_sync(0xffffffffu, besti, o);
            if (oc > bestc || (oc == bestc && oi < besti)) {
                bestc = oc;
                besti = oi;
            }
        }
        if (indmin > indmax) // no vector fell into the histogram (cannot happen for n >= 1 with sane vectors)
            besti = indmin;
        med[c] = besti - (freqSize >> 1);
        __syncwarp();
        for (int i = indmin + lane; i <= indmax; i += 32)
            hist[i] = 0;
        __syncwarp();
    }
    int meanvx = 0, meanvy = 0, num = 0;
    for (int i = lane; i < n; i += 32) {
        VEC t = ldvec(&v[i]);
        if (abs(t.x - med[0]) < 6 && abs(t.y - med[1]) < 6) {
            meanvx += t.x;
            meanvy += t.y;
            num++;
        }
    }
    meanvx = __reduce_add_sync(0xffffffffu, meanvx);
    meanvy = __reduce_add_sync(0xffffffffu, meanvy);
    num = __reduce_add_sync(0xffffffffu, num);
    if (num > 0) {
        g.x = 2 * meanvx / num;
        g.y = 2 * meanvy / num;
    } else {
        g.x = 2 * med[0];
        g.y = 2 * med[1];
    }
}

// pobInterpolatePrediction, PlaneOfBlocks.cpp:1447-1514, lanes split the blocks
__device__ void interpolate_prediction(VEC *dst, int nBlkX, int nBlkY, int logPel, const VEC *v2, int nBlkX2, int nBlkY2, int blkW, int blkH, int ovX, int ovY, int lane) {
    int normFactor = 3 - logPel; // logPel of the coarser level is always 0
    int mulFactor = normFactor < 0 ? -normFactor : 0;
    normFactor = normFactor < 0 ? 0 : normFactor;
    int normov = (blkW - ovX) * (blkH - ovY);
    int aoddx = blkW * 3 - ovX * 2, aevenx = blkW * 3 - ovX * 4;
    int aoddy = blkH * 3 - ovY * 2, aeveny = blkH * 3 - ovY * 4;
    double scaleov = __ddiv_rn(1.0, (double)normov);
    int total = nBlkX * nBlkY;
    for (int index = lane; index < total; index += 32) {
        int l = index / nBlkX, k = index - l * nBlkX;
        int i = k, j = l;
        if (i >= 2 * nBlkX2) i = 2 * nBlkX2 - 1;
        if (j >= 2 * nBlkY2) j = 2 * nBlkY2 - 1;
        int offy = -1 + 2 * (j % 2), offx = -1 + 2 * (i % 2);
        bool edgex = (i == 0) || (i >= 2 * nBlkX2 - 1), edgey = (j == 0) || (j >= 2 * nBlkY2 - 1);
        VEC v1, w2, v3, v4;
        v1 = ldvec(&v2[i / 2 + (j / 2) * nBlkX2]);
        if (edgex) {
            w2 = v1;
            if (edgey) {
                v3 = v4 = v1;
            } else {
                v3 = v4 = ldvec(&v2[i / 2 + (j / 2 + offy) * nBlkX2]);
            }
        } else if (edgey) {
            w2 = v1;
            v3 = v4 = ldvec(&v2[i / 2 + offx + (j / 2) * nBlkX2]);
        } else {
            w2 = ldvec(&v2[i / 2 + offx + (j / 2) * nBlkX2]);
            v3 = ldvec(&v2[i / 2 + (j / 2 + offy) * nBlkX2]);
            v4 = ldvec(&v2[i / 2 + offx + (j / 2 + offy) * nBlkX2]);
        }
        long long temp_sad;
        int vx, vy;
        if (ovX == 0 && ovY == 0) {
            vx = 9 * v1.x + 3 * w2.x + 3 * v3.x + v4.x;
            vy = 9 * v1.y + 3 * w2.y + 3 * v3.y + v4.y;
            temp_sad = 9 * v1.sad + 3 * w2.sad + 3 * v3.sad + v4.sad + 8;
        } else {
            int ax1 = offx > 0 ? aoddx : aevenx;
            int ax2 = (blkW - ovX) * 4 - ax1;
            int ay1 = offy > 0 ? aoddy : aeveny;
            int ay2 = (blkH - ovY) * 4 - ay1;
            long long a11 = ax1 * ay1, a12 = ax1 * ay2, a21 = ax2 * ay1, a22 = ax2 * ay2;
            vx = __double2int_rz(__dmul_rn((double)(a11 * v1.x + a21 * w2.x + a12 * v3.x + a22 * v4.x), scaleov));
            vy = __double2int_rz(__dmul_rn((double)(a11 * v1.y + a21 * w2.y + a12 * v3.y + a22 * v4.y), scaleov));
            temp_sad = __double2ll_rz(__dmul_rn((double)(a11 * v1.sad + a21 * w2.sad + a12 * v3.sad + a22 * v4.sad), scaleov));
        }
        VEC o;
        o.x = (vx >> normFactor) * (1 << mulFactor);
        o.y = (vy >> normFactor) * (1 << mulFactor);
        o.sad = temp_sad >> 4;
        stvec(&dst[index], o);
    }
}

// GetMedian, GroupOfPlanes.c:179-203
__device__ __forceinline__ int med3(int a, int b, int c) {
    if (((b <= a) && (a <= c)) || ((c <= a) && (a <= b))) return a;
    if (((a <= b) && (b <= c)) || ((c <= b) && (b <= a))) return b;
    return c;
}
__device__ __forceinline__ void get_median(int &vx, int &vy, int x1, int y1, int x2, int y2, int x3, int y3) {
    vx = med3(x1, x2, x3);
    vy = med3(y1, y2, y3);
    if ((vx == x1 && vy == y1) || (vx == x2 && vy == y2) || (vx == x3 && vy == y3))
        return;
    vx = x1;
    vy = y1;
}


// blob header + level headers (pobWriteHeaderToArray :413-416); vectors zeroed like pobInit
// (:354-355) or, without a reference frame, the defaults of gopWriteDefaultToArray
// (GroupOfPlanes.c:150-164)
__device__ void chain_prologue(const SearchJob &job, const SearchParams &P, int lane) {
    const int L = P.levels;
    int *hdr = (int *)job.out;
    if (lane == 0) {
        hdr[0] = P.vectors_size;
        hdr[1] = job.ref ? 1 : 0;
    }
    // level headers (pobWriteHeaderToArray :413-416) and, without a reference frame, the
    // default vectors of gopWriteDefaultToArray (GroupOfPlanes.c:150-164)
    for (int l = 0; l < L; l++) {
        int n = P.nBlkX[l] * P.nBlkY[l];
        uint8_t *o = job.out + P.level_off[l];
        if (lane == 0)
            *(int *)o = 4 + n * 16;
        VEC *v = (VEC *)(o + 4);
        VEC init;
        init.x = 0;
        init.y = 0;
        init.sad = job.ref ? 0 : P.verybigSAD; // pobInit zeroes vectors[] (:354-355)
        for (int i = lane; i < n; i += 32)
            stvec(&v[i], init);
        if (l == 0 && P.divide) {
            uint8_t *o2 = o + 4 + n * 16;
            if (lane == 0)
                *(int *)o2 = 4 + n * 16 * 4;
            if (!job.ref) {
                VEC *v2 = (VEC *)(o2 + 4);
                for (int i = lane; i < n * 4; i += 32)
                    stvec(&v2[i], init);
            }
        }
    }
    __syncwarp();
}

// gopExtraDivide, GroupOfPlanes.c:206-302, lanes split the blocks
__device__ void chain_divide(const SearchJob &job, const SearchParams &P, int lane) {
    if (P.divide) {
        const int nBlkX = P.nBlkX[0], nBlkY = P.nBlkY[0], n = nBlkX * nBlkY;
        const VEC *in = (const VEC *)(job.out + P.level_off[0] + 4);
        VEC *out = (VEC *)(job.out + P.level_off[0] + 4 + n * 16 + 4);
        for (int i = lane; i < n; i += 32) {
            int by = i / nBlkX, bx = i - by * nBlkX;
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
    }
}

} // namespace
} // namespace mvb200
