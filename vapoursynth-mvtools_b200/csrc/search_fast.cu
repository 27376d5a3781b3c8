// Specialised search kernel for the common configurations (dct=0, no trymany, search
// types exhaustive / hex2 / UMH -- i.e. every default of mv.Analyse).  Same semantics
// as the generic kernel in search.cu (which stays as the fallback and as a cross-check
// in the tests), restructured for the hardware:
//
//   * the SOURCE block lives in registers for the whole block step (NW 32-bit words per
//     lane: luma + both chroma blocks are one flat list of words dealt round-robin to
//     the 32 lanes), and is prefetched one block ahead;
//   * candidates are evaluated in BATCHES: every pattern step whose candidate positions
//     do not depend on accepts inside the step (the 7 start candidates, the 6 hexagon
//     points, the 3 half-hexagon points, an expanding-square ring, the UMH cross and
//     hex-4 rings) issues all reference loads first -- B x NW x 2 independent loads in
//     flight per lane -- then one REDUX.SUM per candidate, then REPLAYS the accepts in
//     the reference's order with its strict '<' (first minimum wins), so results are
//     bit-identical while the dependent-latency chain per block drops from ~45 round
//     trips to ~5;
//   * 8-bit blocks up to 16x16 pack the luma SAD (low 16 bits) and the chroma SAD (high
//     16 bits) into one accumulator: one VABSDIFF4.U8.ACC per word, one REDUX per
//     candidate (the cost needs the two sums separately because the pnew penalty is
//     shifted separately, PlaneOfBlocks.cpp:238,247);
//   * the predictor vectors of the NEXT block (hierarchical, up, down-right/up-right)
//     are loaded while the current block is searched;
//   * all chain state is in registers; nothing is called out of line.
#include "search_common.cuh"

namespace mvb200 {

namespace {

__constant__ int f_hex2[8][2] = {{-1, -2}, {-2, 0}, {-1, 2}, {1, 2}, {2, 0}, {1, -2}, {-1, -2}, {-2, 0}};
__constant__ int f_hex6[8][2] = {{-2, 0}, {-1, 2}, {1, 2}, {2, 0}, {1, -2}, {-1, -2}, {0, 0}, {0, 0}};
__constant__ int f_hex4[16][2] = {{-4, 2}, {-4, 1}, {-4, 0}, {-4, -1}, {-4, -2}, {4, -2}, {4, -1}, {4, 0},
                                  {4, 1},  {4, 2},  {2, 3},  {0, 4},   {-2, 3},  {-2, -3}, {0, -4}, {2, -3}};

struct FPlane {
    const uint8_t *sbase; // source level base (padded origin, sub-plane 0)
    const uint8_t *rbase; // reference level base
    int pitch, sub, hpad, vpad;
};

template <int NW, int BPS, bool PACK>
struct Fast {
    static constexpr int B = (NW <= 3) ? 8 : 4; // candidates per batch

    // ---- per-lane item tables
    const uint8_t *sp[NW]; // source pointer of the lane's k-th word at block origin (0,0) of its plane
    const uint8_t *rp[NW];
    unsigned srcw[NW], nextw[NW];
    unsigned lumaMask, actMask; // bit k: item k is a luma word / is active
    // ---- level constants
    FPlane pl[3];
    int pw0, ph0;
    int nBlkX, nBlkY, blkW, blkH, ovX, ovY, logPel, logScale, logxR, logyR, chroma, smallest;
    int searchType, searchParam, penaltyZero, pglobal, penaltyNew, badrange;
    long long badSAD, LSAD;
    VEC *vectors;
    VEC globalMV;
    int fieldShift;
    // ---- block state
    int x0, y0, xc, yc;
    int dxMin, dxMax, dyMin, dyMax;
    long long nLambda, nMinCost;
    VEC best, predictor;
    int badcount, lane;

    __device__ __forceinline__ bool ok(int vx, int vy) const { return vx >= dxMin && vy >= dyMin && vx < dxMax && vy < dyMax; }
    __device__ __forceinline__ VEC clip(VEC v) const {
        v.x = min(max(v.x, dxMin), dxMax - 1);
        v.y = min(max(v.y, dyMin), dyMax - 1);
        return v;
    }

    __device__ __forceinline__ static unsigned ldw(const uint8_t *p, unsigned sh) {
        // p is 4-byte aligned; sh = 8 * (byte misalignment), warp-uniform
        unsigned lo = __ldg((const unsigned *)p);
        if (sh == 0)
            return lo;
        unsigned hi = __ldg((const unsigned *)p + 1);
        return __funnelshift_r(lo, hi, sh);
    }

    // byte offset of the block addressed by a luma-unit vector inside plane class c (0 luma, 1 chroma)
    __device__ __forceinline__ int cand_off(int c, int vx, int vy) const {
        const int m = (1 << logPel) - 1;
        int X, Y;
        if (c == 0) {
            X = (x0 << logPel) + vx;
            Y = (y0 << logPel) + vy;
        } else { // pobGetRefBlockU/V, PlaneOfBlocks.cpp:56-101: the luma vector is divided toward zero
            int xbias = (vx < 0) * ((1 << logxR) - 1), ybias = (vy < 0) * ((1 << logyR) - 1);
            X = (xc << logPel) + ((vx + xbias) >> logxR);
            Y = (yc << logPel) + ((vy + ybias) >> logyR);
        }
        int idx = (X & m) | ((Y & m) << logPel);
        return idx * pl[c].sub + (Y >> logPel) * pl[c].pitch + (X >> logPel) * BPS;
    }

    // Evaluates up to B candidates.  lvx/lvy: luma vectors, cvx/cvy: the vectors the chroma position is
    // derived from (differs from the luma vector only for the zero candidate under a field shift).
    // valid: bit j set = evaluate candidate j.  Results are warp-uniform.
    __device__ __forceinline__ void eval_batch(const int (&lvx)[B], const int (&lvy)[B], const int (&cvx)[B], const int (&cvy)[B], unsigned valid,
                                               unsigned (&sadY)[B], unsigned (&sadUV)[B]) const {
        unsigned accA[B], accB[B];
#pragma unroll
        for (int j = 0; j < B; j++) {
            accA[j] = 0;
            accB[j] = 0;
            if (valid & (1u << j)) {
                const int o0 = cand_off(0, lvx[j], lvy[j]);
                const int o1 = chroma ? cand_off(1, cvx[j], cvy[j]) : 0;
                const int a0 = o0 & ~3, a1 = o1 & ~3;
                const unsigned s0 = (o0 & 3) * 8, s1 = (o1 & 3) * 8;
#pragma unroll
                for (int k = 0; k < NW; k++) {
                    if (actMask & (1u << k)) {
                        const bool isl = lumaMask & (1u << k);
                        unsigned w = ldw(rp[k] + (isl ? a0 : a1), isl ? s0 : s1);
                        unsigned d = BPS == 1 ? __vsadu4(w, srcw[k]) : __vsadu2(w, srcw[k]);
                        if (PACK)
                            accA[j] += isl ? d : (d << 16);
                        else {
                            accA[j] += isl ? d : 0u;
                            accB[j] += isl ? 0u : d;
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < B; j++) {
            if (valid & (1u << j)) {
                unsigned r = __reduce_add_sync(0xffffffffu, accA[j]);
                if (PACK) {
                    sadY[j] = r & 0xffffu;
                    sadUV[j] = r >> 16;
                } else {
                    sadY[j] = r;
                    sadUV[j] = chroma ? __reduce_add_sync(0xffffffffu, accB[j]) : 0u;
                }
            } else {
                sadY[j] = 0;
                sadUV[j] = 0;
            }
        }
    }

    // pobMotionDistorsion, PlaneOfBlocks.cpp:105-114
    __device__ __forceinline__ long long mv_cost(int vx, int vy) const {
        int ddx = predictor.x - vx, ddy = predictor.y - vy;
        int dist = (int)((unsigned)(ddx * ddx) + (unsigned)(ddy * ddy));
        return (long long)(int)((nLambda * dist) >> 8);
    }

    // the accept half of pobCheckMV_Template (:219-261) for an already evaluated candidate
    __device__ __forceinline__ void accept(int vx, int vy, unsigned sy, unsigned suv, bool penaltyNew_, bool updateBest, bool updateDir, int &dir, int val) {
        long long cost = mv_cost(vx, vy);
        long long sad = sy, saduv = suv;
        cost += sad + (penaltyNew_ ? ((penaltyNew * sad) >> 8) : 0);
        cost += saduv + (penaltyNew_ ? ((penaltyNew * saduv) >> 8) : 0);
        if (cost >= nMinCost)
            return;
        if (updateBest) {
            best.x = vx;
            best.y = vy;
        }
        nMinCost = cost;
        best.sad = sad + saduv;
        if (updateDir)
            dir = val;
    }

    // ---- candidate patterns, as index -> offset generators in the reference's order
    // pobExpandingSearch (:635-658): ring r, step s; returns the number of points
    __device__ __forceinline__ static int expand_count(int r, int s) {
        int m = (2 * r - s - 1) / s + 1; // iterations of `for (i = -r + s; i < r; i += s)`
        if (-r + s >= r)
            m = 0;
        return 4 * m + 4;
    }
    __device__ __forceinline__ static void expand_pt(int r, int s, int i, int &dx, int &dy) {
        int m = (-r + s >= r) ? 0 : (2 * r - s - 1) / s + 1;
        if (i < 2 * m) {
            dx = -r + s + (i >> 1) * s;
            dy = (i & 1) ? r : -r;
        } else if (i < 4 * m) {
            int t = i - 2 * m;
            dy = -r + s + (t >> 1) * s;
            dx = (t & 1) ? r : -r;
        } else {
            int t = i - 4 * m;
            dx = (t & 2) ? r : -r;
            dy = (t & 1) ? r : -r;
        }
    }

    enum { PAT_EXPAND = 0, PAT_HEX6 = 1, PAT_HEX3 = 2, PAT_CROSS = 3, PAT_HEX4 = 4 };

    // Runs one pattern step: `count` candidates around (cx, cy), CheckMV semantics (penalty new +
    // update best) or CheckMVdir semantics (penalty new + update dir, not the vector).
    template <int PAT>
    __device__ __forceinline__ void run_pattern(int cx, int cy, int count, int a, int b, bool dirMode, int &dir) {
        for (int base = 0; base < count; base += B) {
            int vx[B], vy[B], val[B];
            unsigned valid = 0;
#pragma unroll
            for (int j = 0; j < B; j++) {
                int i = base + j;
                int dx = 0, dy = 0, v = 0;
                if (PAT == PAT_EXPAND) {
                    expand_pt(a, b, i, dx, dy);
                } else if (PAT == PAT_HEX6) { // :682-687
                    dx = f_hex6[i & 7][0];
                    dy = f_hex6[i & 7][1];
                    v = i;
                } else if (PAT == PAT_HEX3) { // :706-708, a = odir
                    int t = (a + i) & 7; // index into hex2[8]
                    dx = f_hex2[t][0];
                    dy = f_hex2[t][1];
                    v = a - 1 + i;
                } else if (PAT == PAT_CROSS) { // pobCrossSearch :727-739 with start 1, x_max = y_max = a
                    int per = a / 2; // iterations of `for (i = 1; i < a; i += 2)`
                    int t = i < 2 * per ? i : i - 2 * per;
                    int d = 1 + 2 * (t >> 1);
                    d = (t & 1) ? d : -d;
                    dx = i < 2 * per ? d : 0;
                    dy = i < 2 * per ? 0 : d;
                } else { // PAT_HEX4 :755-763, a = scale
                    dx = f_hex4[i & 15][0] * a;
                    dy = f_hex4[i & 15][1] * a;
                }
                vx[j] = cx + dx;
                vy[j] = cy + dy;
                val[j] = v;
                if (i < count && ok(vx[j], vy[j]) && mv_cost(vx[j], vy[j]) < nMinCost)
                    valid |= 1u << j;
            }
            if (!valid)
                continue;
            unsigned sy[B], suv[B];
            eval_batch(vx, vy, vx, vy, valid, sy, suv);
#pragma unroll
            for (int j = 0; j < B; j++)
                if (valid & (1u << j))
                    accept(vx[j], vy[j], sy[j], suv[j], true, !dirMode, dirMode, dir, val[j]);
        }
    }

    __device__ __forceinline__ void expanding(int r, int s, int cx, int cy) {
        int dir = 0;
        run_pattern<PAT_EXPAND>(cx, cy, expand_count(r, s), r, s, false, dir);
    }

    // pobHex2Search, PlaneOfBlocks.cpp:661-724
    __device__ __forceinline__ void hex2(int range) {
        int dir = -2;
        int bmx = best.x, bmy = best.y;
        if (range > 1) {
            run_pattern<PAT_HEX6>(bmx, bmy, 6, 0, 0, true, dir);
            if (dir != -2) {
                bmx += f_hex2[(dir + 1) & 7][0];
                bmy += f_hex2[(dir + 1) & 7][1];
                for (int i = 1; i < range / 2 && ok(bmx, bmy); i++) {
                    const int odir = (dir + 1 == 0) ? 5 : ((dir + 1 == 7) ? 0 : dir); // mod6m1[dir + 1]
                    dir = -2;
                    run_pattern<PAT_HEX3>(bmx, bmy, 3, odir, 0, true, dir);
                    if (dir == -2)
                        break;
                    bmx += f_hex2[(dir + 1) & 7][0];
                    bmy += f_hex2[(dir + 1) & 7][1];
                }
            }
            best.x = bmx;
            best.y = bmy;
        }
        expanding(1, 1, bmx, bmy);
    }

    // pobUMHSearch, PlaneOfBlocks.cpp:742-769
    __device__ __forceinline__ void umh(int range, int omx, int omy) {
        int dir = 0;
        run_pattern<PAT_CROSS>(omx, omy, 4 * (range / 2), range, 0, false, dir);
        int i = 1;
        do {
            run_pattern<PAT_HEX4>(omx, omy, 16, i, 0, false, dir);
        } while (++i <= range / 4);
        hex2(range);
    }

    // pobRefine, PlaneOfBlocks.cpp:772-816 (types 3, 4, 5 only on this path)
    __device__ __forceinline__ void refine() {
        if (searchType == 3) {
            int mvx = best.x, mvy = best.y;
            for (int i = 1; i <= searchParam; i++)
                expanding(i, 1, mvx, mvy);
        } else if (searchType == 4) {
            hex2(searchParam);
        } else {
            umh(searchParam, best.x, best.y);
        }
    }

    // pobPseudoEPZSearch, PlaneOfBlocks.cpp:819-968, given the fetched predictors
    __device__ __forceinline__ void epz(const VEC (&pr)[4], int blkIdx) {
        // the seven start candidates in one batch: zero, global, hierarchical, predictors[0..3]
        globalMV = clip(globalMV); // sticky, as in the reference (:859)
        int lvx[8], lvy[8], cvx[8], cvy[8];
        lvx[0] = 0;
        lvy[0] = fieldShift;
        cvx[0] = 0;
        cvy[0] = 0; // zero: luma at (0, fieldShift), chroma at (0, 0)  (:836-839)
        lvx[1] = cvx[1] = globalMV.x;
        lvy[1] = cvy[1] = globalMV.y;
        lvx[2] = cvx[2] = predictor.x;
        lvy[2] = cvy[2] = predictor.y;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            lvx[3 + i] = cvx[3 + i] = pr[i].x;
            lvy[3 + i] = cvy[3 + i] = pr[i].y;
        }
        lvx[7] = lvy[7] = cvx[7] = cvy[7] = 0;
        unsigned sy[8], suv[8];
        if (B == 8) {
            unsigned valid = 7u;
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (ok(pr[i].x, pr[i].y))
                    valid |= 8u << i;
            eval8(lvx, lvy, cvx, cvy, valid, sy, suv);
        } else {
            eval8(lvx, lvy, cvx, cvy, 0, sy, suv);
        }
        // zero (:834-846): penalty on the summed SAD, accepted unconditionally
        long long sad = (long long)sy[0] + suv[0];
        best.x = 0;
        best.y = fieldShift;
        best.sad = sad;
        nMinCost = sad + ((penaltyZero * sad) >> 8);
        // global (:859-877)
        sad = (long long)sy[1] + suv[1];
        long long cost = sad + ((pglobal * sad) >> 8);
        if (cost < nMinCost) {
            best.x = globalMV.x;
            best.y = globalMV.y;
            best.sad = sad;
            nMinCost = cost;
        }
        // hierarchical predictor (:884-901)
        sad = (long long)sy[2] + suv[2];
        if (sad < nMinCost) {
            best.x = predictor.x;
            best.y = predictor.y;
            best.sad = sad;
            nMinCost = sad;
        }
        // spatial predictors via CheckMV0 (:910-922)
        int dummy = 0;
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (ok(pr[i].x, pr[i].y))
                accept(pr[i].x, pr[i].y, sy[3 + i], suv[3 + i], false, true, false, dummy, 0);
        refine();
        // bad-block wide search (:938-963)
        long long foundSAD = best.sad;
        if (blkIdx > 1 && foundSAD > (badSAD + badSAD * badcount / 16)) {
            badcount++;
            if (badrange > 0) {
                umh(badrange * (1 << logPel), 0, 0);
            } else if (badrange < 0) {
                for (int i = 1; i < -badrange * (1 << logPel); i += (1 << logPel)) {
                    expanding(i, 1 << logPel, 0, 0);
                    if (best.sad < foundSAD / 4)
                        break;
                }
            }
            int mvx = best.x, mvy = best.y;
            for (int i = 1; i < (1 << logPel); i++)
                expanding(i, 1, mvx, mvy);
        }
    }

    // evaluates 8 start candidates with batches of B (valid == 0 means "bounds-check 3..6 here")
    __device__ __forceinline__ void eval8(const int (&lvx)[8], const int (&lvy)[8], const int (&cvx)[8], const int (&cvy)[8], unsigned valid8, unsigned (&sy)[8], unsigned (&suv)[8]) const {
        if (B == 8) {
            int a[B], b[B], c[B], d[B];
            unsigned y[B], u[B];
#pragma unroll
            for (int j = 0; j < B; j++) {
                a[j] = lvx[j & 7];
                b[j] = lvy[j & 7];
                c[j] = cvx[j & 7];
                d[j] = cvy[j & 7];
            }
            eval_batch(a, b, c, d, valid8, y, u);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                sy[j] = y[j % B];
                suv[j] = u[j % B];
            }
        } else {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                int a[B], b[B], c[B], d[B];
                unsigned y[B], u[B];
                unsigned valid = 0;
#pragma unroll
                for (int j = 0; j < B; j++) {
                    int i = (h * 4 + j) & 7;
                    a[j] = lvx[i];
                    b[j] = lvy[i];
                    c[j] = cvx[i];
                    d[j] = cvy[i];
                    if (j < 4 && i < 7 && (i < 3 || ok(lvx[i], lvy[i])))
                        valid |= 1u << j;
                }
                eval_batch(a, b, c, d, valid, y, u);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    sy[h * 4 + j] = y[j % B];
                    suv[h * 4 + j] = u[j % B];
                }
            }
        }
    }
};

template <int NW, int BPS, bool PACK>
__global__ void __launch_bounds__(32) k_search_fast(const SuperDesc *__restrict__ sdp, const SearchParams *__restrict__ params, const SearchJob *__restrict__ jobs, int njobs,
                                                    int *__restrict__ hist_all, int hist_ints) {
    const int chain = blockIdx.x;
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
    VEC globalMVec;
    globalMVec.x = 0;
    globalMVec.y = 0;
    globalMVec.sad = -1;
    const int pglobal = P.global ? P.pglobal : P.pzero;

    Fast<NW, BPS, PACK> S;
    S.lane = lane;
    S.blkW = P.blkW;
    S.blkH = P.blkH;
    S.ovX = P.ovX;
    S.ovY = P.ovY;
    S.logxR = sd.logxR;
    S.logyR = sd.logyR;
    S.chroma = P.chroma;
    S.penaltyZero = P.pzero;
    S.pglobal = pglobal;
    S.penaltyNew = P.pnew;
    S.badSAD = P.badSAD;
    S.badrange = P.badrange;
    S.LSAD = P.lsad;
    const int cw = P.blkW >> sd.logxR, ch = P.blkH >> sd.logyR;
    const int lumaWords = (P.blkW * BPS / 4) * P.blkH;
    const int chromaWords = P.chroma ? (cw * BPS / 4) * ch : 0;
    const int stepX = P.blkW - P.ovX, stepY = P.blkH - P.ovY;

    for (int l = L - 1; l >= 0; l--) {
        S.nBlkX = P.nBlkX[l];
        S.nBlkY = P.nBlkY[l];
        S.logPel = l == 0 ? P.logPel : 0;
        S.logScale = l;
        S.smallest = (l == L - 1);
        S.vectors = (VEC *)(job.out + P.level_off[l] + 4);
#pragma unroll
        for (int p = 0; p < 3; p++) {
            const int q = (p < sd.nplanes) ? p : 0;
            const mvb200_level_geom &g = sd.lv[q][l];
            const long long off = sd.offset[q] + (long long)g.row0 * sd.pitch[q];
            S.pl[p].sbase = job.src + off;
            S.pl[p].rbase = job.ref + off;
            S.pl[p].pitch = (int)sd.pitch[q];
            S.pl[p].sub = (int)(sd.pitch[q] * g.padded_h);
            S.pl[p].hpad = g.hpad;
            S.pl[p].vpad = g.vpad;
        }
        S.pw0 = sd.lv[0][l].padded_w;
        S.ph0 = sd.lv[0][l].padded_h;
        // deal the words of the block (luma, then U, then V) round-robin to the lanes
        S.lumaMask = 0;
        S.actMask = 0;
#pragma unroll
        for (int k = 0; k < NW; k++) {
            int i = lane + 32 * k;
            int p = 0, w = 0, wpr = 1;
            bool act = true;
            if (i < lumaWords) {
                p = 0;
                w = i;
                wpr = P.blkW * BPS / 4;
            } else if (i < lumaWords + chromaWords) {
                p = 1;
                w = i - lumaWords;
                wpr = cw * BPS / 4;
            } else if (i < lumaWords + 2 * chromaWords) {
                p = 2;
                w = i - lumaWords - chromaWords;
                wpr = cw * BPS / 4;
            } else {
                act = false;
            }
            int row = w / wpr, col = w - row * wpr;
            // (no dynamic indexing of S.pl[]: that would force the whole state into local memory)
            const int pitch_p = p == 0 ? S.pl[0].pitch : S.pl[1].pitch;
            const uint8_t *sb = p == 0 ? S.pl[0].sbase : (p == 1 ? S.pl[1].sbase : S.pl[2].sbase);
            const uint8_t *rb = p == 0 ? S.pl[0].rbase : (p == 1 ? S.pl[1].rbase : S.pl[2].rbase);
            int o = row * pitch_p + col * 4;
            S.sp[k] = sb + o;
            S.rp[k] = rb + o;
            if (act)
                S.actMask |= 1u << k;
            if (act && p == 0)
                S.lumaMask |= 1u << k;
        }
        int st, stp, tmFs;
        if (l == L - 1) {
            st = (L == 1) ? P.search : P.search_coarse;
            stp = (L == 1) ? P.nPelSearch : P.nSearchParam;
            tmFs = (L - 1 == 0) ? job.field_shift : 0;
        } else {
            st = (l == 0) ? P.search : P.search_coarse;
            stp = (l == 0) ? P.nPelSearch : P.nSearchParam;
            const VEC *coarse = (const VEC *)(job.out + P.level_off[l + 1] + 4);
            if (P.global)
                estimate_global_mv(coarse, P.nBlkX[l + 1] * P.nBlkY[l + 1], 8192 * 1 * 2, hist, globalMVec, lane);
            interpolate_prediction(S.vectors, S.nBlkX, S.nBlkY, S.logPel, coarse, P.nBlkX[l + 1], P.nBlkY[l + 1], P.blkW, P.blkH, P.ovX, P.ovY, lane);
            __syncwarp();
            tmFs = (l == 0) ? job.field_shift : 0;
        }
        // ---- doPobSearchMVs, PlaneOfBlocks.cpp:971-1131
        S.searchType = st;
        S.searchParam = stp;
        S.fieldShift = tmFs;
        S.globalMV.x = (1 << S.logPel) * globalMVec.x;
        S.globalMV.y = (1 << S.logPel) * globalMVec.y + tmFs;
        S.globalMV.sad = globalMVec.sad;
        const int nScale = 1 << l;
        int nLambdaLevel = P.nLambda / ((1 << S.logPel) * (1 << S.logPel));
        if (P.plevel == 1)
            nLambdaLevel = nLambdaLevel * nScale;
        else if (P.plevel == 2)
            nLambdaLevel = nLambdaLevel * nScale * nScale;
        S.badcount = 0;
        const int hp0 = S.pl[0].hpad, vp0 = S.pl[0].vpad, hpc = S.pl[1].hpad, vpc = S.pl[1].vpad;
        const int hps = hp0 >> l, vps = vp0 >> l;
        const int nBlk = S.nBlkX * S.nBlkY;
        VEC zeroFs;
        zeroFs.x = 0;
        zeroFs.y = tmFs;
        zeroFs.sad = 0;

        // scan-order bookkeeping for block `t` of the chain
        auto blk_of = [&](int t, int &bx, int &by, int &dir) {
            by = t / S.nBlkX;
            int i = t - by * S.nBlkX;
            dir = (by % 2 == 0 || P.meander == 0) ? 1 : -1;
            bx = dir == 1 ? i : S.nBlkX - 1 - i;
        };
        auto src_off = [&](int c, int bx, int by) {
            int x = c == 0 ? hp0 + stepX * bx : hpc + (stepX >> sd.logxR) * bx;
            int y = c == 0 ? vp0 + stepY * by : vpc + (stepY >> sd.logyR) * by;
            return y * S.pl[c].pitch + x * BPS;
        };
        auto load_src = [&](int bx, int by, unsigned (&dst)[NW]) {
            const int o0 = src_off(0, bx, by), o1 = P.chroma ? src_off(1, bx, by) : 0;
            const unsigned s0 = (o0 & 3) * 8, s1 = (o1 & 3) * 8;
#pragma unroll
            for (int k = 0; k < NW; k++) {
                dst[k] = 0;
                if (S.actMask & (1u << k)) {
                    const bool isl = S.lumaMask & (1u << k);
                    dst[k] = Fast<NW, BPS, PACK>::ldw(S.sp[k] + ((isl ? o0 : o1) & ~3), isl ? s0 : s1);
                }
            }
        };
        // predictor vectors of a block that do not depend on the previous block's result
        struct Pre {
            VEC hier, up, p3;
        };
        auto prefetch = [&](int bx, int by, int dir, Pre &pre) {
            const int idx = by * S.nBlkX + bx;
            pre.hier = ldvec(&S.vectors[idx]);
            const bool hasNext = (dir == 1 && bx < S.nBlkX - 1) || (dir == -1 && bx > 0);
            pre.up = zeroFs;
            if (by > 0)
                pre.up = ldvec(&S.vectors[idx - S.nBlkX]);
            pre.p3 = zeroFs;
            if (by < S.nBlkY - 1 && hasNext)
                pre.p3 = ldvec(&S.vectors[idx + S.nBlkX + dir]);
            else if (by > 0 && hasNext)
                pre.p3 = ldvec(&S.vectors[idx - S.nBlkX + dir]);
        };

        Pre cur, nxt;
        int bx, by, dir;
        blk_of(0, bx, by, dir);
        prefetch(bx, by, dir, cur);
        load_src(bx, by, S.srcw);
        VEC prev = zeroFs; // result of the previous block in scan order
        for (int t = 0; t < nBlk; t++) {
            const int blkIdx = by * S.nBlkX + bx;
            int nbx = 0, nby = 0, ndir = 1;
            const bool hasMore = t + 1 < nBlk;
            if (hasMore) {
                blk_of(t + 1, nbx, nby, ndir);
                prefetch(nbx, nby, ndir, nxt);
                load_src(nbx, nby, S.nextw);
            }
            S.x0 = hp0 + stepX * bx;
            S.y0 = vp0 + stepY * by;
            S.xc = hpc + (stepX >> sd.logxR) * bx;
            S.yc = vpc + (stepY >> sd.logyR) * by;
            S.nLambda = by == 0 ? 0 : nLambdaLevel;
            S.dxMax = (S.pw0 - S.x0 - P.blkW - hp0 + hps) << S.logPel;
            S.dyMax = (S.ph0 - S.y0 - P.blkH - vp0 + vps) << S.logPel;
            S.dxMin = -((S.x0 - hp0 + hps) << S.logPel);
            S.dyMin = -((S.y0 - vp0 + vps) << S.logPel);
            S.predictor = S.clip(cur.hier);
            // pobFetchPredictors, PlaneOfBlocks.cpp:419-463
            VEC pr[4];
            const bool hasPrev = (dir == 1 && bx > 0) || (dir == -1 && bx < S.nBlkX - 1);
            pr[1] = S.clip(hasPrev ? prev : zeroFs);
            pr[2] = S.clip(cur.up);
            pr[3] = S.clip(cur.p3);
            if (by > 0) {
                pr[0].x = max(min(pr[1].x, pr[2].x), min(max(pr[1].x, pr[2].x), pr[3].x));
                pr[0].y = max(min(pr[1].y, pr[2].y), min(max(pr[1].y, pr[2].y), pr[3].y));
                pr[0].sad = max(pr[1].sad, max(pr[2].sad, pr[3].sad));
            } else {
                pr[0] = pr[1];
            }
            if (S.smallest)
                S.predictor = pr[0];
            {
                double scale = __ddiv_rn((double)S.LSAD, (double)(S.LSAD + (S.predictor.sad >> 1)));
                S.nLambda = __double2ll_rz(__dmul_rn(__dmul_rn((double)S.nLambda, scale), scale));
            }
            S.epz(pr, blkIdx);
            if (lane == 0)
                stvec(&S.vectors[blkIdx], S.best);
            __syncwarp(); // orders the vector store before the next blocks' (other lanes') loads
            prev = S.best;
            if (hasMore) {
                // the next block's "up" neighbour is THIS block at a meander turn: its prefetched copy is stale
                if (nby > 0 && (nby * S.nBlkX + nbx) - S.nBlkX == blkIdx)
                    nxt.up = S.best;
                cur = nxt;
#pragma unroll
                for (int k = 0; k < NW; k++)
                    S.srcw[k] = S.nextw[k];
                bx = nbx;
                by = nby;
                dir = ndir;
            }
        }
        __syncwarp();
        // (sumLumaChange / meanLumaChange, :1109-1110,1129-1130, only feed dctweight16: unused for dct=0)
    }
    chain_divide(job, P, lane);
}

} // namespace

// Returns the template selector for a parameter set, or -1 when only the generic kernel applies.
static int fast_config(const SearchParams &P, const SuperDesc &sd, int *nw, int *pack) {
    if (P.dctmode != 0 || P.tryMany)
        return -1;
    auto okType = [](int t) { return t == 3 || t == 4 || t == 5; };
    if (!okType(P.search) || !okType(P.search_coarse))
        return -1;
    const int cw = P.blkW >> sd.logxR, ch = P.blkH >> sd.logyR;
    if ((P.blkW * P.bps) % 4 || (P.chroma && (cw * P.bps) % 4))
        return -1;
    const int words = (P.blkW * P.bps / 4) * P.blkH + (P.chroma ? 2 * (cw * P.bps / 4) * ch : 0);
    const int need = (words + 31) / 32;
    static const int sizes[] = {1, 2, 3, 4, 6, 8};
    *nw = 0;
    for (int s : sizes)
        if (need <= s) {
            *nw = s;
            break;
        }
    if (!*nw)
        return -1;
    *pack = (P.bps == 1 && P.blkW * P.blkH <= 256 && (!P.chroma || 2 * cw * ch <= 256)) ? 1 : 0;
    return (*nw) * 4 + (P.bps == 2 ? 2 : 0) + *pack;
}

bool fast_applies(const SuperDesc &sd, const SearchParams *host_params, int nparams) {
    int key = -2;
    for (int i = 0; i < nparams; i++) {
        int n = 0, p = 0;
        const int k = fast_config(host_params[i], sd, &n, &p);
        if (k < 0 || (key != -2 && k != key))
            return false;
        key = k;
    }
    return key >= 0;
}

int launch_search_fast(Ctx *ctx, const SuperDesc &sd, const SearchParams *host_params, int nparams, const SuperDesc *dev_sd, const SearchParams *dev_params,
                       const SearchJob *dev_jobs, int njobs, int *hist, int hist_ints, bool *handled) {
    *handled = false;
    int key = -2, nw = 0, pack = 0;
    for (int i = 0; i < nparams; i++) {
        int n = 0, p = 0;
        int k = fast_config(host_params[i], sd, &n, &p);
        if (k < 0 || (key != -2 && k != key))
            return 0;
        key = k;
        nw = n;
        pack = p;
    }
    if (key < 0)
        return 0;
    const int bps = host_params[0].bps;
#define LAUNCH(NW_, BPS_, PACK_)                                                                                                   \
    do {                                                                                                                           \
        k_search_fast<NW_, BPS_, PACK_><<<njobs, 32, 0, ctx->stream>>>(dev_sd, dev_params, dev_jobs, njobs, hist, hist_ints); \
        *handled = true;                                                                                                           \
    } while (0)
#ifdef MVB_FAST_QUICK
    if (bps == 1 && pack && nw == 3) LAUNCH(3, 1, true);
#else
    if (bps == 1 && pack) {
        if (nw == 1) LAUNCH(1, 1, true);
        else if (nw == 2) LAUNCH(2, 1, true);
        else if (nw == 3) LAUNCH(3, 1, true);
        else if (nw == 4) LAUNCH(4, 1, true);
    } else if (bps == 1) {
        if (nw == 1) LAUNCH(1, 1, false);
        else if (nw == 2) LAUNCH(2, 1, false);
        else if (nw == 3) LAUNCH(3, 1, false);
        else if (nw == 4) LAUNCH(4, 1, false);
        else if (nw == 6) LAUNCH(6, 1, false);
        else if (nw == 8) LAUNCH(8, 1, false);
    } else {
        if (nw == 1) LAUNCH(1, 2, false);
        else if (nw == 2) LAUNCH(2, 2, false);
        else if (nw == 3) LAUNCH(3, 2, false);
        else if (nw == 4) LAUNCH(4, 2, false);
        else if (nw == 6) LAUNCH(6, 2, false);
        else if (nw == 8) LAUNCH(8, 2, false);
    }
#endif
#undef LAUNCH
    if (*handled) {
        ctx->launches++;
        MVB_CUDA_CHECK(cudaGetLastError());
    }
    return 0;
}

} // namespace mvb200
