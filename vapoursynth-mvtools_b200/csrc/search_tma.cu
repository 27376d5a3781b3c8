// The production search kernel for the default mv.Analyse configurations (dct=0, no
// trymany, search types exhaustive / hex2 / UMH, pel <= 2, compile-time block shapes).
// Same semantics -- bit for bit -- as the generic kernel (search.cu); three ideas make it
// fast on sm_100a:
//
// 1. TMA-staged reference windows.  The hierarchical predictor of a block is known
//    long before the serial chain reaches it, so while block t is being searched, one
//    elected lane issues `cp.async.bulk.tensor.3d` tile loads (UBLKCP/UTMALDG in SASS)
//    that bring a (blk + 2R)^2 window of EVERY pel sub-plane of the three planes around
//    block t+1's predictor from the HBM-resident super clip into this warp's shared
//    memory (double buffered, one mbarrier per buffer).  Candidate blocks are then read
//    with 2 LDS + 1 funnel shift per word instead of gathers that touch one 128-byte
//    line per block row.  Candidates that leave the window (rare: far global/zero
//    predictors, the bad-block wide search) fall back to direct global loads.
// 2. Lane-parallel candidate bookkeeping.  In a pattern step lane j owns candidate j:
//    it generates the coordinates, does the bounds test, the lambda*distance cost, the
//    sub-plane / window arithmetic, and after the SADs are reduced it computes
//    candidate j's total cost.  The reference's sequential strict-'<' accepts over the
//    step are exactly "first minimum wins", so they are replayed as ONE warp arg-min
//    (ties to the lower index) instead of 8 serial compare/update sequences.
// 3. The pixel work of a candidate is spread over the 32 lanes (word granularity,
//    VABSDIFF4.U8.ACC / 2x16-bit), source block in registers, one REDUX.SUM per
//    candidate with luma/chroma packed where they fit 16 bits each.
#include "search_common.cuh"

#include <cuda.h>

#ifndef MVB_TMA_MIN_BLOCKS
#define MVB_TMA_MIN_BLOCKS 16 // caps registers at 128 so that 16 chains fit on an SM
#endif

namespace mvb200 {

namespace {

// ------------------------------------------------------------------- PTX
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *tm, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst), "l"(tm), "r"(c0),
                 "r"(c1), "r"(c2), "r"(bar)
                 : "memory");
}

__constant__ int t_hex2[8][2] = {{-1, -2}, {-2, 0}, {-1, 2}, {1, 2}, {2, 0}, {1, -2}, {-1, -2}, {-2, 0}};
__constant__ int t_hex6[8][2] = {{-2, 0}, {-1, 2}, {1, 2}, {2, 0}, {1, -2}, {-1, -2}, {0, 0}, {0, 0}};
__constant__ int t_hex4[16][2] = {{-4, 2}, {-4, 1}, {-4, 0}, {-4, -1}, {-4, -2}, {4, -2}, {4, -1}, {4, 0},
                                  {4, 1},  {4, 2},  {2, 3},  {0, 4},   {-2, 3},  {-2, -3}, {0, -4}, {2, -3}};

template <int BPS, int BW, int BH, int LXR, int LYR, bool CHROMA>
struct Cfg {
    static constexpr int CW = BW >> LXR, CH = BH >> LYR;
    static constexpr int R = 4; // guaranteed luma window radius in full pixels
    static constexpr int RCX = R >> LXR, RCY = R >> LYR;
    // window row bytes: block + 2R, plus up to 15 bytes because a TMA tile must START on a 16-byte
    // boundary of the innermost dimension (measured: an unaligned start coordinate raises
    // "illegal instruction" on sm_100), rounded up to the 16-byte box granularity
    static constexpr int LWB = ((BW + 2 * R) * BPS + 15 + 15) / 16 * 16;
    static constexpr int LWH = BH + 2 * R;
    static constexpr int CWB = ((CW + 2 * RCX) * BPS + 15 + 15) / 16 * 16;
    static constexpr int CWH = CH + 2 * RCY;
    static constexpr int LT = (LWB * LWH + 127) / 128 * 128; // tile bytes (TMA destinations are 128-byte aligned)
    static constexpr int CT = CHROMA ? (CWB * CWH + 127) / 128 * 128 : 0;
    static constexpr int NSUB = 4; // pel <= 2
    static constexpr int BUF = NSUB * LT + 2 * NSUB * CT;
    static constexpr int LW = BW * BPS / 4, LN = LW * BH, NLI = (LN + 31) / 32;
    static constexpr int CWw = CW * BPS / 4, CN = CWw * CH, NCI = CHROMA ? (2 * CN + 31) / 32 : 0;
    static constexpr bool PACK = BPS == 1 && BW * BH <= 256 && 2 * CW * CH <= 256;
    static constexpr int SMEM = 2 * BUF + 256;
};

struct TmaMaps {
    CUtensorMap plane[3];
};

struct Pre {
    VEC hier, up, p3;
};

template <int BPS, int BW, int BH, int LXR, int LYR, bool CHROMA>
struct Chain {
    using C = Cfg<BPS, BW, BH, LXR, LYR, CHROMA>;
    static constexpr int NLI = C::NLI, NCI = C::NCI;
    static constexpr long long INF = 0x7fffffffffffffffLL;

    // ---- per-lane pixel mapping (constant for the whole chain)
    int lo_l[NLI], lo_c[NCI > 0 ? NCI : 1]; // smem byte offsets of the lane's words inside a window tile
    bool act_l[NLI], act_c[NCI > 0 ? NCI : 1];
    int go_l[NLI], go_c[NCI > 0 ? NCI : 1]; // global byte offsets inside the plane (row * pitch + 4 * col), per level
    int cplane[NCI > 0 ? NCI : 1];          // 1 = U, 2 = V for the lane's chroma words
    unsigned srcL[NLI], srcC[NCI > 0 ? NCI : 1], nextL[NLI], nextC[NCI > 0 ? NCI : 1];
    // ---- level constants
    const uint8_t *sbase[3], *rbase[3]; // level bases (padded origin, sub-plane 0) of source / reference frame
    int pitch0, pitch1, sub0, sub1, hp0, vp0, hpc, vpc, pw0, ph0, row0_0, row0_1, ph_0, ph_1;
    int nBlkX, nBlkY, ovX, ovY, logPel, smallest, searchType, searchParam;
    int penaltyZero, pglobal, penaltyNew, badrange, fieldShift, ref_frame;
    long long badSAD, LSAD;
    VEC *vectors;
    VEC globalMV;
    // ---- block state
    int x0, y0, xc, yc, dxMin, dxMax, dyMin, dyMax;
    int wx0, wy0, wxc0, wyc0; // window origins (full-pel, padded-plane coordinates) of the CURRENT buffer
    long long nLambda, nMinCost;
    VEC best, predictor;
    int badcount, lane;
    uint32_t sbuf; // shared-space address of the current window buffer

    __device__ __forceinline__ bool ok(int vx, int vy) const { return vx >= dxMin && vy >= dyMin && vx < dxMax && vy < dyMax; }
    __device__ __forceinline__ VEC clip(VEC v) const {
        v.x = min(max(v.x, dxMin), dxMax - 1);
        v.y = min(max(v.y, dyMin), dyMax - 1);
        return v;
    }
    // pobMotionDistorsion, PlaneOfBlocks.cpp:105-114
    __device__ __forceinline__ long long mv_cost(int vx, int vy) const {
        int ddx = predictor.x - vx, ddy = predictor.y - vy;
        int dist = (int)((unsigned)(ddx * ddx) + (unsigned)(ddy * ddy));
        return (long long)(int)((nLambda * dist) >> 8);
    }
    // chroma vector of a luma vector: division toward zero (pobGetRefBlockU/V, :56-101)
    __device__ __forceinline__ static int cdiv(int v, int lg) { return (v + ((v < 0) ? ((1 << lg) - 1) : 0)) >> lg; }

    // window reads use 32-bit shared-space addresses (no 64-bit generic pointer arithmetic per word).
    // `volatile` keeps them after the mbarrier wait that publishes the TMA data.
    __device__ __forceinline__ static unsigned lds(uint32_t a) {
        unsigned v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
        return v;
    }
    __device__ __forceinline__ static void lds2(uint32_t a, unsigned &lo, unsigned &hi) {
        asm volatile("ld.shared.u32 %0, [%2];\n\tld.shared.u32 %1, [%2+4];" : "=r"(lo), "=r"(hi) : "r"(a));
    }
    __device__ __forceinline__ static unsigned ldg_un(const uint8_t *p) { // unaligned 32-bit global load
        unsigned long long a = (unsigned long long)p;
        const unsigned *q = (const unsigned *)(a & ~3ull);
        unsigned sh = ((unsigned)a & 3u) * 8u;
        return __funnelshift_r(__ldg(q), __ldg(q + 1), sh);
    }
    __device__ __forceinline__ static unsigned absd(unsigned a, unsigned b) { return BPS == 1 ? __vsadu4(a, b) : __vsadu2(a, b); }

    // window placement for a block at (x0_, y0_)/(xc_, yc_) whose (clipped) centre vector is (cx, cy)
    __device__ __forceinline__ void window_origin(int x0_, int y0_, int xc_, int yc_, int cx, int cy, int &ox, int &oy, int &oxc, int &oyc) const {
        ox = (((x0_ << logPel) + cx) >> logPel) - C::R;
        oy = (((y0_ << logPel) + cy) >> logPel) - C::R;
        oxc = (((xc_ << logPel) + cdiv(cx, LXR)) >> logPel) - C::RCX;
        oyc = (((yc_ << logPel) + cdiv(cy, LYR)) >> logPel) - C::RCY;
        // 16-byte aligned start of the innermost (x) coordinate, rounding toward -infinity
        ox = ((ox * BPS) & ~15) >> (BPS == 2 ? 1 : 0);
        oxc = ((oxc * BPS) & ~15) >> (BPS == 2 ? 1 : 0);
    }

    // one lane issues the tile loads of a window into buffer `dst`
    __device__ __forceinline__ void issue_window(const TmaMaps *maps, uint32_t dst, uint32_t bar, int ox, int oy, int oxc, int oyc) const {
        const int nsub = 1 << (2 * logPel);
        mbar_expect_tx(bar, (uint32_t)(nsub * (C::LWB * C::LWH + (CHROMA ? 2 * C::CWB * C::CWH : 0))));
        for (int s = 0; s < nsub; s++) {
            tma_load_3d(dst + s * C::LT, &maps->plane[0], ox, row0_0 + s * ph_0 + oy, ref_frame, bar);
            if (CHROMA) {
                tma_load_3d(dst + C::NSUB * C::LT + s * C::CT, &maps->plane[1], oxc, row0_1 + s * ph_1 + oyc, ref_frame, bar);
                tma_load_3d(dst + C::NSUB * C::LT + (C::NSUB + s) * C::CT, &maps->plane[2], oxc, row0_1 + s * ph_1 + oyc, ref_frame, bar);
            }
        }
    }

    // ---- SAD of ONE candidate from global memory (fallback; warp-uniform arguments)
    __device__ __forceinline__ void eval_global(int lvx, int lvy, int cvx, int cvy, unsigned &sy, unsigned &suv) const {
        const int m = (1 << logPel) - 1;
        int X = (x0 << logPel) + lvx, Y = (y0 << logPel) + lvy;
        long long o0 = (long long)((X & m) | ((Y & m) << logPel)) * sub0 + (long long)(Y >> logPel) * pitch0 + (X >> logPel) * BPS;
        unsigned aL = 0, aC = 0;
#pragma unroll
        for (int k = 0; k < NLI; k++)
            if (act_l[k])
                aL += absd(ldg_un(rbase[0] + o0 + go_l[k]), srcL[k]);
        if (CHROMA) {
            int Xc = (xc << logPel) + cdiv(cvx, LXR), Yc = (yc << logPel) + cdiv(cvy, LYR);
            long long o1 = (long long)((Xc & m) | ((Yc & m) << logPel)) * sub1 + (long long)(Yc >> logPel) * pitch1 + (Xc >> logPel) * BPS;
#pragma unroll
            for (int k = 0; k < NCI; k++)
                if (act_c[k])
                    aC += absd(ldg_un((cplane[k] == 1 ? rbase[1] : rbase[2]) + o1 + go_c[k]), srcC[k]);
        }
        sy = __reduce_add_sync(0xffffffffu, aL);
        suv = CHROMA ? __reduce_add_sync(0xffffffffu, aC) : 0u;
    }

    // Per-lane description of "my" candidate in a step
    struct Cand {
        int lvx, lvy, cvx, cvy; // luma vector, vector the chroma position derives from
        bool valid;             // in bounds (or exempt from the bounds test)
    };

    // Evaluates the candidates held by lanes 0..7 (lane j: candidate j).  Returns, in every lane,
    // this lane's own candidate's (sadY, sadUV) -- lanes without a valid candidate get garbage.
    template <int NS>
    __device__ __forceinline__ void eval_step(const Cand &c, unsigned &mySy, unsigned &mySuv) const {
        // my candidate's window addressing
        const int m = (1 << logPel) - 1;
        int X = (x0 << logPel) + c.lvx, Y = (y0 << logPel) + c.lvy;
        int fx = X >> logPel, fy = Y >> logPel;
        bool inside = fx >= wx0 && fx + BW <= wx0 + C::LWB / BPS && fy >= wy0 && fy + BH <= wy0 + C::LWH;
        int cbl = ((X & m) | ((Y & m) << logPel)) * C::LT + (fy - wy0) * C::LWB + (fx - wx0) * BPS; // low 2 bits = byte misalignment
        int cbc = C::NSUB * C::LT;
        if (CHROMA) {
            int Xc = (xc << logPel) + cdiv(c.cvx, LXR), Yc = (yc << logPel) + cdiv(c.cvy, LYR);
            int fxc = Xc >> logPel, fyc = Yc >> logPel;
            inside = inside && fxc >= wxc0 && fxc + C::CW <= wxc0 + C::CWB / BPS && fyc >= wyc0 && fyc + C::CH <= wyc0 + C::CWH;
            cbc += ((Xc & m) | ((Yc & m) << logPel)) * C::CT + (fyc - wyc0) * C::CWB + (fxc - wxc0) * BPS;
        }
        const bool mine = lane < NS && c.valid;
        if (!mine) { // harmless in-window dummy
            cbl = 0;
            cbc = C::NSUB * C::LT;
        }
        const unsigned needMask = __ballot_sync(0xffffffffu, mine);
        const unsigned outside = __ballot_sync(0xffffffffu, mine && !inside);
        unsigned res[NS];
        if (outside == 0) {
            unsigned aL[NS], aC[NS];
#pragma unroll
            for (int j = 0; j < NS; j++) {
                const int pl = __shfl_sync(0xffffffffu, cbl, j);
                const uint32_t al = sbuf + (pl & ~3);
                const unsigned shl = (pl & 3) * 8;
                aL[j] = 0;
                aC[j] = 0;
#pragma unroll
                for (int k = 0; k < NLI; k++) {
                    unsigned w0, w1;
                    lds2(al + lo_l[k], w0, w1);
                    unsigned w = __funnelshift_r(w0, w1, shl);
                    unsigned d = absd(w, srcL[k]);
                    aL[j] += act_l[k] ? d : 0u;
                }
                if (CHROMA) {
                    const int pc = __shfl_sync(0xffffffffu, cbc, j);
                    const uint32_t ac = sbuf + (pc & ~3);
                    const unsigned shc = (pc & 3) * 8;
#pragma unroll
                    for (int k = 0; k < NCI; k++) {
                        unsigned w0, w1;
                        lds2(ac + lo_c[k], w0, w1);
                        unsigned w = __funnelshift_r(w0, w1, shc);
                        unsigned d = absd(w, srcC[k]);
                        aC[j] += act_c[k] ? d : 0u;
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < NS; j++) {
                if (C::PACK) {
                    res[j] = __reduce_add_sync(0xffffffffu, aL[j] + (aC[j] << 16));
                } else {
                    res[j] = __reduce_add_sync(0xffffffffu, aL[j]);
                    aC[j] = CHROMA ? __reduce_add_sync(0xffffffffu, aC[j]) : 0u;
                }
            }
            if (C::PACK) {
                unsigned r = res[0];
#pragma unroll
                for (int j = 1; j < NS; j++)
                    r = lane == j ? res[j] : r;
                mySy = r & 0xffffu;
                mySuv = r >> 16;
            } else {
                unsigned r = res[0], u = aC[0];
#pragma unroll
                for (int j = 1; j < NS; j++) {
                    r = lane == j ? res[j] : r;
                    u = lane == j ? aC[j] : u;
                }
                mySy = r;
                mySuv = u;
            }
        } else {
            // at least one candidate leaves the staged window: evaluate the step from global memory
            unsigned ry = 0, ruv = 0;
            for (int j = 0; j < NS; j++) {
                if (!(needMask & (1u << j)))
                    continue;
                unsigned sy, suv;
                eval_global(__shfl_sync(0xffffffffu, c.lvx, j), __shfl_sync(0xffffffffu, c.lvy, j), __shfl_sync(0xffffffffu, c.cvx, j), __shfl_sync(0xffffffffu, c.cvy, j), sy, suv);
                if (lane == j) {
                    ry = sy;
                    ruv = suv;
                }
            }
            mySy = ry;
            mySuv = ruv;
        }
    }

    // Warp arg-min over lanes 0..7 of (cost, lane); ties to the lower lane.  Returns the winner lane.
    __device__ __forceinline__ static int argmin8(long long cost, long long &best_cost) {
        int idx = threadIdx.x & 31;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            long long oc = __shfl_xor_sync(0xffffffffu, cost, o);
            int oi = __shfl_xor_sync(0xffffffffu, idx, o);
            bool take = oc < cost || (oc == cost && oi < idx);
            cost = take ? oc : cost;
            idx = take ? oi : idx;
        }
        best_cost = __shfl_sync(0xffffffffu, cost, 0);
        return __shfl_sync(0xffffffffu, idx, 0);
    }

    // Replays the sequential strict-'<' accepts of one step: first minimum wins.
    // Returns true if a candidate was accepted; winner's data are broadcast.
    __device__ __forceinline__ bool resolve(const Cand &c, long long myCost, unsigned mySy, unsigned mySuv, bool updateBest, int myVal, int &dirOut) {
        long long cmin;
        const int w = argmin8((lane < 8 && c.valid) ? myCost : INF, cmin);
        if (!(cmin < nMinCost))
            return false;
        const int wx = __shfl_sync(0xffffffffu, c.lvx, w), wy = __shfl_sync(0xffffffffu, c.lvy, w);
        const unsigned wsy = __shfl_sync(0xffffffffu, mySy, w), wsuv = __shfl_sync(0xffffffffu, mySuv, w);
        const int wval = __shfl_sync(0xffffffffu, myVal, w);
        if (updateBest) {
            best.x = wx;
            best.y = wy;
        }
        nMinCost = cmin;
        best.sad = (long long)wsy + wsuv;
        dirOut = wval;
        return true;
    }

    __device__ __forceinline__ long long cost_new(int vx, int vy, unsigned sy, unsigned suv) const { // CheckMV / CheckMVdir cost (:228-247)
        return mv_cost(vx, vy) + sy + (long long)(((unsigned long long)(unsigned)penaltyNew * sy) >> 8) + suv + (long long)(((unsigned long long)(unsigned)penaltyNew * suv) >> 8);
    }

    // pobExpandingSearch point generator (:635-658)
    __device__ __forceinline__ static int expand_count(int r, int s) {
        int m = (-r + s >= r) ? 0 : (2 * r - s - 1) / s + 1;
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

    // One pattern step of `count` candidates around (cx, cy); CheckMV (updateBest) or CheckMVdir (dir) semantics.
    template <int PAT>
    __device__ __forceinline__ void run_pattern(int cx, int cy, int count, int a, int b, bool dirMode, int &dir) {
        for (int base = 0; base < count; base += 8) {
            const int i = base + (lane & 7);
            int dx = 0, dy = 0, v = 0;
            if (PAT == PAT_EXPAND) {
                expand_pt(a, b, i, dx, dy);
            } else if (PAT == PAT_HEX6) { // :682-687
                dx = t_hex6[i & 7][0];
                dy = t_hex6[i & 7][1];
                v = i;
            } else if (PAT == PAT_HEX3) { // :706-708, a = odir
                dx = t_hex2[(a + i) & 7][0];
                dy = t_hex2[(a + i) & 7][1];
                v = a - 1 + i;
            } else if (PAT == PAT_CROSS) { // pobCrossSearch :727-739, start 1, x_max = y_max = a
                int per = a / 2;
                int t = i < 2 * per ? i : i - 2 * per;
                int d = 1 + 2 * (t >> 1);
                d = (t & 1) ? d : -d;
                dx = i < 2 * per ? d : 0;
                dy = i < 2 * per ? 0 : d;
            } else { // PAT_HEX4 :755-763, a = scale
                dx = t_hex4[i & 15][0] * a;
                dy = t_hex4[i & 15][1] * a;
            }
            Cand c;
            c.lvx = c.cvx = cx + dx;
            c.lvy = c.cvy = cy + dy;
            c.valid = i < count && ok(c.lvx, c.lvy);
            if (__ballot_sync(0xffffffffu, lane < 8 && c.valid) == 0)
                continue;
            unsigned sy, suv;
            if (PAT == PAT_HEX6)
                eval_step<6>(c, sy, suv);
            else if (PAT == PAT_HEX3)
                eval_step<3>(c, sy, suv);
            else
                eval_step<8>(c, sy, suv);
            int d2 = dir;
            if (resolve(c, cost_new(c.lvx, c.lvy, sy, suv), sy, suv, !dirMode, v, d2) && dirMode)
                dir = d2;
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
                bmx += t_hex2[(dir + 1) & 7][0];
                bmy += t_hex2[(dir + 1) & 7][1];
                for (int i = 1; i < range / 2 && ok(bmx, bmy); i++) {
                    const int odir = (dir + 1 == 0) ? 5 : ((dir + 1 == 7) ? 0 : dir); // mod6m1[dir + 1]
                    dir = -2;
                    run_pattern<PAT_HEX3>(bmx, bmy, 3, odir, 0, true, dir);
                    if (dir == -2)
                        break;
                    bmx += t_hex2[(dir + 1) & 7][0];
                    bmy += t_hex2[(dir + 1) & 7][1];
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

    __device__ __forceinline__ void refine() { // pobRefine :772-816, types 3/4/5
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
        globalMV = clip(globalMV); // sticky (:859)
        // lane j owns start candidate j: 0 zero, 1 global, 2 hierarchical, 3..6 predictors[0..3]
        Cand c;
        const int j = lane & 7;
        int px = 0, py = 0;
        if (j == 1) {
            px = globalMV.x;
            py = globalMV.y;
        } else if (j == 2) {
            px = predictor.x;
            py = predictor.y;
        }
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (j == 3 + i) {
                px = pr[i].x;
                py = pr[i].y;
            }
        c.lvx = c.cvx = px;
        c.lvy = c.cvy = py;
        if (j == 0) { // zero: luma at (0, fieldShift), chroma at (0, 0)  (:836-839)
            c.lvy = fieldShift;
        }
        c.valid = j < 3 || (j < 7 && ok(px, py));
        unsigned sy, suv;
        eval_step<7>(c, sy, suv);
        // per-candidate cost rules (:846, :870, :894, CheckMV0)
        const long long sad = (long long)sy + suv;
        long long cost;
        if (j == 0)
            cost = sad + ((penaltyZero * sad) >> 8);
        else if (j == 1)
            cost = sad + ((pglobal * sad) >> 8);
        else if (j == 2)
            cost = sad;
        else
            cost = mv_cost(px, py) + sad;
        nMinCost = INF; // zero is accepted unconditionally, then strict '<' in order == first minimum
        int dummy = 0;
        resolve(c, cost, sy, suv, true, 0, dummy);
        refine();
        // bad-block wide search (:938-963)
        const long long foundSAD = best.sad;
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
};

template <int BPS, int BW, int BH, int LXR, int LYR, bool CHROMA>
__global__ void __launch_bounds__(32, MVB_TMA_MIN_BLOCKS) k_search_tma(const __grid_constant__ TmaMaps maps, const SuperDesc *__restrict__ sdp, const SearchParams *__restrict__ params,
                                                   const SearchJob *__restrict__ jobs, int njobs, int *__restrict__ hist_all, int hist_ints) {
    using C = Cfg<BPS, BW, BH, LXR, LYR, CHROMA>;
    using CH = Chain<BPS, BW, BH, LXR, LYR, CHROMA>;
    extern __shared__ __align__(128) uint8_t smem_raw[];
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

    // shared memory: [buffer 0][buffer 1][2 mbarriers]
    const uint32_t smem0 = (smem_u32(smem_raw) + 127u) & ~127u;
    const uint32_t bar0 = smem0 + 2 * C::BUF;
    if (lane == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
    }
    fence_proxy_async();
    __syncwarp();
    unsigned phase[2] = {0, 0};

    VEC globalMVec;
    globalMVec.x = 0;
    globalMVec.y = 0;
    globalMVec.sad = -1;

    CH S;
    S.lane = lane;
    S.ovX = P.ovX;
    S.ovY = P.ovY;
    S.penaltyZero = P.pzero;
    S.pglobal = P.global ? P.pglobal : P.pzero;
    S.penaltyNew = P.pnew;
    S.badSAD = P.badSAD;
    S.badrange = P.badrange;
    S.LSAD = P.lsad;
    S.ref_frame = job.ref_index;
    const int stepX = BW - P.ovX, stepY = BH - P.ovY;
    // the lane's words: window offsets are level independent
#pragma unroll
    for (int k = 0; k < C::NLI; k++) {
        int w = lane + 32 * k;
        S.act_l[k] = w < C::LN;
        w = S.act_l[k] ? w : 0;
        S.lo_l[k] = (w / C::LW) * C::LWB + (w % C::LW) * 4;
    }
    if (CHROMA) {
#pragma unroll
        for (int k = 0; k < C::NCI; k++) {
            int w = lane + 32 * k;
            S.act_c[k] = w < 2 * C::CN;
            w = S.act_c[k] ? w : 0;
            int pl = w / C::CN, w2 = w % C::CN;
            S.cplane[k] = 1 + pl;
            S.lo_c[k] = pl * (C::NSUB * C::CT) + (w2 / C::CWw) * C::CWB + (w2 % C::CWw) * 4;
        }
    }

    for (int l = L - 1; l >= 0; l--) {
        S.nBlkX = P.nBlkX[l];
        S.nBlkY = P.nBlkY[l];
        S.logPel = l == 0 ? P.logPel : 0;
        S.smallest = (l == L - 1);
        S.vectors = (VEC *)(job.out + P.level_off[l] + 4);
#pragma unroll
        for (int p = 0; p < 3; p++) {
            const int q = (CHROMA || p == 0) ? p : 0;
            const mvb200_level_geom &g = sd.lv[q][l];
            const long long off = sd.offset[q] + (long long)g.row0 * sd.pitch[q];
            S.sbase[p] = job.src + off;
            S.rbase[p] = job.ref + off;
        }
        {
            const mvb200_level_geom &g0 = sd.lv[0][l];
            const mvb200_level_geom &g1 = sd.lv[CHROMA ? 1 : 0][l];
            S.pitch0 = (int)sd.pitch[0];
            S.pitch1 = (int)sd.pitch[CHROMA ? 1 : 0];
            S.sub0 = S.pitch0 * g0.padded_h;
            S.sub1 = S.pitch1 * g1.padded_h;
            S.hp0 = g0.hpad;
            S.vp0 = g0.vpad;
            S.hpc = g1.hpad;
            S.vpc = g1.vpad;
            S.pw0 = g0.padded_w;
            S.ph0 = g0.padded_h;
            S.row0_0 = g0.row0;
            S.row0_1 = g1.row0;
            S.ph_0 = g0.padded_h;
            S.ph_1 = g1.padded_h;
        }
#pragma unroll
        for (int k = 0; k < C::NLI; k++) {
            int w = S.act_l[k] ? lane + 32 * k : 0;
            S.go_l[k] = (w / C::LW) * S.pitch0 + (w % C::LW) * 4;
        }
        if (CHROMA) {
#pragma unroll
            for (int k = 0; k < C::NCI; k++) {
                int w = S.act_c[k] ? lane + 32 * k : 0;
                int w2 = w % C::CN;
                S.go_c[k] = (w2 / C::CWw) * S.pitch1 + (w2 % C::CWw) * 4;
            }
        }
        int st, stp, fs;
        if (l == L - 1) {
            st = (L == 1) ? P.search : P.search_coarse;
            stp = (L == 1) ? P.nPelSearch : P.nSearchParam;
            fs = (L - 1 == 0) ? job.field_shift : 0;
        } else {
            st = (l == 0) ? P.search : P.search_coarse;
            stp = (l == 0) ? P.nPelSearch : P.nSearchParam;
            const VEC *coarse = (const VEC *)(job.out + P.level_off[l + 1] + 4);
            if (P.global)
                estimate_global_mv(coarse, P.nBlkX[l + 1] * P.nBlkY[l + 1], 8192 * 1 * 2, hist, globalMVec, lane);
            interpolate_prediction(S.vectors, S.nBlkX, S.nBlkY, S.logPel, coarse, P.nBlkX[l + 1], P.nBlkY[l + 1], BW, BH, P.ovX, P.ovY, lane);
            __syncwarp();
            fs = (l == 0) ? job.field_shift : 0;
        }
        // ---- doPobSearchMVs, PlaneOfBlocks.cpp:971-1131
        S.searchType = st;
        S.searchParam = stp;
        S.fieldShift = fs;
        S.globalMV.x = (1 << S.logPel) * globalMVec.x;
        S.globalMV.y = (1 << S.logPel) * globalMVec.y + fs;
        S.globalMV.sad = globalMVec.sad;
        const int nScale = 1 << l;
        int nLambdaLevel = P.nLambda / ((1 << S.logPel) * (1 << S.logPel));
        if (P.plevel == 1)
            nLambdaLevel = nLambdaLevel * nScale;
        else if (P.plevel == 2)
            nLambdaLevel = nLambdaLevel * nScale * nScale;
        S.badcount = 0;
        const int hps = S.hp0 >> l, vps = S.vp0 >> l;
        const int nBlk = S.nBlkX * S.nBlkY;
        VEC zeroFs;
        zeroFs.x = 0;
        zeroFs.y = fs;
        zeroFs.sad = 0;

        auto blk_of = [&](int t, int &bx, int &by, int &dir) {
            by = t / S.nBlkX;
            int i = t - by * S.nBlkX;
            dir = (by % 2 == 0 || P.meander == 0) ? 1 : -1;
            bx = dir == 1 ? i : S.nBlkX - 1 - i;
        };
        auto block_geom = [&](int bx, int by, int &x0, int &y0, int &xc, int &yc, int &dxMin, int &dxMax, int &dyMin, int &dyMax) {
            x0 = S.hp0 + stepX * bx;
            y0 = S.vp0 + stepY * by;
            xc = S.hpc + (stepX >> LXR) * bx;
            yc = S.vpc + (stepY >> LYR) * by;
            dxMax = (S.pw0 - x0 - BW - S.hp0 + hps) << S.logPel;
            dyMax = (S.ph0 - y0 - BH - S.vp0 + vps) << S.logPel;
            dxMin = -((x0 - S.hp0 + hps) << S.logPel);
            dyMin = -((y0 - S.vp0 + vps) << S.logPel);
        };
        auto load_src = [&](int x0, int y0, int xc, int yc, unsigned (&dl)[C::NLI], unsigned (&dc)[C::NCI > 0 ? C::NCI : 1]) {
            const long long o0 = (long long)y0 * S.pitch0 + x0 * BPS;
#pragma unroll
            for (int k = 0; k < C::NLI; k++)
                dl[k] = CH::ldg_un(S.sbase[0] + o0 + S.go_l[k]);
            if (CHROMA) {
                const long long o1 = (long long)yc * S.pitch1 + xc * BPS;
#pragma unroll
                for (int k = 0; k < C::NCI; k++)
                    dc[k] = CH::ldg_un((S.cplane[k] == 1 ? S.sbase[1] : S.sbase[2]) + o1 + S.go_c[k]);
            }
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
        // stages the window of the block at scan position (bx, by) around its clipped hierarchical predictor
        auto stage = [&](int bx, int by, const VEC &hier, int buf, int &ox, int &oy, int &oxc, int &oyc) {
            int x0, y0, xc, yc, a, b, c, d;
            block_geom(bx, by, x0, y0, xc, yc, a, b, c, d);
            int cx = min(max(hier.x, a), b - 1), cy = min(max(hier.y, c), d - 1);
            S.window_origin(x0, y0, xc, yc, cx, cy, ox, oy, oxc, oyc);
            fence_proxy_async(); // earlier generic-proxy reads of this buffer are ordered before the async writes
            __syncwarp();
            if (lane == 0)
                S.issue_window(&maps, smem0 + buf * C::BUF, bar0 + 8 * buf, ox, oy, oxc, oyc);
        };

        Pre cur, nxt;
        int bx, by, dir;
        blk_of(0, bx, by, dir);
        prefetch(bx, by, dir, cur);
        {
            int x0, y0, xc, yc, a, b, c, d;
            block_geom(bx, by, x0, y0, xc, yc, a, b, c, d);
            load_src(x0, y0, xc, yc, S.srcL, S.srcC);
        }
        int wo[2][4]; // window origins of the two buffers
        int cbuf = 0;
        if (!S.smallest)
            stage(bx, by, cur.hier, cbuf, wo[cbuf][0], wo[cbuf][1], wo[cbuf][2], wo[cbuf][3]);
        VEC prev = zeroFs;
        for (int t = 0; t < nBlk; t++) {
            const int blkIdx = by * S.nBlkX + bx;
            int nbx = 0, nby = 0, ndir = 1;
            const bool hasMore = t + 1 < nBlk;
            if (hasMore) {
                blk_of(t + 1, nbx, nby, ndir);
                prefetch(nbx, nby, ndir, nxt);
                int x0, y0, xc, yc, a, b, c, d;
                block_geom(nbx, nby, x0, y0, xc, yc, a, b, c, d);
                load_src(x0, y0, xc, yc, S.nextL, S.nextC);
            }
            block_geom(bx, by, S.x0, S.y0, S.xc, S.yc, S.dxMin, S.dxMax, S.dyMin, S.dyMax);
            S.nLambda = by == 0 ? 0 : nLambdaLevel;
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
            if (S.smallest) {
                S.predictor = pr[0];
                // the window centre depends on the previous block here: stage now (few blocks on this level)
                stage(bx, by, S.predictor, cbuf, wo[cbuf][0], wo[cbuf][1], wo[cbuf][2], wo[cbuf][3]);
            }
            {
                double scale = __ddiv_rn((double)S.LSAD, (double)(S.LSAD + (S.predictor.sad >> 1)));
                S.nLambda = __double2ll_rz(__dmul_rn(__dmul_rn((double)S.nLambda, scale), scale));
            }
            // the window of this block was issued one block ago: wait for it
            mbar_wait(bar0 + 8 * cbuf, phase[cbuf]);
            phase[cbuf] ^= 1;
            S.sbuf = smem0 + cbuf * C::BUF;
            S.wx0 = wo[cbuf][0];
            S.wy0 = wo[cbuf][1];
            S.wxc0 = wo[cbuf][2];
            S.wyc0 = wo[cbuf][3];
            // ... and the next block's window goes into the other buffer while this block is searched
            if (hasMore && !S.smallest)
                stage(nbx, nby, nxt.hier, cbuf ^ 1, wo[cbuf ^ 1][0], wo[cbuf ^ 1][1], wo[cbuf ^ 1][2], wo[cbuf ^ 1][3]);
            S.epz(pr, blkIdx);
            if (lane == 0)
                stvec(&S.vectors[blkIdx], S.best);
            __syncwarp(); // orders the vector store before later loads by other lanes
            prev = S.best;
            if (hasMore) {
                if (nby > 0 && (nby * S.nBlkX + nbx) - S.nBlkX == blkIdx)
                    nxt.up = S.best; // meander turn: the next block's "up" neighbour is this block
                cur = nxt;
#pragma unroll
                for (int k = 0; k < C::NLI; k++)
                    S.srcL[k] = S.nextL[k];
                if (CHROMA) {
#pragma unroll
                    for (int k = 0; k < C::NCI; k++)
                        S.srcC[k] = S.nextC[k];
                }
                bx = nbx;
                by = nby;
                dir = ndir;
                if (!S.smallest)
                    cbuf ^= 1;
            }
        }
        __syncwarp();
    }
    chain_divide(job, P, lane);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

template <int BPS, int BW, int BH, int LXR, int LYR, bool CHROMA>
int launch_cfg(Ctx *ctx, const SuperDesc &sd, const uint8_t *frames_base, int nframes, const SuperDesc *dev_sd, const SearchParams *dev_params, const SearchJob *dev_jobs,
               int njobs, int *hist, int hist_ints, const mvb200_super_geom &geom) {
    using C = Cfg<BPS, BW, BH, LXR, LYR, CHROMA>;
    EncodeTiledFn enc = get_encode_fn();
    if (!enc)
        return set_error("mvtools_b200: cuTensorMapEncodeTiled is not available from this driver");
    TmaMaps maps;
    memset(&maps, 0, sizeof(maps));
    for (int p = 0; p < (CHROMA ? 3 : 1); p++) {
        cuuint64_t dims[3] = {(cuuint64_t)(sd.pitch[p] / BPS), (cuuint64_t)geom.plane_h[p], (cuuint64_t)nframes};
        cuuint64_t strides[2] = {(cuuint64_t)sd.pitch[p], (cuuint64_t)sd.frame_bytes};
        cuuint32_t box[3] = {(cuuint32_t)((p ? C::CWB : C::LWB) / BPS), (cuuint32_t)(p ? C::CWH : C::LWH), 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(&maps.plane[p], BPS == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, (void *)(frames_base + sd.offset[p]), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS)
            return set_error("mvtools_b200: cuTensorMapEncodeTiled failed (%d) for plane %d", (int)r, p);
    }
    auto kern = k_search_tma<BPS, BW, BH, LXR, LYR, CHROMA>;
    static bool configured = false;
    if (!configured) {
        MVB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        MVB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        configured = true;
    }
    kern<<<njobs, 32, C::SMEM, ctx->stream>>>(maps, dev_sd, dev_params, dev_jobs, njobs, hist, hist_ints);
    ctx->launches++;
    MVB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

} // namespace

// Launches the TMA kernel when every parameter set matches one compiled configuration.
int launch_search_tma(Ctx *ctx, const SuperDesc &sd, const mvb200_super_geom &geom, const uint8_t *frames_base, int nframes, const SearchParams *host_params, int nparams,
                      const SuperDesc *dev_sd, const SearchParams *dev_params, const SearchJob *dev_jobs, int njobs, int *hist, int hist_ints, bool *handled) {
    *handled = false;
    if (!frames_base || nframes <= 0)
        return 0;
    const SearchParams &P0 = host_params[0];
    for (int i = 0; i < nparams; i++) {
        const SearchParams &P = host_params[i];
        auto okType = [](int t) { return t == 3 || t == 4 || t == 5; };
        if (P.dctmode != 0 || P.tryMany || !okType(P.search) || !okType(P.search_coarse) || P.pel > 2)
            return 0;
        if (P.blkW != P0.blkW || P.blkH != P0.blkH || P.chroma != P0.chroma || P.bps != P0.bps)
            return 0;
    }
    const int bw = P0.blkW, bh = P0.blkH, bps = P0.bps, lx = sd.logxR, ly = sd.logyR;
    const bool chroma = P0.chroma;
#define TRY(BPS_, BW_, BH_, LX_, LY_, CH_)                                                                                                                       \
    if (!*handled && bps == BPS_ && bw == BW_ && bh == BH_ && chroma == CH_ && (!CH_ || (lx == LX_ && ly == LY_))) {                                             \
        *handled = true;                                                                                                                                          \
        return launch_cfg<BPS_, BW_, BH_, LX_, LY_, CH_>(ctx, sd, frames_base, nframes, dev_sd, dev_params, dev_jobs, njobs, hist, hist_ints, geom);             \
    }
    TRY(1, 16, 16, 1, 1, true)
    TRY(2, 16, 16, 1, 1, true)
    TRY(1, 8, 8, 1, 1, true)
    TRY(2, 8, 8, 1, 1, true)
    TRY(1, 16, 16, 0, 0, false)
    TRY(1, 8, 8, 0, 0, false)
#undef TRY
    return 0;
}

} // namespace mvb200
