// The production search kernel for the default mv.Analyse configurations (dct=0, no
// trymany, search types exhaustive / hex2 / UMH, pel <= 2, compile-time block shapes).
// Same semantics -- bit for bit -- as the generic kernel (search.cu).  One warp owns one
// serial search chain (one (frame, direction, delta) instance); what makes it fast on sm_100a:
//
// 1. TMA-staged reference windows.  The hierarchical predictor of a block is known
//    long before the serial chain reaches it, so while block t is being searched, one
//    elected lane issues three `cp.async.bulk.tensor` tile loads (UTMALDG in SASS; a 4-D map
//    whose third dimension is the pel sub-plane brings all pel^2 sub-planes of a plane with
//    ONE instruction) that bring a (blk + 2R)^2 window of the three planes around block
//    t+1's predictor from the HBM-resident super clip into this warp's shared memory
//    (double buffered, one mbarrier per buffer).  Candidates that leave the window (rare:
//    far global/zero predictors, the bad-block wide search) fall back to global loads.
// 2. A pattern step evaluates up to 8 candidates at once: lane = (candidate j = lane / 4,
//    quarter q = lane % 4).  The four lanes of a candidate split the block's rows
//    (q, q+4, q+8, ...), read the candidate's rows from the window with 32-bit LDS + funnel
//    shifts (any byte alignment), accumulate with VABSDIFF4.U8.ACC against the source block
//    held in registers, and combine with two shuffles.  Every lane then knows its own
//    candidate's SAD and computes that candidate's bounds test, lambda*distance and cost.
// 3. The reference's sequential strict-'<' accepts over a step are exactly "first minimum
//    wins", so a step is resolved by ONE warp arg-min: REDUX.MIN over (cost << 3 | j).
#include "search_common.cuh"

#include <cuda.h>

#ifndef MVB_TMA_MIN_BLOCKS
#define MVB_TMA_MIN_BLOCKS 12 // caps registers at 168: the kernel needs ~170 to stay out of local memory (measured: 128 regs = spills in the block loop, 1.8x slower)
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
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *tm, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst), "l"(tm),
                 "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
                 : "memory");
}
// shared-space loads/stores with 32-bit addresses.  `volatile` keeps the loads after the mbarrier
// wait that publishes the TMA data.
__device__ __forceinline__ void sts32(uint32_t a, unsigned v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
template <int OFF>
__device__ __forceinline__ unsigned lds_o(uint32_t a) {
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF));
    return v;
}
template <int OFF>
__device__ __forceinline__ void lds_o4(uint32_t a, unsigned *o) {
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+%5];" : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]) : "r"(a), "n"(OFF));
}
template <int N, int I = 0>
struct RowLd { // N consecutive words at a 4-byte aligned address, immediate offsets
    __device__ __forceinline__ static void run(uint32_t a, unsigned *o) {
        o[I] = lds_o<4 * I>(a);
        RowLd<N, I + 1>::run(a, o);
    }
};
template <int N>
struct RowLd<N, N> {
    __device__ __forceinline__ static void run(uint32_t, unsigned *) {}
};
template <int N, int I = 0>
struct VecLd { // N (multiple of 4) words at a 16-byte aligned address
    __device__ __forceinline__ static void run(uint32_t a, unsigned *o) {
        lds_o4<4 * I>(a, o + I);
        VecLd<N, I + 4>::run(a, o);
    }
};
template <int N>
struct VecLd<N, N> {
    __device__ __forceinline__ static void run(uint32_t, unsigned *) {}
};
template <int N>
__device__ __forceinline__ void lds_row(uint32_t a, unsigned (&o)[N]) {
    RowLd<N>::run(a, o);
}
template <int N>
__device__ __forceinline__ void lds_vec(uint32_t a, unsigned (&o)[N]) { // N words, address aligned to min(16, 4N) bytes
    static_assert(N == 1 || N == 2 || N % 4 == 0, "row words");
    if constexpr (N == 1) {
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(o[0]) : "r"(a));
    } else if constexpr (N == 2) {
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(o[0]), "=r"(o[1]) : "r"(a));
    } else {
        VecLd<N>::run(a, o);
    }
}

__constant__ int t_hex2[8][2] = {{-1, -2}, {-2, 0}, {-1, 2}, {1, 2}, {2, 0}, {1, -2}, {-1, -2}, {-2, 0}};
__constant__ int t_hex6[8][2] = {{-2, 0}, {-1, 2}, {1, 2}, {2, 0}, {1, -2}, {-1, -2}, {0, 0}, {0, 0}};
__constant__ int t_hex4[16][2] = {{-4, 2}, {-4, 1}, {-4, 0}, {-4, -1}, {-4, -2}, {4, -2}, {4, -1}, {4, 0},
                                  {4, 1},  {4, 2},  {2, 3},  {0, 4},   {-2, 3},  {-2, -3}, {0, -4}, {2, -3}};

constexpr int r128(int v) { return (v + 127) / 128 * 128; }

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
    static constexpr int LT = LWB * LWH;               // bytes of one sub-plane's luma tile
    static constexpr int CT = CHROMA ? CWB * CWH : 0;  // ... chroma tile
    static constexpr int NSUB = 4;                     // pel <= 2
    static constexpr int LREG = r128(NSUB * LT);       // luma region of a buffer (TMA destinations are 128-byte aligned)
    static constexpr int CREG = r128(NSUB * CT);       // U region, V region
    static constexpr int BUF = LREG + 2 * CREG;
    static constexpr int RW = BW * BPS / 4, LR = BH / 4;                  // words per luma row, luma rows per lane
    static constexpr int CRW = CHROMA ? CW * BPS / 4 : 1, CR = CHROMA ? CH / 4 : 1; // same for one chroma plane
    static constexpr int LBYTES = BW * BH * BPS, CBYTES = CHROMA ? CW * CH * BPS : 0;
    static constexpr int SRCW = (LBYTES + 2 * CBYTES) / 4, NSI = (SRCW + 31) / 32; // source block words, per lane for the cooperative load
    static constexpr bool PACK = BPS == 1 && BW * BH <= 256 && 2 * CW * CH <= 256;
    static constexpr bool SRC_SMEM = BPS == 2 && BW * BH >= 256; // too many registers: read source rows from shared memory
    static constexpr int SRCOFF = 2 * BUF, BAROFF = SRCOFF + r128(LBYTES + 2 * CBYTES);
    static constexpr int SMEM = BAROFF + 64 + 128;
    static_assert(BH % 4 == 0 && (BW * BPS) % 4 == 0, "block shape");
    static_assert(!CHROMA || (CH % 4 == 0 && (CW * BPS) % 4 == 0), "chroma block shape");
};

struct TmaMaps {
    CUtensorMap plane[3];  // 3-D {x, row of the whole plane, frame}: levels > 0 and pel 1
    CUtensorMap plane4[3]; // 4-D {x, row, sub-plane, frame} over level 0: pel 2
};

struct Pre {
    VEC hier, up, p3;
};

struct Geom {
    int x0, y0, xc, yc, dxMin, dxMax, dyMin, dyMax;
};

template <int BPS, int BW, int BH, int LXR, int LYR, bool CHROMA>
struct Chain {
    using C = Cfg<BPS, BW, BH, LXR, LYR, CHROMA>;
    static constexpr long long INF = 0x7fffffffffffffffLL;
    static constexpr unsigned FULL = 0xffffffffu;

    // ---- source block rows of this lane (rows q, q+4, ...), unless C::SRC_SMEM
    unsigned srcL[C::SRC_SMEM ? 1 : C::LR][C::RW], srcC[2][C::SRC_SMEM ? 1 : C::CR][C::CRW];
    uint32_t srcbuf; // shared-space address of the source block
    // ---- level constants
    const uint8_t *rbase[3]; // level bases (padded origin, sub-plane 0) of the reference frame
    int pitch0, pitch1, sub0, sub1;
    int logPel, searchType, searchParam;
    int penaltyZero, pglobal, penaltyNew, badrange, fieldShift;
    long long badSAD, LSAD;
    VEC globalMV;
    // ---- block state
    Geom g;
    int wx0, wy0, wxc0, wyc0; // window origins (full-pel, padded-plane coordinates) of the CURRENT buffer
    long long nLambda, nMinCost;
    VEC best, predictor;
    int badcount, lane, j, q;
    uint32_t sbuf; // shared-space address of the current window buffer

    __device__ __forceinline__ bool ok(int vx, int vy) const { return vx >= g.dxMin && vy >= g.dyMin && vx < g.dxMax && vy < g.dyMax; }
    __device__ __forceinline__ VEC clip(VEC v) const {
        v.x = min(max(v.x, g.dxMin), g.dxMax - 1);
        v.y = min(max(v.y, g.dyMin), g.dyMax - 1);
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

    __device__ __forceinline__ static unsigned ldg_un(const uint8_t *p) { // unaligned 32-bit global load
        unsigned long long a = (unsigned long long)p;
        const unsigned *w = (const unsigned *)(a & ~3ull);
        unsigned sh = ((unsigned)a & 3u) * 8u;
        return __funnelshift_r(__ldg(w), __ldg(w + 1), sh);
    }
    __device__ __forceinline__ static unsigned absd(unsigned a, unsigned b) { return BPS == 1 ? __vsadu4(a, b) : __vsadu2(a, b); }

    // window placement for a block whose (clipped) centre vector is (cx, cy)
    __device__ __forceinline__ void window_origin(const Geom &gg, int cx, int cy, int &ox, int &oy, int &oxc, int &oyc) const {
        ox = (((gg.x0 << logPel) + cx) >> logPel) - C::R;
        oy = (((gg.y0 << logPel) + cy) >> logPel) - C::R;
        oxc = (((gg.xc << logPel) + cdiv(cx, LXR)) >> logPel) - C::RCX;
        oyc = (((gg.yc << logPel) + cdiv(cy, LYR)) >> logPel) - C::RCY;
        // 16-byte aligned start of the innermost (x) coordinate, rounding toward -infinity
        ox = ((ox * BPS) & ~15) >> (BPS == 2 ? 1 : 0);
        oxc = ((oxc * BPS) & ~15) >> (BPS == 2 ? 1 : 0);
    }

    // Per-lane description of "my" candidate in a step
    struct Cand {
        int lvx, lvy, cvx, cvy; // luma vector, vector the chroma position derives from
        bool valid;             // in bounds (or exempt from the bounds test)
    };

    // Evaluates the candidates of a step: the four lanes 4j..4j+3 hold candidate j (identical Cand).
    // Returns, in every lane, its own candidate's (sadY, sadUV) -- lanes without a valid candidate get garbage.
    template <int NS>
    __device__ __forceinline__ void eval_step(const Cand &c, unsigned &mySy, unsigned &mySuv) const {
        const bool mine = j < NS && c.valid;
        const int lvx = mine ? c.lvx : 0, lvy = mine ? c.lvy : 0, cvx = mine ? c.cvx : 0, cvy = mine ? c.cvy : 0;
        const int m = (1 << logPel) - 1;
        const int X = (g.x0 << logPel) + lvx, Y = (g.y0 << logPel) + lvy;
        const int fx = X >> logPel, fy = Y >> logPel;
        const int subl = (X & m) | ((Y & m) << logPel);
        bool inside = fx >= wx0 && fx + BW <= wx0 + C::LWB / BPS && fy >= wy0 && fy + BH <= wy0 + C::LWH;
        int cbl = subl * C::LT + (fy - wy0) * C::LWB + (fx - wx0) * BPS; // low 2 bits = byte misalignment
        int cbc = 0, fxc = 0, fyc = 0, subc = 0;
        if (CHROMA) {
            const int Xc = (g.xc << logPel) + cdiv(cvx, LXR), Yc = (g.yc << logPel) + cdiv(cvy, LYR);
            fxc = Xc >> logPel;
            fyc = Yc >> logPel;
            subc = (Xc & m) | ((Yc & m) << logPel);
            inside = inside && fxc >= wxc0 && fxc + C::CW <= wxc0 + C::CWB / BPS && fyc >= wyc0 && fyc + C::CH <= wyc0 + C::CWH;
            cbc = subc * C::CT + (fyc - wyc0) * C::CWB + (fxc - wxc0) * BPS;
        }
        const unsigned outside = __ballot_sync(FULL, mine && !inside);
        unsigned aL = 0, aC = 0;
        if (outside == 0) {
            if (!mine) { // harmless in-window dummy
                cbl = 0;
                cbc = 0;
            }
            const uint32_t al = sbuf + (cbl & ~3) + q * C::LWB;
            const unsigned shl = (unsigned)cbl << 3; // funnel shifts use the amount mod 32
#pragma unroll
            for (int i = 0; i < C::LR; i++) {
                unsigned r[C::RW + 1], sw[C::RW];
                lds_row(al + 4 * i * C::LWB, r);
                if (C::SRC_SMEM)
                    lds_vec(srcbuf + (q + 4 * i) * (C::RW * 4), sw);
#pragma unroll
                for (int w = 0; w < C::RW; w++)
                    aL += absd(__funnelshift_r(r[w], r[w + 1], shl), C::SRC_SMEM ? sw[w] : srcL[i][w]);
            }
            if (CHROMA) {
                const uint32_t ac = sbuf + C::LREG + (cbc & ~3) + q * C::CWB;
                const unsigned shc = (unsigned)cbc << 3;
#pragma unroll
                for (int p = 0; p < 2; p++) {
#pragma unroll
                    for (int i = 0; i < C::CR; i++) {
                        unsigned r[C::CRW + 1], sw[C::CRW];
                        lds_row(ac + p * C::CREG + 4 * i * C::CWB, r);
                        if (C::SRC_SMEM)
                            lds_vec(srcbuf + C::LBYTES + p * C::CBYTES + (q + 4 * i) * (C::CRW * 4), sw);
#pragma unroll
                        for (int w = 0; w < C::CRW; w++)
                            aC += absd(__funnelshift_r(r[w], r[w + 1], shc), C::SRC_SMEM ? sw[w] : srcC[p][i][w]);
                    }
                }
            }
        } else {
            // at least one candidate leaves the staged window: the whole step reads global memory
            const uint8_t *pl = rbase[0] + (long long)subl * sub0 + (long long)(fy + q) * pitch0 + fx * BPS;
#pragma unroll
            for (int i = 0; i < C::LR; i++) {
                unsigned sw[C::RW];
                if (C::SRC_SMEM)
                    lds_vec(srcbuf + (q + 4 * i) * (C::RW * 4), sw);
#pragma unroll
                for (int w = 0; w < C::RW; w++)
                    aL += absd(ldg_un(pl + (long long)(4 * i) * pitch0 + 4 * w), C::SRC_SMEM ? sw[w] : srcL[i][w]);
            }
            if (CHROMA) {
                const long long oc = (long long)subc * sub1 + (long long)(fyc + q) * pitch1 + fxc * BPS;
#pragma unroll
                for (int p = 0; p < 2; p++) {
#pragma unroll
                    for (int i = 0; i < C::CR; i++) {
                        unsigned sw[C::CRW];
                        if (C::SRC_SMEM)
                            lds_vec(srcbuf + C::LBYTES + p * C::CBYTES + (q + 4 * i) * (C::CRW * 4), sw);
#pragma unroll
                        for (int w = 0; w < C::CRW; w++)
                            aC += absd(ldg_un(rbase[1 + p] + oc + (long long)(4 * i) * pitch1 + 4 * w), C::SRC_SMEM ? sw[w] : srcC[p][i][w]);
                    }
                }
            }
        }
        if (C::PACK) {
            unsigned v = aL + (aC << 16);
            v += __shfl_xor_sync(FULL, v, 1);
            v += __shfl_xor_sync(FULL, v, 2);
            mySy = v & 0xffffu;
            mySuv = v >> 16;
        } else {
            aL += __shfl_xor_sync(FULL, aL, 1);
            aL += __shfl_xor_sync(FULL, aL, 2);
            if (CHROMA) {
                aC += __shfl_xor_sync(FULL, aC, 1);
                aC += __shfl_xor_sync(FULL, aC, 2);
            }
            mySy = aL;
            mySuv = aC;
        }
    }

    // Warp arg-min of (cost, candidate slot) over the lanes whose `valid` is set; ties to the lower slot.
    // Returns false when no lane is valid.
    __device__ __forceinline__ bool argmin(bool valid, long long cost, long long &best_cost, int &slot) const {
        const bool small = cost >= 0 && cost < (1ll << 28);
        if (__all_sync(FULL, small || !valid)) { // the usual case: one REDUX.MIN
            const unsigned key = valid ? (((unsigned)cost << 3) | (unsigned)j) : 0xffffffffu;
            const unsigned kmin = __reduce_min_sync(FULL, key);
            best_cost = (long long)(kmin >> 3);
            slot = (int)(kmin & 7u);
            return kmin != 0xffffffffu;
        }
        long long cst = valid ? cost : INF;
        int idx = valid ? j : 8;
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
            const long long oc = __shfl_xor_sync(FULL, cst, o);
            const int oi = __shfl_xor_sync(FULL, idx, o);
            const bool take = oc < cst || (oc == cst && oi < idx);
            cst = take ? oc : cst;
            idx = take ? oi : idx;
        }
        best_cost = cst;
        slot = idx;
        return idx < 8;
    }

    // Replays the sequential strict-'<' accepts of one step: first minimum wins.
    // Returns true if a candidate was accepted; the winner's data are broadcast.
    template <int NS>
    __device__ __forceinline__ bool resolve(const Cand &c, long long myCost, unsigned mySy, unsigned mySuv, bool updateBest, int myVal, int &dirOut) {
        long long cmin;
        int w;
        if (!argmin(j < NS && c.valid, myCost, cmin, w))
            return false;
        if (!(cmin < nMinCost))
            return false;
        const int src = w << 2;
        const int wxy = __shfl_sync(FULL, (c.lvx & 0xffff) | (c.lvy << 16), src); // |v| < 2^15 (checked at launch)
        unsigned wsy, wsuv;
        if (C::PACK) {
            const unsigned ws = __shfl_sync(FULL, mySy | (mySuv << 16), src);
            wsy = ws & 0xffffu;
            wsuv = ws >> 16;
        } else {
            wsy = __shfl_sync(FULL, mySy, src);
            wsuv = __shfl_sync(FULL, mySuv, src);
        }
        const int wval = __shfl_sync(FULL, myVal, src);
        if (updateBest) {
            best.x = (int)(short)(wxy & 0xffff);
            best.y = wxy >> 16;
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

    enum { PAT_EXPAND = 0, PAT_HEX6 = 1, PAT_HEX3 = 2, PAT_CROSS = 3, PAT_HEX4 = 4, PAT_SQUARE1 = 5 };

    // One pattern step of `count` candidates around (cx, cy); CheckMV (updateBest) or CheckMVdir (dir) semantics.
    template <int PAT>
    __device__ __forceinline__ void run_pattern(int cx, int cy, int count, int a, int b, bool dirMode, int &dir) {
        for (int base = 0; base < count; base += 8) {
            const int i = base + j;
            int dx = 0, dy = 0, v = 0;
            if (PAT == PAT_EXPAND) {
                expand_pt(a, b, i, dx, dy);
            } else if (PAT == PAT_SQUARE1) { // pobExpandingSearch(1, 1): the generator with compile-time radius/step
                expand_pt(1, 1, i, dx, dy);
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
            if (__ballot_sync(FULL, c.valid) == 0)
                continue;
            unsigned sy, suv;
            int d2 = dir;
            bool acc;
            if (PAT == PAT_HEX6) {
                eval_step<6>(c, sy, suv);
                acc = resolve<6>(c, cost_new(c.lvx, c.lvy, sy, suv), sy, suv, !dirMode, v, d2);
            } else if (PAT == PAT_HEX3) {
                eval_step<3>(c, sy, suv);
                acc = resolve<3>(c, cost_new(c.lvx, c.lvy, sy, suv), sy, suv, !dirMode, v, d2);
            } else {
                eval_step<8>(c, sy, suv);
                acc = resolve<8>(c, cost_new(c.lvx, c.lvy, sy, suv), sy, suv, !dirMode, v, d2);
            }
            if (acc && dirMode)
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
        int d0 = 0;
        run_pattern<PAT_SQUARE1>(bmx, bmy, 8, 1, 1, false, d0); // pobExpandingSearch(1, 1, bmx, bmy)
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
        // slot j owns start candidate j: 0 zero, 1 global, 2 hierarchical, 3..6 predictors[0..3]
        Cand c;
        int px = 0, py = 0;
        px = j == 1 ? globalMV.x : px;
        py = j == 1 ? globalMV.y : py;
        px = j == 2 ? predictor.x : px;
        py = j == 2 ? predictor.y : py;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            px = j == 3 + i ? pr[i].x : px;
            py = j == 3 + i ? pr[i].y : py;
        }
        c.lvx = c.cvx = px;
        c.cvy = py;
        c.lvy = j == 0 ? fieldShift : py; // zero: luma at (0, fieldShift), chroma at (0, 0)  (:836-839)
        c.valid = j < 3 || (j < 7 && ok(px, py));
        unsigned sy, suv;
        eval_step<7>(c, sy, suv);
        // per-candidate cost rules (:846, :870, :894, CheckMV0)
        const long long sad = (long long)sy + suv;
        const int pen = j == 0 ? penaltyZero : (j == 1 ? pglobal : 0);
        const long long cost = sad + ((pen * sad) >> 8) + (j >= 3 ? mv_cost(px, py) : 0ll);
        nMinCost = INF; // zero is accepted unconditionally, then strict '<' in order == first minimum
        int dummy = 0;
        resolve<7>(c, cost, sy, suv, true, 0, dummy);
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

    // shared memory: [window buffer 0][window buffer 1][source block][2 mbarriers]
    const uint32_t smem0 = (smem_u32(smem_raw) + 127u) & ~127u;
    const uint32_t bar0 = smem0 + C::BAROFF;
    if (lane == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
    }
    fence_proxy_async();
    __syncwarp();
    unsigned phasebits = 0; // bit b = parity to wait for on buffer b's mbarrier

    VEC globalMVec;
    globalMVec.x = 0;
    globalMVec.y = 0;
    globalMVec.sad = -1;

    CH S;
    S.lane = lane;
    S.j = lane >> 2;
    S.q = lane & 3;
    S.srcbuf = smem0 + C::SRCOFF;
    S.penaltyZero = P.pzero;
    S.pglobal = P.global ? P.pglobal : P.pzero;
    S.penaltyNew = P.pnew;
    S.badSAD = P.badSAD;
    S.badrange = P.badrange;
    S.LSAD = P.lsad;
    const int ref_frame = job.ref_index;
    const int stepX = BW - P.ovX, stepY = BH - P.ovY;
    const int meander = P.meander;

    // cooperative source-block load: word w of the block (luma rows, then U rows, then V rows) -> lane w % 32
    int sw_plane[C::NSI], sw_row[C::NSI], sw_col[C::NSI];
#pragma unroll
    for (int k = 0; k < C::NSI; k++) {
        int w = lane + 32 * k;
        w = w < C::SRCW ? w : 0;
        const int lw = C::LBYTES / 4, cw = C::CBYTES / 4;
        if (w < lw) {
            sw_plane[k] = 0;
            sw_row[k] = w / C::RW;
            sw_col[k] = (w % C::RW) * 4;
        } else {
            const int w2 = w - lw;
            sw_plane[k] = (CHROMA && w2 >= cw) ? 2 : 1;
            const int w3 = (CHROMA && w2 >= cw) ? w2 - cw : w2;
            sw_row[k] = w3 / C::CRW;
            sw_col[k] = (w3 % C::CRW) * 4;
        }
    }

    for (int l = L - 1; l >= 0; l--) {
        const int nBlkX = P.nBlkX[l], nBlkY = P.nBlkY[l];
        S.logPel = l == 0 ? P.logPel : 0;
        const bool smallest = (l == L - 1);
        VEC *const vectors = (VEC *)(job.out + P.level_off[l] + 4);
        const uint8_t *sbase[3];
#pragma unroll
        for (int p = 0; p < 3; p++) {
            const int qp = (CHROMA || p == 0) ? p : 0;
            const mvb200_level_geom &gl = sd.lv[qp][l];
            const long long off = sd.offset[qp] + (long long)gl.row0 * sd.pitch[qp];
            sbase[p] = job.src + off;
            S.rbase[p] = job.ref + off;
        }
        const mvb200_level_geom &g0 = sd.lv[0][l];
        const mvb200_level_geom &g1 = sd.lv[CHROMA ? 1 : 0][l];
        S.pitch0 = (int)sd.pitch[0];
        S.pitch1 = (int)sd.pitch[CHROMA ? 1 : 0];
        S.sub0 = S.pitch0 * g0.padded_h;
        S.sub1 = S.pitch1 * g1.padded_h;
        const int hp0 = g0.hpad, vp0 = g0.vpad, hpc = g1.hpad, vpc = g1.vpad, pw0 = g0.padded_w, ph0 = g0.padded_h;
        const int row0_0 = g0.row0, row0_1 = g1.row0;
        const bool use4d = S.logPel == 1; // level 0 of a pel-2 clip: the 4-D maps (their row origin is level 0's row0 = 0)
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
            interpolate_prediction(vectors, nBlkX, nBlkY, S.logPel, coarse, P.nBlkX[l + 1], P.nBlkY[l + 1], BW, BH, P.ovX, P.ovY, lane);
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
        const int hps = hp0 >> l, vps = vp0 >> l;
        const int nBlk = nBlkX * nBlkY;
        VEC zeroFs;
        zeroFs.x = 0;
        zeroFs.y = fs;
        zeroFs.sad = 0;

        auto block_geom = [&](int bx, int by, Geom &gg) {
            gg.x0 = hp0 + stepX * bx;
            gg.y0 = vp0 + stepY * by;
            gg.xc = hpc + (stepX >> LXR) * bx;
            gg.yc = vpc + (stepY >> LYR) * by;
            gg.dxMax = (pw0 - gg.x0 - BW - hp0 + hps) << S.logPel;
            gg.dyMax = (ph0 - gg.y0 - BH - vp0 + vps) << S.logPel;
            gg.dxMin = -((gg.x0 - hp0 + hps) << S.logPel);
            gg.dyMin = -((gg.y0 - vp0 + vps) << S.logPel);
        };
        // raw (unshifted) words of the source block at gg; the funnel shift happens when they are consumed
        auto load_src_raw = [&](const Geom &gg, unsigned (&lo)[C::NSI], unsigned (&hi)[C::NSI]) {
#pragma unroll
            for (int k = 0; k < C::NSI; k++) {
                const int pl = sw_plane[k];
                const uint8_t *p = (pl == 0 ? sbase[0] : (pl == 1 ? sbase[1] : sbase[2])) +
                                   (long long)((pl == 0 ? gg.y0 : gg.yc) + sw_row[k]) * (pl == 0 ? S.pitch0 : S.pitch1) + (pl == 0 ? gg.x0 : gg.xc) * BPS + sw_col[k];
                const unsigned *w = (const unsigned *)((unsigned long long)p & ~3ull);
                lo[k] = __ldg(w);
                hi[k] = __ldg(w + 1);
            }
        };
        // shifts the raw words, publishes the block in shared memory and loads this lane's rows
        auto commit_src = [&](const Geom &gg, const unsigned (&lo)[C::NSI], const unsigned (&hi)[C::NSI]) {
            const unsigned shl = ((unsigned)(gg.x0 * BPS) & 3u) * 8u, shc = ((unsigned)(gg.xc * BPS) & 3u) * 8u;
            __syncwarp(); // every lane is done reading the previous block from shared memory
#pragma unroll
            for (int k = 0; k < C::NSI; k++)
                if (lane + 32 * k < C::SRCW)
                    sts32(S.srcbuf + 4 * (lane + 32 * k), __funnelshift_r(lo[k], hi[k], sw_plane[k] == 0 ? shl : shc));
            __syncwarp();
            if (!C::SRC_SMEM) {
#pragma unroll
                for (int i = 0; i < C::LR; i++)
                    lds_vec(S.srcbuf + (S.q + 4 * i) * (C::RW * 4), S.srcL[i]);
                if (CHROMA) {
#pragma unroll
                    for (int p = 0; p < 2; p++)
#pragma unroll
                        for (int i = 0; i < C::CR; i++)
                            lds_vec(S.srcbuf + C::LBYTES + p * C::CBYTES + (S.q + 4 * i) * (C::CRW * 4), S.srcC[p][i]);
                }
            }
        };
        auto prefetch = [&](int bx, int by, int dir, Pre &pre) {
            const int idx = by * nBlkX + bx;
            pre.hier = ldvec(&vectors[idx]);
            const bool hasNext = (dir == 1 && bx < nBlkX - 1) || (dir == -1 && bx > 0);
            pre.up = zeroFs;
            if (by > 0)
                pre.up = ldvec(&vectors[idx - nBlkX]);
            pre.p3 = zeroFs;
            if (by < nBlkY - 1 && hasNext)
                pre.p3 = ldvec(&vectors[idx + nBlkX + dir]);
            else if (by > 0 && hasNext)
                pre.p3 = ldvec(&vectors[idx - nBlkX + dir]);
        };
        // stages the window of the block with geometry gg around its clipped hierarchical predictor
        auto stage = [&](const Geom &gg, const VEC &hier, int buf, int (&wo)[4]) {
            const int cx = min(max(hier.x, gg.dxMin), gg.dxMax - 1), cy = min(max(hier.y, gg.dyMin), gg.dyMax - 1);
            S.window_origin(gg, cx, cy, wo[0], wo[1], wo[2], wo[3]);
            fence_proxy_async(); // earlier generic-proxy reads of this buffer are ordered before the async writes
            __syncwarp();
            if (lane == 0) {
                const uint32_t dst = smem0 + buf * C::BUF, bar = bar0 + 8 * buf;
                if (use4d) {
                    mbar_expect_tx(bar, (uint32_t)(4 * (C::LT + 2 * C::CT)));
                    tma_load_4d(dst, &maps.plane4[0], wo[0], wo[1], 0, ref_frame, bar);
                    if (CHROMA) {
                        tma_load_4d(dst + C::LREG, &maps.plane4[1], wo[2], wo[3], 0, ref_frame, bar);
                        tma_load_4d(dst + C::LREG + C::CREG, &maps.plane4[2], wo[2], wo[3], 0, ref_frame, bar);
                    }
                } else {
                    mbar_expect_tx(bar, (uint32_t)(C::LT + 2 * C::CT));
                    tma_load_3d(dst, &maps.plane[0], wo[0], row0_0 + wo[1], ref_frame, bar);
                    if (CHROMA) {
                        tma_load_3d(dst + C::LREG, &maps.plane[1], wo[2], row0_1 + wo[3], ref_frame, bar);
                        tma_load_3d(dst + C::LREG + C::CREG, &maps.plane[2], wo[2], row0_1 + wo[3], ref_frame, bar);
                    }
                }
            }
        };

        Pre cur, nxt;
        Geom gn;
        int bx = 0, by = 0, dir = 1;
        prefetch(bx, by, dir, cur);
        block_geom(bx, by, S.g);
        {
            unsigned lo[C::NSI], hi[C::NSI];
            load_src_raw(S.g, lo, hi);
            commit_src(S.g, lo, hi);
        }
        int wcur[4], wnxt[4]; // window origins of the current / the other buffer
        int cbuf = 0;
        if (!smallest)
            stage(S.g, cur.hier, cbuf, wcur);
        VEC prev = zeroFs;
        for (int t = 0; t < nBlk; t++) {
            const int blkIdx = by * nBlkX + bx;
            // next scan position (meander order, :1037-1055)
            int nbx = bx + dir, nby = by, ndir = dir;
            if (nbx < 0 || nbx >= nBlkX) {
                nby = by + 1;
                ndir = (meander == 0 || (nby & 1) == 0) ? 1 : -1;
                nbx = ndir == 1 ? 0 : nBlkX - 1;
            }
            const bool hasMore = t + 1 < nBlk;
            unsigned nlo[C::NSI], nhi[C::NSI];
            if (hasMore) {
                prefetch(nbx, nby, ndir, nxt);
                block_geom(nbx, nby, gn);
                load_src_raw(gn, nlo, nhi);
            }
            S.nLambda = by == 0 ? 0 : nLambdaLevel;
            S.predictor = S.clip(cur.hier);
            // pobFetchPredictors, PlaneOfBlocks.cpp:419-463
            VEC pr[4];
            const bool hasPrev = (dir == 1 && bx > 0) || (dir == -1 && bx < nBlkX - 1);
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
            if (smallest) {
                S.predictor = pr[0];
                // the window centre depends on the previous block here: stage now (few blocks on this level)
                stage(S.g, S.predictor, cbuf, wcur);
            }
            {
                double scale = __ddiv_rn((double)S.LSAD, (double)(S.LSAD + (S.predictor.sad >> 1)));
                S.nLambda = __double2ll_rz(__dmul_rn(__dmul_rn((double)S.nLambda, scale), scale));
            }
            // the window of this block was issued one block ago: wait for it
            mbar_wait(bar0 + 8 * cbuf, (phasebits >> cbuf) & 1u);
            phasebits ^= 1u << cbuf;
            S.sbuf = smem0 + cbuf * C::BUF;
            S.wx0 = wcur[0];
            S.wy0 = wcur[1];
            S.wxc0 = wcur[2];
            S.wyc0 = wcur[3];
            // ... and the next block's window goes into the other buffer while this block is searched
            if (hasMore && !smallest)
                stage(gn, nxt.hier, cbuf ^ 1, wnxt);
            S.epz(pr, blkIdx);
            if (lane == 0)
                stvec(&vectors[blkIdx], S.best);
            __syncwarp(); // orders the vector store before later loads by other lanes
            prev = S.best;
            if (hasMore) {
                if (nby > 0 && (nby * nBlkX + nbx) - nBlkX == blkIdx)
                    nxt.up = S.best; // meander turn: the next block's "up" neighbour is this block
                cur = nxt;
                commit_src(gn, nlo, nhi);
                S.g = gn;
                bx = nbx;
                by = nby;
                dir = ndir;
                if (!smallest) {
                    cbuf ^= 1;
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        wcur[k] = wnxt[k];
                }
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
               int njobs, int *hist, int hist_ints, const mvb200_super_geom &geom, int pel) {
    using C = Cfg<BPS, BW, BH, LXR, LYR, CHROMA>;
    EncodeTiledFn enc = get_encode_fn();
    if (!enc)
        return set_error("mvtools_b200: cuTensorMapEncodeTiled is not available from this driver");
    TmaMaps maps;
    memset(&maps, 0, sizeof(maps));
    const CUtensorMapDataType dt = BPS == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_UINT16;
    for (int p = 0; p < (CHROMA ? 3 : 1); p++) {
        const cuuint32_t bw = (cuuint32_t)((p ? C::CWB : C::LWB) / BPS), bh = (cuuint32_t)(p ? C::CWH : C::LWH);
        {
            cuuint64_t dims[3] = {(cuuint64_t)(sd.pitch[p] / BPS), (cuuint64_t)geom.plane_h[p], (cuuint64_t)nframes};
            cuuint64_t strides[2] = {(cuuint64_t)sd.pitch[p], (cuuint64_t)sd.frame_bytes};
            cuuint32_t box[3] = {bw, bh, 1};
            cuuint32_t estr[3] = {1, 1, 1};
            CUresult r = enc(&maps.plane[p], dt, 3, (void *)(frames_base + sd.offset[p]), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS)
                return set_error("mvtools_b200: cuTensorMapEncodeTiled failed (%d) for plane %d", (int)r, p);
        }
        if (pel == 2) { // level 0 as {x, row, sub-plane, frame}: one tile load brings the window of all four sub-planes
            const mvb200_level_geom &g0 = sd.lv[p][0];
            cuuint64_t dims[4] = {(cuuint64_t)(sd.pitch[p] / BPS), (cuuint64_t)g0.padded_h, 4, (cuuint64_t)nframes};
            cuuint64_t strides[3] = {(cuuint64_t)sd.pitch[p], (cuuint64_t)(sd.pitch[p] * g0.padded_h), (cuuint64_t)sd.frame_bytes};
            cuuint32_t box[4] = {bw, bh, 4, 1};
            cuuint32_t estr[4] = {1, 1, 1, 1};
            CUresult r = enc(&maps.plane4[p], dt, 4, (void *)(frames_base + sd.offset[p] + (long long)g0.row0 * sd.pitch[p]), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS)
                return set_error("mvtools_b200: cuTensorMapEncodeTiled (4-D) failed (%d) for plane %d", (int)r, p);
        }
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
        if (P.blkW != P0.blkW || P.blkH != P0.blkH || P.chroma != P0.chroma || P.bps != P0.bps || P.pel != P0.pel)
            return 0;
    }
    // the winner's vector travels as two 16-bit halves of one shuffle
    if ((sd.lv[0][0].padded_w << P0.logPel) >= 32768 || (sd.lv[0][0].padded_h << P0.logPel) >= 32768)
        return 0;
    const int bw = P0.blkW, bh = P0.blkH, bps = P0.bps, lx = sd.logxR, ly = sd.logyR;
    const bool chroma = P0.chroma;
#define TRY(BPS_, BW_, BH_, LX_, LY_, CH_)                                                                                                                       \
    if (!*handled && bps == BPS_ && bw == BW_ && bh == BH_ && chroma == CH_ && (!CH_ || (lx == LX_ && ly == LY_))) {                                             \
        *handled = true;                                                                                                                                          \
        return launch_cfg<BPS_, BW_, BH_, LX_, LY_, CH_>(ctx, sd, frames_base, nframes, dev_sd, dev_params, dev_jobs, njobs, hist, hist_ints, geom, P0.pel);     \
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
