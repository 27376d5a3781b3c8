// The production search kernel for the default mv.Analyse configurations (dct=0 or dct=5 / SATD on 8-bit clips, no
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

#include <atomic>

#ifndef MVB_TMA_MIN_BLOCKS
#define MVB_TMA_MIN_BLOCKS 12 // chains (warps) per SM: caps registers at 168 (16 chains at 128 registers measured slower: instruction-cache and LSU pressure)
#endif
#ifndef MVB_TMA_MIN_BLOCKS_SMALL
#define MVB_TMA_MIN_BLOCKS_SMALL 16 // 8-bit 8x8 blocks
#endif
#ifndef MVB_TMA_MIN_BLOCKS_BIG
#define MVB_TMA_MIN_BLOCKS_BIG 12 // 16-bit 16x16 blocks: shared memory allows 9 chains per SM anyway; 168 registers
#endif
#ifndef MVB_CLD
#define MVB_CLD 0 // (experiment, measured slower) chroma windows of the 8-bit 16x16 configuration by LDGSTS into compact 16-byte rows (see Cfg::CLD)
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
__device__ __forceinline__ void cp_async4(uint32_t dst, const void *src) { // LDGSTS: global -> shared without a register
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
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

// SATD_ (dct = 5, pobLumaSAD<5>: the luma cost is Satd_C, SADFunctions.cpp:581-710): the four lanes of a candidate no longer take
// rows q, q + 4, ... of the block but whole 4x4 Hadamard tiles -- lane q owns tiles q * T .. q * T + T - 1 of the block's TCxTR tile
// grid (16x16: tile row q, all four tile columns; 8x8: the one tile (q & 1, q >> 1)).  Chroma stays a SAD.
template <int BPS, int BW, int BH, int LXR, int LYR, bool CHROMA, bool SATD_ = false>
struct Cfg {
    static constexpr bool SATD = SATD_;
    static_assert(!SATD_ || BPS == 1, "the SATD lane layout is written for 8-bit samples");
    static constexpr int TC = BW / 4, TR = BH / 4, T = (TC * TR) / 4; // tile grid, tiles per lane
    static constexpr int SRW = T < TC ? T : TC;                        // words (tile columns) a lane reads per row
    static_assert(!SATD_ || (TC * TR) % 4 == 0, "tiles per lane");
    static_assert(!SATD_ || T <= TC, "a lane's tiles lie in one tile row");
    static constexpr int CW = BW >> LXR, CH = BH >> LYR;
    static constexpr int R = 4; // guaranteed luma window radius in full pixels
    static constexpr int RCX = R >> LXR, RCY = R >> LYR;
    // window row bytes: block + 2R, plus up to 15 bytes because a TMA tile must START on a 16-byte
    // boundary of the innermost dimension (measured: an unaligned start coordinate raises
    // "illegal instruction" on sm_100), rounded up to the 16-byte box granularity
    static constexpr int LWB = ((BW + 2 * R) * BPS + 15 + 15) / 16 * 16;
    static constexpr int LWH = BH + 2 * R;
    // CLD: the chroma windows do not come by TMA but by 4-byte LDGSTS of all lanes, into rows that carry 3 bytes of
    // alignment slack instead of 15.  For 8-bit 16x16 blocks that halves the chroma part of a window buffer and brings a
    // chain's shared memory from 17.6 KB to 13.1 KB -- 16 chains per SM instead of 12 (228 KB, 1 KB reserved per CTA).
    static constexpr bool CLD = MVB_CLD && CHROMA && BPS == 1 && BW == 16 && BH == 16;
    static constexpr int CWB = CLD ? ((CW + 2 * RCX) * BPS + 3 + 3) / 4 * 4 : ((CW + 2 * RCX) * BPS + 15 + 15) / 16 * 16;
    static constexpr int CWH = CH + 2 * RCY;
    static_assert(!CLD || (CWB == 16 && CWH == 12), "the LDGSTS lane mapping below is written for 16-byte x 12-row chroma windows");
    // chains per SM the kernel is built for (the register cap follows).  8-bit 8x8 blocks: 16 chains at <= 128 registers,
    // no spills, 9 KB of shared memory -- measured 12.5 % faster than 12 chains (1080p, blk 8, ov 4: 2198 -> 1922 ms for
    // 7104 chains).  8-bit 16x16: 12 (at 16 chains the kernel was slower).
    static constexpr int MINB = (BPS == 2 && BW * BH >= 256) ? MVB_TMA_MIN_BLOCKS_BIG : ((BPS == 1 && BW * BH <= 64) ? MVB_TMA_MIN_BLOCKS_SMALL : MVB_TMA_MIN_BLOCKS);
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
    static constexpr bool PACK = !SATD_ && BPS == 1 && BW * BH <= 256 && 2 * CW * CH <= 256; // (a SATD does not fit 16 bits)
    static constexpr bool SRC_SMEM = BPS == 2 && BW * BH >= 256; // too many registers: read source rows from shared memory
    static constexpr int SRCB = r128(LBYTES + 2 * CBYTES);                   // one source-block buffer
    static constexpr int SRCOFF = 2 * BUF;                                   // [src 0][src 1]
    static constexpr int VECOFF = SRCOFF + 2 * SRCB;                          // hier[2], up, p3: 4 x 16 bytes
    static constexpr int COLDOFF = VECOFF + 64, BAROFF = COLDOFF + 256;       // level constants, 2 mbarriers
#ifndef MVB_ZB
#define MVB_ZB 1
#endif
    // ZB: the reference block at the zero vector -- the one start candidate that is usually far from a window placed
    // around the hierarchical predictor of fast motion -- is prefetched beside the source block (same position, same
    // word layout, LDGSTS), so that it alone no longer sends the whole first step to global memory
    static constexpr bool ZB = MVB_ZB != 0 && BPS == 1; // (16-bit 16x16 blocks would drop from 9 to 8 chains per SM for its 1.5 KB)
    static constexpr int GEOOFF = BAROFF + 16;                                // the next block's Geom (8 ints), computed once at prefetch time
    static constexpr int ZBOFF = GEOOFF + 32;                                 // [zero block 0][zero block 1]
    static constexpr int RAWOFF = ZBOFF + (ZB ? 2 * SRCB : 0);                // raw lo/hi words of an unaligned source block (optional)
    static constexpr int SMEM = r128(RAWOFF + 16), SMEM_RAW = r128(RAWOFF + 8 * SRCW + 16); // (+16: a row read fetches one word past the last block row)
    static_assert(BH % 4 == 0 && (BW * BPS) % 4 == 0, "block shape");
    static_assert(!CHROMA || (CH % 4 == 0 && (CW * BPS) % 4 == 0), "chroma block shape");
};

struct TmaMaps {
    CUtensorMap plane[3];  // 3-D {x, row of the whole plane, frame}: levels > 0 and pel 1
    CUtensorMap plane4[3]; // 4-D {x, row, sub-plane, frame} over level 0: pel 2
};

struct Geom {
    int x0, y0, xc, yc, dxMin, dxMax, dyMin, dyMax;
};

// Level constants and other state that is touched once per block (or less): kept in this warp's
// shared memory instead of registers -- the block loop needs its registers for the search itself.
struct Cold {
    const uint8_t *sbase[3], *rbase[3]; // level bases (padded origin, sub-plane 0) of the source / reference frame
    VEC *vectors;
    long long badSAD, LSAD;
    int pitch0, pitch1, sub0, sub1;
    int hp0, vp0, hpc, vpc, pw0, ph0, hps, vps, row0_0, row0_1;
    int nBlkX, nBlkY, nBlk, stepX, stepY, meander, nLambdaLevel, fs, ref_frame, use4d, smallest, srcAligned;
    int penaltyZero, pglobal, badrange, searchType, searchParam;
    int gmx, gmy;  // globalMV (sticky clipping, :859)
    int wo[2][4];  // window origins of the two window buffers
    int cxmax, cymin, cymax; // CLD: limits of a chroma window origin that keep every LDGSTS address inside the plane
    int zb;                  // this level prefetches the zero-vector reference block (aligned source blocks, no field shift)
};
static_assert(sizeof(Cold) <= 256, "Cold");

template <int BPS, int BW, int BH, int LXR, int LYR, bool CHROMA, bool SATD_ = false>
struct Chain {
    using C = Cfg<BPS, BW, BH, LXR, LYR, CHROMA, SATD_>;
    static constexpr long long INF = 0x7fffffffffffffffLL;
    static constexpr unsigned FULL = 0xffffffffu;

    // ---- source block rows of this lane (rows q, q+4, ...), unless C::SRC_SMEM
    // (SATD: rows 4 * trow + i, i = 0 .. 3, of the lane's tile row, words tcol .. tcol + SRW - 1)
    unsigned srcL[C::SRC_SMEM ? 1 : (C::SATD ? 4 : C::LR)][C::SATD ? C::SRW : C::RW], srcC[2][C::SRC_SMEM ? 1 : C::CR][C::CRW];
    uint32_t srcbuf; // shared-space address of the current source block
    Cold *K;
    int logPel, penaltyNew;
    // ---- block state
    Geom g;
    int wx0, wy0, wxc0, wyc0; // window origins (full-pel, padded-plane coordinates) of the CURRENT buffer
    long long nLambda, nMinCost;
    VEC best, predictor;
    int badcount, j, q;
    uint32_t sbuf; // shared-space address of the current window buffer
    uint32_t zbuf; // shared-space address of the current block's zero-vector reference block, or 0

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

    // Real_Satd_4x4_C (SADFunctions.cpp:605-641) on one tile held as four words of source and four of reference samples; the same
    // 32-bit pseudo-SIMD arithmetic (two 16-bit sums per register, abs2), so the result is the reference's bit for bit
    __device__ __forceinline__ static unsigned abs2(unsigned a) {
        const unsigned s = ((a >> 15) & 0x10001u) * 0xffffu;
        return (a + s) ^ s;
    }
    __device__ __forceinline__ static unsigned satd_tile(const unsigned s0, const unsigned s1, const unsigned s2, const unsigned s3, const unsigned r0, const unsigned r1, const unsigned r2,
                                                         const unsigned r3) {
        const unsigned sv[4] = {s0, s1, s2, s3}, rv[4] = {r0, r1, r2, r3};
        unsigned t0[4], t1[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const unsigned a0 = (sv[i] & 0xffu) - (rv[i] & 0xffu), a1 = ((sv[i] >> 8) & 0xffu) - ((rv[i] >> 8) & 0xffu);
            const unsigned a2 = ((sv[i] >> 16) & 0xffu) - ((rv[i] >> 16) & 0xffu), a3 = (sv[i] >> 24) - (rv[i] >> 24);
            const unsigned b0 = (a0 + a1) + ((a0 - a1) << 16), b1 = (a2 + a3) + ((a2 - a3) << 16);
            t0[i] = b0 + b1;
            t1[i] = b0 - b1;
        }
        unsigned sum = 0;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const unsigned *t = h ? t1 : t0;
            const unsigned u0 = t[0] + t[1], u1 = t[0] - t[1], u2 = t[2] + t[3], u3 = t[2] - t[3];
            const unsigned a = abs2(u0 + u2) + abs2(u0 - u2) + abs2(u1 + u3) + abs2(u1 - u3);
            sum += (a & 0xffffu) + (a >> 16);
        }
        return sum >> 1;
    }

    // window placement for a block whose (clipped) centre vector is (cx, cy)
    __device__ __forceinline__ void window_origin(const Geom &gg, int cx, int cy, int &ox, int &oy, int &oxc, int &oyc) const {
        ox = (((gg.x0 << logPel) + cx) >> logPel) - C::R;
        oy = (((gg.y0 << logPel) + cy) >> logPel) - C::R;
        oxc = (((gg.xc << logPel) + cdiv(cx, LXR)) >> logPel) - C::RCX;
        oyc = (((gg.yc << logPel) + cdiv(cy, LYR)) >> logPel) - C::RCY;
        // 16-byte aligned start of the innermost (x) coordinate, rounding toward -infinity
        ox = ((ox * BPS) & ~15) >> (BPS == 2 ? 1 : 0);
        if (C::CLD) {
            // LDGSTS windows start on a 4-byte boundary and must lie inside the plane (TMA zero-fills what is outside,
            // a plain load may not go there): a clamped window still serves every candidate it covers -- the rest
            // takes the global-memory path like any other candidate outside its window
            oxc = min(max(((oxc * BPS) & ~3) >> (BPS == 2 ? 1 : 0), 0), K->cxmax);
            oyc = min(max(oyc, K->cymin), K->cymax);
        } else {
            oxc = ((oxc * BPS) & ~15) >> (BPS == 2 ? 1 : 0);
        }
    }

    // Per-lane description of "my" candidate in a step
    struct Cand {
        int lvx, lvy, cvx, cvy; // luma vector, vector the chroma position derives from
        bool valid;             // in bounds (or exempt from the bounds test)
    };

    // Evaluates the candidates of a step: the four lanes 4j..4j+3 hold candidate j (identical Cand).
    // Returns, in every lane, its own candidate's (sadY, sadUV) -- lanes without a valid candidate get garbage.
    // ZSLOT: slot 0 is the zero-vector candidate (the first step of a block) and may be served by the prefetched block.
    template <int NS, bool ZSLOT = false>
    __device__ __forceinline__ void eval_step(const Cand &c, unsigned &mySy, unsigned &mySuv) const {
        const bool mine = j < NS && c.valid;
        const bool zlane = ZSLOT && C::ZB && j == 0 && zbuf != 0;
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
        const unsigned outside = __ballot_sync(FULL, mine && !inside && !zlane);
        unsigned aL = 0, aC = 0;
        if (outside == 0) {
            unsigned accL[C::LR], accC[2][C::CR]; // one accumulator per row: independent VABSDIFF4 chains
            if (!mine) { // harmless in-window dummy
                cbl = 0;
                cbc = 0;
            }
            uint32_t al = sbuf + (cbl & ~3) + q * C::LWB;
            unsigned shl = (unsigned)cbl << 3; // funnel shifts use the amount mod 32
            int lrow = 4 * C::LWB;             // bytes between two rows of this lane
            if (zlane) {                       // the prefetched block: rows of BW samples, no misalignment
                al = zbuf + q * (C::RW * 4);
                shl = 0;
                lrow = 4 * (C::RW * 4);
            }
            if constexpr (C::SATD) {
                // this lane's tiles: rows 4 * trow .. + 3, words tcol .. tcol + SRW - 1 of the candidate block
                const int trow = (q * C::T) / C::TC, tcol = (q * C::T) % C::TC;
                uint32_t at = sbuf + (cbl & ~3) + 4 * trow * C::LWB + 4 * tcol;
                int trs = C::LWB;
                if (zlane) {
                    at = zbuf + 4 * trow * (C::RW * 4) + 4 * tcol;
                    trs = C::RW * 4;
                }
                unsigned ref[4][C::SRW];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    unsigned r[C::SRW + 1];
                    lds_row(at + i * trs, r);
#pragma unroll
                    for (int w = 0; w < C::SRW; w++)
                        ref[i][w] = __funnelshift_r(r[w], r[w + 1], shl);
                }
#pragma unroll
                for (int w = 0; w < C::SRW; w++)
                    aL += satd_tile(srcL[0][w], srcL[1][w], srcL[2][w], srcL[3][w], ref[0][w], ref[1][w], ref[2][w], ref[3][w]);
            }
#pragma unroll
            for (int i = 0; i < (C::SATD ? 0 : C::LR); i++) {
                unsigned r[C::RW + 1], sw[C::RW];
                if (ZSLOT && C::ZB)
                    lds_row(al + i * lrow, r);
                else
                    lds_row(al + 4 * i * C::LWB, r);
                if (C::SRC_SMEM)
                    lds_vec(srcbuf + (q + 4 * i) * (C::RW * 4), sw);
                accL[i] = 0;
#pragma unroll
                for (int w = 0; w < C::RW; w++)
                    accL[i] += absd(__funnelshift_r(r[w], r[w + 1], shl), C::SRC_SMEM ? sw[w] : srcL[i][w]);
            }
#pragma unroll
            for (int i = 0; i < (C::SATD ? 0 : C::LR); i++)
                aL += accL[i];
            if (CHROMA) {
                uint32_t ac = sbuf + C::LREG + (cbc & ~3) + q * C::CWB;
                unsigned shc = (unsigned)cbc << 3;
                int crow = 4 * C::CWB, cplane = C::CREG;
                if (zlane) {
                    ac = zbuf + C::LBYTES + q * (C::CRW * 4);
                    shc = 0;
                    crow = 4 * (C::CRW * 4);
                    cplane = C::CBYTES;
                }
#pragma unroll
                for (int p = 0; p < 2; p++) {
#pragma unroll
                    for (int i = 0; i < C::CR; i++) {
                        unsigned r[C::CRW + 1], sw[C::CRW];
                        if (ZSLOT && C::ZB)
                            lds_row(ac + p * cplane + i * crow, r);
                        else
                            lds_row(ac + p * C::CREG + 4 * i * C::CWB, r);
                        if (C::SRC_SMEM)
                            lds_vec(srcbuf + C::LBYTES + p * C::CBYTES + (q + 4 * i) * (C::CRW * 4), sw);
                        accC[p][i] = 0;
#pragma unroll
                        for (int w = 0; w < C::CRW; w++)
                            accC[p][i] += absd(__funnelshift_r(r[w], r[w + 1], shc), C::SRC_SMEM ? sw[w] : srcC[p][i][w]);
                    }
                }
#pragma unroll
                for (int p = 0; p < 2; p++)
#pragma unroll
                    for (int i = 0; i < C::CR; i++)
                        aC += accC[p][i];
            }
        } else {
            // at least one candidate leaves the staged window: the whole step reads global memory.  Like the window
            // path: RW + 1 aligned words per row (the row pitch carries 4 bytes of slack for the last one), funnel
            // shifts for the byte misalignment, one accumulator per row.
            const int pitch0 = K->pitch0, pitch1 = K->pitch1;
            const unsigned long long la = (unsigned long long)(K->rbase[0] + (long long)subl * K->sub0 + (long long)(fy + q) * pitch0 + fx * BPS);
            const uint8_t *lw = (const uint8_t *)(la & ~3ull);
            const unsigned shl = ((unsigned)la & 3u) * 8u;
            unsigned accL[C::LR], accC[2][C::CR];
            if constexpr (C::SATD) {
                const int trow = (q * C::T) / C::TC, tcol = (q * C::T) % C::TC;
                const unsigned long long ta = (unsigned long long)(K->rbase[0] + (long long)subl * K->sub0 + (long long)(fy + 4 * trow) * pitch0 + fx * BPS + 4 * tcol);
                const uint8_t *tw = (const uint8_t *)(ta & ~3ull);
                const unsigned sht = ((unsigned)ta & 3u) * 8u;
                unsigned ref[4][C::SRW];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const unsigned *rw = (const unsigned *)(tw + (long long)i * pitch0);
                    unsigned r[C::SRW + 1];
#pragma unroll
                    for (int w = 0; w <= C::SRW; w++)
                        r[w] = __ldg(rw + w);
#pragma unroll
                    for (int w = 0; w < C::SRW; w++)
                        ref[i][w] = __funnelshift_r(r[w], r[w + 1], sht);
                }
#pragma unroll
                for (int w = 0; w < C::SRW; w++)
                    aL += satd_tile(srcL[0][w], srcL[1][w], srcL[2][w], srcL[3][w], ref[0][w], ref[1][w], ref[2][w], ref[3][w]);
            }
#pragma unroll
            for (int i = 0; i < (C::SATD ? 0 : C::LR); i++) {
                const unsigned *rw = (const unsigned *)(lw + (long long)(4 * i) * pitch0);
                unsigned r[C::RW + 1], sw[C::RW];
#pragma unroll
                for (int w = 0; w <= C::RW; w++)
                    r[w] = __ldg(rw + w);
                if (C::SRC_SMEM)
                    lds_vec(srcbuf + (q + 4 * i) * (C::RW * 4), sw);
                accL[i] = 0;
#pragma unroll
                for (int w = 0; w < C::RW; w++)
                    accL[i] += absd(__funnelshift_r(r[w], r[w + 1], shl), C::SRC_SMEM ? sw[w] : srcL[i][w]);
            }
#pragma unroll
            for (int i = 0; i < (C::SATD ? 0 : C::LR); i++)
                aL += accL[i];
            if (CHROMA) {
                const long long oc = (long long)subc * K->sub1 + (long long)(fyc + q) * pitch1 + fxc * BPS;
                const unsigned shc = ((unsigned)oc & 3u) * 8u; // the plane bases are 4-byte aligned
#pragma unroll
                for (int p = 0; p < 2; p++) {
                    const uint8_t *cw = K->rbase[1 + p] + (oc & ~3ll);
#pragma unroll
                    for (int i = 0; i < C::CR; i++) {
                        const unsigned *rw = (const unsigned *)(cw + (long long)(4 * i) * pitch1);
                        unsigned r[C::CRW + 1], sw[C::CRW];
#pragma unroll
                        for (int w = 0; w <= C::CRW; w++)
                            r[w] = __ldg(rw + w);
                        if (C::SRC_SMEM)
                            lds_vec(srcbuf + C::LBYTES + p * C::CBYTES + (q + 4 * i) * (C::CRW * 4), sw);
                        accC[p][i] = 0;
#pragma unroll
                        for (int w = 0; w < C::CRW; w++)
                            accC[p][i] += absd(__funnelshift_r(r[w], r[w + 1], shc), C::SRC_SMEM ? sw[w] : srcC[p][i][w]);
                    }
                }
#pragma unroll
                for (int p = 0; p < 2; p++)
#pragma unroll
                    for (int i = 0; i < C::CR; i++)
                        aC += accC[p][i];
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
        const int searchType = K->searchType, searchParam = K->searchParam;
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
    __device__ __forceinline__ void epz(const int (&prx)[4], const int (&pry)[4], int blkIdx) {
        // globalMV is clipped to every block's bounds in turn and stays clipped (:859)
        const int gmx = min(max(K->gmx, g.dxMin), g.dxMax - 1), gmy = min(max(K->gmy, g.dyMin), g.dyMax - 1);
        __syncwarp();
        if (j == 0 && q == 0) {
            K->gmx = gmx;
            K->gmy = gmy;
        }
        // slot j owns start candidate j: 0 zero, 1 global, 2 hierarchical, 3..6 predictors[0..3]
        Cand c;
        int px = 0, py = 0;
        px = j == 1 ? gmx : px;
        py = j == 1 ? gmy : py;
        px = j == 2 ? predictor.x : px;
        py = j == 2 ? predictor.y : py;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            px = j == 3 + i ? prx[i] : px;
            py = j == 3 + i ? pry[i] : py;
        }
        c.lvx = c.cvx = px;
        c.cvy = py;
        c.lvy = j == 0 ? K->fs : py; // zero: luma at (0, fieldShift), chroma at (0, 0)  (:836-839)
        c.valid = j < 3 || (j < 7 && ok(px, py));
        unsigned sy, suv;
        eval_step<7, true>(c, sy, suv);
        // per-candidate cost rules (:846, :870, :894, CheckMV0)
        const long long sad = (long long)sy + suv;
        const int pen = j == 0 ? K->penaltyZero : (j == 1 ? K->pglobal : 0);
        const long long cost = sad + ((pen * sad) >> 8) + (j >= 3 ? mv_cost(px, py) : 0ll);
        nMinCost = INF; // zero is accepted unconditionally, then strict '<' in order == first minimum
        int dummy = 0;
        resolve<7>(c, cost, sy, suv, true, 0, dummy);
        refine();
        // bad-block wide search (:938-963)
        const long long foundSAD = best.sad;
        const long long badSAD = K->badSAD;
        if (blkIdx > 1 && foundSAD > (badSAD + badSAD * badcount / 16)) {
            badcount++;
            const int badrange = K->badrange;
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

template <int BPS, int BW, int BH, int LXR, int LYR, bool CHROMA, bool SATD_ = false>
__global__ void __launch_bounds__(32, (Cfg<BPS, BW, BH, LXR, LYR, CHROMA, SATD_>::MINB)) k_search_tma(const __grid_constant__ TmaMaps maps, const SuperDesc *__restrict__ sdp, const SearchParams *__restrict__ params,
                                                   const SearchJob *__restrict__ jobs, int njobs, int *__restrict__ hist_all, int hist_ints, int chroma_rows) {
    using C = Cfg<BPS, BW, BH, LXR, LYR, CHROMA, SATD_>;
    using CH = Chain<BPS, BW, BH, LXR, LYR, CHROMA, SATD_>;
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

    // shared memory: [window 0][window 1][source 0][source 1][raw source words][vector slots][Cold][2 mbarriers]
    const uint32_t smem0 = (smem_u32(smem_raw) + 127u) & ~127u;
    uint8_t *const smem_gen = smem_raw + (smem0 - smem_u32(smem_raw));
    Cold &K = *(Cold *)(smem_gen + C::COLDOFF);
    const uint32_t bar0 = smem0 + C::BAROFF;
    const uint32_t vslot = smem0 + C::VECOFF; // hier(even t), hier(odd t), up, p3
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
    S.j = lane >> 2;
    S.q = lane & 3;
    S.K = &K;
    S.penaltyNew = P.pnew;

    // cooperative source-block load: word w of the block (luma rows, then U rows, then V rows) -> lane w % 32;
    // bit k of sw_luma: this lane's k-th word is a luma word
    unsigned sw_luma = 0;
#pragma unroll
    for (int k = 0; k < C::NSI; k++) {
        int w = lane + 32 * k;
        w = w < C::SRCW ? w : 0;
        sw_luma |= (w < C::LBYTES / 4 ? 1u : 0u) << k;
    }
    const long long refdelta = job.ref - job.src; // frames share one layout: a reference word lies this far from the source word

    for (int l = L - 1; l >= 0; l--) {
        S.logPel = l == 0 ? P.logPel : 0;
        const int nBlkX = P.nBlkX[l], nBlkY = P.nBlkY[l];
        int fs;
        {
            const mvb200_level_geom &g0 = sd.lv[0][l];
            const mvb200_level_geom &g1 = sd.lv[CHROMA ? 1 : 0][l];
            int st, stp;
            if (l == L - 1) {
                st = (L == 1) ? P.search : P.search_coarse;
                stp = (L == 1) ? P.nPelSearch : P.nSearchParam;
                fs = (L - 1 == 0) ? job.field_shift : 0;
            } else {
                st = (l == 0) ? P.search : P.search_coarse;
                stp = (l == 0) ? P.nPelSearch : P.nSearchParam;
                fs = (l == 0) ? job.field_shift : 0;
            }
            const int nScale = 1 << l;
            int nLambdaLevel = P.nLambda / ((1 << S.logPel) * (1 << S.logPel));
            if (P.plevel == 1)
                nLambdaLevel = nLambdaLevel * nScale;
            else if (P.plevel == 2)
                nLambdaLevel = nLambdaLevel * nScale * nScale;
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (int p = 0; p < 3; p++) {
                    const int qp = (CHROMA || p == 0) ? p : 0;
                    const long long off = sd.offset[qp] + (long long)sd.lv[qp][l].row0 * sd.pitch[qp];
                    K.sbase[p] = job.src + off;
                    K.rbase[p] = job.ref + off;
                }
                K.vectors = (VEC *)(job.out + P.level_off[l] + 4);
                K.badSAD = P.badSAD;
                K.LSAD = P.lsad;
                K.pitch0 = (int)sd.pitch[0];
                K.pitch1 = (int)sd.pitch[CHROMA ? 1 : 0];
                K.sub0 = K.pitch0 * g0.padded_h;
                K.sub1 = K.pitch1 * g1.padded_h;
                K.hp0 = g0.hpad;
                K.vp0 = g0.vpad;
                K.hpc = g1.hpad;
                K.vpc = g1.vpad;
                K.pw0 = g0.padded_w;
                K.ph0 = g0.padded_h;
                K.hps = g0.hpad >> l;
                K.vps = g0.vpad >> l;
                K.row0_0 = g0.row0;
                K.row0_1 = g1.row0;
                K.nBlkX = nBlkX;
                K.nBlkY = nBlkY;
                K.nBlk = nBlkX * nBlkY;
                K.stepX = BW - P.ovX;
                K.stepY = BH - P.ovY;
                K.meander = P.meander;
                K.nLambdaLevel = nLambdaLevel;
                K.fs = fs;
                K.ref_frame = job.ref_index;
                K.use4d = S.logPel == 1; // level 0 of a pel-2 clip: the 4-D maps (their row origin is level 0's row0)
                K.smallest = (l == L - 1);
                if (C::CLD) {
                    const int nsub = S.logPel == 1 ? 4 : 1;
                    K.cxmax = (int)sd.pitch[1] / BPS - C::CWB / BPS;
                    K.cymin = -g1.row0;
                    K.cymax = chroma_rows - g1.row0 - (nsub - 1) * g1.padded_h - C::CWH;
                }
                K.srcAligned = (((g0.hpad * BPS) | ((BW - P.ovX) * BPS) | (g1.hpad * BPS) | (((BW - P.ovX) >> LXR) * BPS)) & 3) == 0;
                K.zb = C::ZB && K.srcAligned && fs == 0;
                K.penaltyZero = P.pzero;
                K.pglobal = P.global ? P.pglobal : P.pzero;
                K.badrange = P.badrange;
                K.searchType = st;
                K.searchParam = stp;
            }
            __syncwarp();
        }
        if (l != L - 1) {
            const VEC *coarse = (const VEC *)(job.out + P.level_off[l + 1] + 4);
            if (P.global)
                estimate_global_mv(coarse, P.nBlkX[l + 1] * P.nBlkY[l + 1], 8192 * 1 * 2, hist, globalMVec, lane);
            interpolate_prediction(K.vectors, nBlkX, nBlkY, S.logPel, coarse, P.nBlkX[l + 1], P.nBlkY[l + 1], BW, BH, P.ovX, P.ovY, lane);
            __syncwarp();
        }
        // ---- doPobSearchMVs, PlaneOfBlocks.cpp:971-1131
        if (lane == 0) {
            K.gmx = (1 << S.logPel) * globalMVec.x;
            K.gmy = (1 << S.logPel) * globalMVec.y + fs;
        }
        __syncwarp();
        S.badcount = 0;
        const int nBlk = nBlkX * nBlkY;
        const bool smallest = (l == L - 1);

        // this lane's words of a source block at the level's origin: a block adds only its own offset (issue_src)
        const uint8_t *srcp[C::NSI];
#pragma unroll
        for (int k = 0; k < C::NSI; k++) {
            int w = lane + 32 * k;
            w = w < C::SRCW ? w : 0;
            constexpr int lw = C::LBYTES / 4, cw = C::CBYTES / 4;
            if (w < lw) {
                srcp[k] = K.sbase[0] + (long long)(w / C::RW) * K.pitch0 + (w % C::RW) * 4;
            } else {
                const int w2 = w - lw, pl = (CHROMA && w2 >= cw) ? 2 : 1, w3 = pl == 2 ? w2 - cw : w2;
                srcp[k] = K.sbase[pl] + (long long)(w3 / C::CRW) * K.pitch1 + (w3 % C::CRW) * 4;
            }
        }
        auto block_geom = [&](int bx, int by, Geom &gg) {
            const int hp0 = K.hp0, vp0 = K.vp0;
            gg.x0 = hp0 + K.stepX * bx;
            gg.y0 = vp0 + K.stepY * by;
            gg.xc = K.hpc + (K.stepX >> LXR) * bx;
            gg.yc = K.vpc + (K.stepY >> LYR) * by;
            gg.dxMax = (K.pw0 - gg.x0 - BW - hp0 + K.hps) << S.logPel;
            gg.dyMax = (K.ph0 - gg.y0 - BH - vp0 + K.vps) << S.logPel;
            gg.dxMin = -((gg.x0 - hp0 + K.hps) << S.logPel);
            gg.dyMin = -((gg.y0 - vp0 + K.vps) << S.logPel);
        };
        // a block's geometry is computed once, when the block is prefetched, and handed to the next iteration through
        // shared memory (8 ints): one lane stores, every lane reloads after the end-of-block __syncwarp
        auto store_geom = [&](const Geom &gg) {
            __syncwarp(); // every lane has read the previous contents
            if (lane == 0) {
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(smem0 + C::GEOOFF), "r"(gg.x0), "r"(gg.y0), "r"(gg.xc), "r"(gg.yc) : "memory");
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(smem0 + C::GEOOFF + 16), "r"(gg.dxMin), "r"(gg.dxMax), "r"(gg.dyMin), "r"(gg.dyMax) : "memory");
            }
        };
        auto load_geom = [&](Geom &gg) {
            unsigned a[4], b[4];
            lds_vec(smem0 + C::GEOOFF, a);
            lds_vec(smem0 + C::GEOOFF + 16, b);
            gg.x0 = (int)a[0];
            gg.y0 = (int)a[1];
            gg.xc = (int)a[2];
            gg.yc = (int)a[3];
            gg.dxMin = (int)b[0];
            gg.dxMax = (int)b[1];
            gg.dyMin = (int)b[2];
            gg.dyMax = (int)b[3];
        };
        // meander scan order (:1037-1055): successor of (bx, by, dir)
        auto succ = [&](int bx, int by, int dir, int &nbx, int &nby, int &ndir) {
            nbx = bx + dir;
            nby = by;
            ndir = dir;
            if (nbx < 0 || nbx >= nBlkX) {
                nby = by + 1;
                ndir = (K.meander == 0 || (nby & 1) == 0) ? 1 : -1;
                nbx = ndir == 1 ? 0 : nBlkX - 1;
            }
        };
        // Asynchronous copy (LDGSTS) of the source block at (x0, y0)/(xc, yc) into source buffer `buf`:
        // straight into place when every block of the level is 4-byte aligned, else as raw lo/hi words.
        auto issue_src = [&](int x0, int y0, int xc, int yc, int buf) {
            const bool aligned = K.srcAligned;
            const bool zb = K.zb != 0;
            const int orgl = y0 * K.pitch0 + x0 * BPS, orgc = yc * K.pitch1 + xc * BPS; // (a plane is far below 2 GB)
#pragma unroll
            for (int k = 0; k < C::NSI; k++) {
                if (lane + 32 * k < C::SRCW) {
                    const uint8_t *p = srcp[k] + (((sw_luma >> k) & 1u) ? orgl : orgc);
                    if (aligned) {
                        cp_async4(smem0 + C::SRCOFF + buf * C::SRCB + 4 * (lane + 32 * k), p);
                        if (C::ZB && zb) // same position in the reference frame: the block at the zero vector (sub-plane 0)
                            cp_async4(smem0 + C::ZBOFF + buf * C::SRCB + 4 * (lane + 32 * k), p + refdelta);
                    } else {
                        const unsigned *w = (const unsigned *)((unsigned long long)p & ~3ull);
                        cp_async4(smem0 + C::RAWOFF + 8 * (lane + 32 * k), w);
                        cp_async4(smem0 + C::RAWOFF + 8 * (lane + 32 * k) + 4, w + 1);
                    }
                }
            }
        };
        // completes the copy of the source block (x0, xc give the byte misalignment) and loads this lane's rows
        auto commit_src = [&](int x0, int xc, int buf) {
            cp_async_wait_all();
            __syncwarp();
            const uint32_t dst = smem0 + C::SRCOFF + buf * C::SRCB;
            if (!K.srcAligned) {
                const unsigned shl = ((unsigned)(x0 * BPS) & 3u) * 8u, shc = ((unsigned)(xc * BPS) & 3u) * 8u;
#pragma unroll
                for (int k = 0; k < C::NSI; k++)
                    if (lane + 32 * k < C::SRCW) {
                        unsigned lh[2];
                        lds_vec(smem0 + C::RAWOFF + 8 * (lane + 32 * k), lh);
                        sts32(dst + 4 * (lane + 32 * k), __funnelshift_r(lh[0], lh[1], ((sw_luma >> k) & 1u) ? shl : shc));
                    }
                __syncwarp();
            }
            S.srcbuf = dst;
            if constexpr (C::SATD) { // the lane's tile row: rows 4 * trow .. + 3, words tcol .. tcol + SRW - 1
                const int trow = (S.q * C::T) / C::TC, tcol = (S.q * C::T) % C::TC;
#pragma unroll
                for (int i = 0; i < 4; i++)
                    lds_vec(dst + (4 * trow + i) * (C::RW * 4) + 4 * tcol, S.srcL[i]);
            }
            if (!C::SRC_SMEM) {
#pragma unroll
                for (int i = 0; i < (C::SATD ? 0 : C::LR); i++)
                    lds_vec(dst + (S.q + 4 * i) * (C::RW * 4), S.srcL[i]);
                if (CHROMA) {
#pragma unroll
                    for (int p = 0; p < 2; p++)
#pragma unroll
                        for (int i = 0; i < C::CR; i++)
                            lds_vec(dst + C::LBYTES + p * C::CBYTES + (S.q + 4 * i) * (C::CRW * 4), S.srcC[p][i]);
                }
            }
        };
        // The word this lane fetches for the vector slots: lanes 0-3 the hierarchical predictor of the block at
        // (hx, hy) (if hv), lanes 4-7 / 8-11 the "up" and "down-right / up-right" neighbours of the block at
        // (bx, by, dir) (if nv) -- pobFetchPredictors' reads (:419-463); zeroMVfieldShifted where there is none.
        auto fetch_word = [&](bool hv, int hx, int hy, bool nv, int bx, int by, int dir) -> int {
            const int grp = lane >> 2, wi = lane & 3;
            const int *vw = (const int *)K.vectors;
            const int zero = wi == 1 ? K.fs : 0;
            int idx = -1;
            if (grp == 0) {
                idx = hv ? hy * nBlkX + hx : -1;
            } else if (grp == 1) {
                idx = (nv && by > 0) ? (by - 1) * nBlkX + bx : -1;
            } else if (grp == 2 && nv) {
                const bool hasNext = (dir == 1 && bx < nBlkX - 1) || (dir == -1 && bx > 0);
                if (by < nBlkY - 1 && hasNext)
                    idx = (by + 1) * nBlkX + bx + dir;
                else if (by > 0 && hasNext)
                    idx = (by - 1) * nBlkX + bx + dir;
            }
            return idx >= 0 ? __ldcg(vw + 4 * idx + wi) : zero; // written by this warp earlier in the chain: L2-coherent loads
        };
        auto slot_of_lane = [&](int hslot) -> uint32_t { // where fetch_word's result goes
            const int grp = lane >> 2, wi = lane & 3;
            return vslot + 16 * (grp == 0 ? hslot : grp + 1) + 4 * wi;
        };
        auto read_vec = [&](int slot) -> VEC {
            unsigned v[4];
            lds_vec(vslot + 16 * slot, v);
            VEC r;
            r.x = (int)v[0];
            r.y = (int)v[1];
            r.sad = (long long)(((unsigned long long)v[3] << 32) | v[2]);
            return r;
        };
        // stages the window of the block with geometry gg around its clipped hierarchical predictor (hx, hy)
        auto stage = [&](const Geom &gg, int hx, int hy, int buf) {
            const int cx = min(max(hx, gg.dxMin), gg.dxMax - 1), cy = min(max(hy, gg.dyMin), gg.dyMax - 1);
            int wo[4];
            S.window_origin(gg, cx, cy, wo[0], wo[1], wo[2], wo[3]);
            fence_proxy_async(); // earlier generic-proxy reads of this buffer are ordered before the async writes
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < 4; k++)
                    K.wo[buf][k] = wo[k];
                const uint32_t dst = smem0 + buf * C::BUF, bar = bar0 + 8 * buf;
                const int ref_frame = K.ref_frame;
                constexpr int TMA_CT = (CHROMA && !C::CLD) ? 2 * C::CT : 0; // chroma bytes per sub-plane that arrive by TMA
                if (K.use4d) {
                    mbar_expect_tx(bar, (uint32_t)(4 * (C::LT + TMA_CT)));
                    tma_load_4d(dst, &maps.plane4[0], wo[0], wo[1], 0, ref_frame, bar);
                    if (CHROMA && !C::CLD) {
                        tma_load_4d(dst + C::LREG, &maps.plane4[1], wo[2], wo[3], 0, ref_frame, bar);
                        tma_load_4d(dst + C::LREG + C::CREG, &maps.plane4[2], wo[2], wo[3], 0, ref_frame, bar);
                    }
                } else {
                    mbar_expect_tx(bar, (uint32_t)(C::LT + TMA_CT));
                    tma_load_3d(dst, &maps.plane[0], wo[0], K.row0_0 + wo[1], ref_frame, bar);
                    if (CHROMA && !C::CLD) {
                        tma_load_3d(dst + C::LREG, &maps.plane[1], wo[2], K.row0_1 + wo[3], ref_frame, bar);
                        tma_load_3d(dst + C::LREG + C::CREG, &maps.plane[2], wo[2], K.row0_1 + wo[3], ref_frame, bar);
                    }
                }
            }
            if (C::CLD) {
                // chroma windows: 16-byte rows, 12 rows per sub-plane, word w of a plane's window (sub-plane, row, column
                // in that order) travels in lane w % 32: lane = (row within a group of 8) * 4 + column word.  The data is
                // complete at the next cp.async.wait_all -- commit_src() at the end of the block, or the callers' own.
                const int a = lane >> 2;
                const int pitch1 = K.pitch1, sub1 = K.sub1;
                const bool all = K.use4d; // level 0 of a pel-2 clip: four sub-planes, else one
                const int off0 = a * pitch1;                                           // words 0..31: sub-plane 0 (and 2), rows 0..7
                const int off1 = a < 4 ? off0 + 8 * pitch1 : off0 - 4 * pitch1 + sub1; // words 32..63: rows 8..11, then rows 0..3 of the next sub-plane
                const int off2 = off0 + 4 * pitch1 + sub1;                             // words 64..95: rows 4..11 of sub-plane 1 (and 3)
                const long long org = (long long)wo[3] * pitch1 + wo[2] * BPS + (lane & 3) * 4;
#pragma unroll
                for (int p = 0; p < 2; p++) {
                    const uint8_t *gp = K.rbase[1 + p] + org;
                    const uint32_t d = smem0 + buf * C::BUF + C::LREG + p * C::CREG + 4 * lane;
                    cp_async4(d, gp + off0);
                    if (all || a < 4)
                        cp_async4(d + 128, gp + off1);
                    if (all) {
                        cp_async4(d + 256, gp + off2);
                        gp += 2 * (long long)sub1;
                        cp_async4(d + 384, gp + off0);
                        cp_async4(d + 512, gp + off1);
                        cp_async4(d + 640, gp + off2);
                    }
                }
            }
            __syncwarp();
        };

        // ---- level prologue: vectors of block 0 (and the hierarchical predictor of block 1), source block 0, window 0
        int bx = 0, by = 0, dir = 1;
        {
            int n1x, n1y, n1d;
            succ(bx, by, dir, n1x, n1y, n1d);
            const int w0 = fetch_word(true, 0, 0, true, 0, 0, 1);
            const int w1 = fetch_word(nBlk > 1, n1x, n1y, false, 0, 0, 1);
            if (lane < 12)
                sts32(slot_of_lane(0), (unsigned)w0);
            if (lane < 4)
                sts32(slot_of_lane(1), (unsigned)w1);
            block_geom(0, 0, S.g);
            store_geom(S.g);
            issue_src(S.g.x0, S.g.y0, S.g.xc, S.g.yc, 0);
            commit_src(S.g.x0, S.g.xc, 0); // (its __syncwarp also publishes the vector slots)
            if (!smallest) {
                const VEC h0 = read_vec(0);
                stage(S.g, h0.x, h0.y, 0);
                if (C::CLD) {
                    cp_async_wait_all();
                    __syncwarp();
                }
            }
        }
        int cbuf = 0;
        for (int t = 0; t < nBlk; t++) {
            const int blkIdx = by * nBlkX + bx;
            const bool hasMore = t + 1 < nBlk;
            int nbx, nby, ndir;
            succ(bx, by, dir, nbx, nby, ndir);
            // this block's vectors (fetched during the previous block)
            const VEC hier = read_vec(t & 1), up = read_vec(2), p3 = read_vec(3);
            __syncwarp(); // slots are rewritten at the end of this block
            // fetch for later: hier(t+2), up/p3(t+1), source block t+1
            int pv;
            {
                int n2x, n2y, n2d;
                succ(nbx, nby, ndir, n2x, n2y, n2d);
                pv = fetch_word(t + 2 < nBlk, n2x, n2y, hasMore, nbx, nby, ndir);
            }
            load_geom(S.g); // = block_geom(bx, by): stored when this block was prefetched
            S.nLambda = by == 0 ? 0 : K.nLambdaLevel;
            S.predictor = S.clip(hier);
            // pobFetchPredictors, PlaneOfBlocks.cpp:419-463
            int prx[4], pry[4];
            {
                const int fs = K.fs;
                const bool hasPrev = (dir == 1 && bx > 0) || (dir == -1 && bx < nBlkX - 1);
                VEC p1; // predictors[1]: the previous block of the scan
                p1.x = hasPrev ? S.best.x : 0;
                p1.y = hasPrev ? S.best.y : fs;
                p1.sad = hasPrev ? S.best.sad : 0;
                p1 = S.clip(p1);
                const VEC p2 = S.clip(up), p3c = S.clip(p3);
                prx[1] = p1.x;
                pry[1] = p1.y;
                prx[2] = p2.x;
                pry[2] = p2.y;
                prx[3] = p3c.x;
                pry[3] = p3c.y;
                long long p0sad = p1.sad;
                if (by > 0) {
                    prx[0] = max(min(p1.x, p2.x), min(max(p1.x, p2.x), p3c.x));
                    pry[0] = max(min(p1.y, p2.y), min(max(p1.y, p2.y), p3c.y));
                    p0sad = max(p1.sad, max(p2.sad, p3c.sad));
                } else {
                    prx[0] = p1.x;
                    pry[0] = p1.y;
                }
                if (smallest) {
                    S.predictor.x = prx[0];
                    S.predictor.y = pry[0];
                    S.predictor.sad = p0sad;
                    // the window centre depends on the previous block here: stage now (few blocks on this level)
                    stage(S.g, S.predictor.x, S.predictor.y, cbuf);
                    if (C::CLD) {
                        cp_async_wait_all();
                        __syncwarp();
                    }
                }
                const long long LSAD = K.LSAD;
                double scale = __ddiv_rn((double)LSAD, (double)(LSAD + (S.predictor.sad >> 1)));
                S.nLambda = __double2ll_rz(__dmul_rn(__dmul_rn((double)S.nLambda, scale), scale));
            }
            // the window of this block was issued one block ago: wait for it
            mbar_wait(bar0 + 8 * cbuf, (phasebits >> cbuf) & 1u);
            phasebits ^= 1u << cbuf;
            S.sbuf = smem0 + cbuf * C::BUF;
            S.zbuf = (C::ZB && K.zb) ? smem0 + C::ZBOFF + (t & 1) * C::SRCB : 0u;
            S.wx0 = K.wo[cbuf][0];
            S.wy0 = K.wo[cbuf][1];
            S.wxc0 = K.wo[cbuf][2];
            S.wyc0 = K.wo[cbuf][3];
            // ... and the next block's window and source block are fetched while this block is searched
            if (hasMore) {
                Geom gn;
                block_geom(nbx, nby, gn);
                store_geom(gn);
                issue_src(gn.x0, gn.y0, gn.xc, gn.yc, (t + 1) & 1);
                if (!smallest) {
                    const VEC hn = read_vec((t + 1) & 1);
                    stage(gn, hn.x, hn.y, cbuf ^ 1);
                }
            }
            S.epz(prx, pry, blkIdx);
            if (lane == 0)
                stvec(&K.vectors[blkIdx], S.best);
            if (hasMore) {
                // publish the fetched vector words; at a meander turn the next block's "up" neighbour is this block
                if (nby > 0 && (nby * nBlkX + nbx) - nBlkX == blkIdx && (lane >> 2) == 1) {
                    const int wi = lane & 3;
                    pv = wi == 0 ? S.best.x : (wi == 1 ? S.best.y : (wi == 2 ? (int)(unsigned)(unsigned long long)S.best.sad : (int)(unsigned)((unsigned long long)S.best.sad >> 32)));
                }
                if (lane < 12)
                    sts32(slot_of_lane(t & 1), (unsigned)pv);
                commit_src(K.hp0 + K.stepX * nbx, K.hpc + (K.stepX >> LXR) * nbx, (t + 1) & 1); // (syncs: publishes the slots, orders the vector store)
                bx = nbx;
                by = nby;
                dir = ndir;
                if (!smallest)
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

template <int BPS, int BW, int BH, int LXR, int LYR, bool CHROMA, bool SATD_>
int launch_cfg(Ctx *ctx, const SuperDesc &sd, const uint8_t *frames_base, int nframes, const SuperDesc *dev_sd, const SearchParams *dev_params, const SearchJob *dev_jobs,
               int njobs, int *hist, int hist_ints, const mvb200_super_geom &geom, int pel, bool need_raw, bool *handled) {
    using C = Cfg<BPS, BW, BH, LXR, LYR, CHROMA, SATD_>;
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) {
        *handled = false;
        return 0;
    }
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
            if (r != CUDA_SUCCESS) { // (a layout the tensor-map encoder rejects: the register-blocked kernel takes the job)
                *handled = false;
                return 0;
            }
        }
        if (pel == 2) { // level 0 as {x, row, sub-plane, frame}: one tile load brings the window of all four sub-planes
            const mvb200_level_geom &g0 = sd.lv[p][0];
            cuuint64_t dims[4] = {(cuuint64_t)(sd.pitch[p] / BPS), (cuuint64_t)g0.padded_h, 4, (cuuint64_t)nframes};
            cuuint64_t strides[3] = {(cuuint64_t)sd.pitch[p], (cuuint64_t)(sd.pitch[p] * g0.padded_h), (cuuint64_t)sd.frame_bytes};
            cuuint32_t box[4] = {bw, bh, 4, 1};
            cuuint32_t estr[4] = {1, 1, 1, 1};
            CUresult r = enc(&maps.plane4[p], dt, 4, (void *)(frames_base + sd.offset[p] + (long long)g0.row0 * sd.pitch[p]), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) {
                *handled = false;
                return 0;
            }
        }
    }
    auto kern = k_search_tma<BPS, BW, BH, LXR, LYR, CHROMA, SATD_>;
    // function attributes are per device: one bit per device ordinal (a second context on another GPU of the same process)
    static std::atomic<unsigned long long> configured{0};
    const unsigned long long devbit = 1ull << (ctx->device & 63);
    if (!(configured.load(std::memory_order_acquire) & devbit)) {
        MVB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_RAW));
        MVB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        configured.fetch_or(devbit, std::memory_order_release);
    }
    kern<<<njobs, 32, need_raw ? C::SMEM_RAW : C::SMEM, ctx->stream>>>(maps, dev_sd, dev_params, dev_jobs, njobs, hist, hist_ints, geom.plane_h[CHROMA ? 1 : 0]);
    ctx->launches++;
    MVB_CUDA_CHECK(cudaGetLastError());
    return 0;
}


// ------------------------------------------------------------------ Recalculate with the chain's candidate layout
// doPobRecalculateMVs (PlaneOfBlocks.cpp:1158-1424): a block's predictor comes from the OLD vector field, there is no neighbour
// predictor and no bad-block path, so blocks are independent -- one warp per block, four per CTA.  What a block does is what a
// chain's block does after its predictors: one candidate (the predictor, cost = its SAD / SATD, no penalty), then, above
// thSAD, pobRefine around it -- so the warp runs Chain's own eval_step / refine (8 candidates x 4 lanes per step, REDUX arg-min,
// the SATD tile layout for dct = 5) with no staged window: every candidate reads the reference frame through the global path.
// The generic k_recalc (search.cu) evaluated one candidate at a time with one lane per 4x4 tile or row group; it stays for
// the configurations this kernel is not compiled for (and other search types / dct modes).
template <int BPS, int BW, int BH, int LXR, int LYR, bool CHROMA, bool SATD_>
__global__ void __launch_bounds__(128) k_recalc_tma(const SuperDesc *__restrict__ sdp, const RecalcParams *__restrict__ pp, const RecalcJob *__restrict__ jobs) {
    using C = Cfg<BPS, BW, BH, LXR, LYR, CHROMA, SATD_>;
    using CH = Chain<BPS, BW, BH, LXR, LYR, CHROMA, SATD_>;
    static_assert(!C::SRC_SMEM, "the source block lives in registers here");
    __shared__ Cold cold[4];
    const SuperDesc &sd = *sdp;
    const RecalcParams &P = *pp;
    const RecalcJob job = jobs[blockIdx.y];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = P.nBlkX * P.nBlkY;
    const int blkIdx = blockIdx.x * 4 + warp;
    if (blkIdx >= n)
        return;
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
    const int by = blkIdx / P.nBlkX, bx = blkIdx - by * P.nBlkX;
    if (!live) { // gopWriteDefaultToArray, GroupOfPlanes.c:150-164
        VEC d;
        d.x = 0;
        d.y = 0;
        d.sad = P.verybigSAD;
        if (lane == 0)
            stvec(&res[blkIdx], d);
        if (P.divide && lane < 4) {
            VEC *q = (VEC *)(job.out + 12 + n * 16 + 4);
            stvec(&q[(by * 2 + (lane >> 1)) * (P.nBlkX * 2) + bx * 2 + (lane & 1)], d);
        }
        return;
    }
    const mvb200_level_geom &g0 = sd.lv[0][0];
    const mvb200_level_geom &g1 = sd.lv[CHROMA ? 1 : 0][0];
    Cold &K = cold[warp];
    if (lane == 0) {
#pragma unroll
        for (int p = 0; p < 3; p++) {
            const int qp = (CHROMA || p == 0) ? p : 0;
            const long long off = sd.offset[qp] + (long long)sd.lv[qp][0].row0 * sd.pitch[qp];
            K.sbase[p] = job.src + off;
            K.rbase[p] = job.ref + off;
        }
        K.pitch0 = (int)sd.pitch[0];
        K.pitch1 = (int)sd.pitch[CHROMA ? 1 : 0];
        K.sub0 = K.pitch0 * g0.padded_h;
        K.sub1 = K.pitch1 * g1.padded_h;
        K.searchType = P.search;
        K.searchParam = P.nSearchParam;
    }
    __syncwarp();
    CH S;
    S.j = lane >> 2;
    S.q = lane & 3;
    S.K = &K;
    S.logPel = P.logPel;
    S.penaltyNew = P.pnew;
    S.badcount = 0;
    S.srcbuf = 0;
    S.sbuf = 0;
    S.zbuf = 0;
    S.wx0 = S.wy0 = S.wxc0 = S.wyc0 = 1 << 20; // no staged window: no candidate is "inside"
    const int stepX = BW - P.ovX, stepY = BH - P.ovY;
    S.g.x0 = g0.hpad + stepX * bx;
    S.g.y0 = g0.vpad + stepY * by;
    S.g.xc = g1.hpad + (stepX >> LXR) * bx;
    S.g.yc = g1.vpad + (stepY >> LYR) * by;
    S.g.dxMax = (g0.padded_w - S.g.x0 - BW) << P.logPel; // :1274-1277, the whole padded plane
    S.g.dyMax = (g0.padded_h - S.g.y0 - BH) << P.logPel;
    S.g.dxMin = -(S.g.x0 << P.logPel);
    S.g.dyMin = -(S.g.y0 << P.logPel);
    S.nLambda = by == 0 ? 0 : P.nLambda / ((1 << P.logPel) * (1 << P.logPel));
    // this lane's rows (or tiles) of the source block
    {
        const uint8_t *sl = K.sbase[0] + (long long)S.g.y0 * K.pitch0 + S.g.x0 * BPS;
        if constexpr (C::SATD) {
            const int trow = (S.q * C::T) / C::TC, tcol = (S.q * C::T) % C::TC;
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int w = 0; w < C::SRW; w++)
                    S.srcL[i][w] = CH::ldg_un(sl + (long long)(4 * trow + i) * K.pitch0 + 4 * (tcol + w));
        } else {
#pragma unroll
            for (int i = 0; i < C::LR; i++)
#pragma unroll
                for (int w = 0; w < C::RW; w++)
                    S.srcL[i][w] = CH::ldg_un(sl + (long long)(S.q + 4 * i) * K.pitch0 + 4 * w);
        }
        if (CHROMA) {
#pragma unroll
            for (int p = 0; p < 2; p++) {
                const uint8_t *sc = K.sbase[1 + p] + (long long)S.g.yc * K.pitch1 + S.g.xc * BPS;
#pragma unroll
                for (int i = 0; i < C::CR; i++)
#pragma unroll
                    for (int w = 0; w < C::CRW; w++)
                        S.srcC[p][i][w] = CH::ldg_un(sc + (long long)(S.q + 4 * i) * K.pitch1 + 4 * w);
            }
        }
    }
    // predictor from the old field (:1279-1330)
    const VEC *oldv = (const VEC *)(job.old + __ldg((const int *)job.old) - 16 * P.oBlkX * P.oBlkY);
    const int nStepXold = P.oBlkW - P.oOvX, nStepYold = P.oBlkH - P.oOvY;
    const int centerX = BW / 2 + stepX * bx;
    const int blkxold = (centerX - P.oBlkW / 2) / nStepXold;
    const int centerY = BH / 2 + stepY * by;
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
    S.predictor = S.clip(vo);
    S.predictor.sad = vo.sad * (BW * BH) / (P.oBlkW * P.oBlkH);
    S.best = S.predictor;
    // the predictor itself: one candidate in slot 0, cost = its SAD (:1340-1346)
    typename CH::Cand c;
    c.lvx = c.cvx = S.predictor.x;
    c.lvy = c.cvy = S.predictor.y;
    c.valid = S.j == 0;
    unsigned sy, suv;
    S.template eval_step<1>(c, sy, suv);
    const unsigned s0 = __shfl_sync(CH::FULL, sy, 0), s1 = __shfl_sync(CH::FULL, suv, 0);
    S.best.sad = (long long)s0 + s1;
    S.nMinCost = S.best.sad;
    if (S.best.sad > P.thSAD)
        S.refine();
    if (lane == 0)
        stvec(&res[blkIdx], S.best);
}

// The compiled configurations, in the order of TMA_CONFIGS below; -1 when a parameter set needs another kernel.
// (ID, bytes per sample, block, log2 chroma subsampling, chroma, SATD luma cost: dct = 5)
#define TMA_CONFIGS(X)                                                                                                                                                         \
    X(0, 1, 16, 16, 1, 1, true, false) X(1, 2, 16, 16, 1, 1, true, false) X(2, 1, 8, 8, 1, 1, true, false) X(3, 2, 8, 8, 1, 1, true, false) X(4, 1, 16, 16, 0, 0, false, false) \
    X(5, 1, 8, 8, 0, 0, false, false) X(6, 1, 16, 16, 1, 1, true, true) X(7, 1, 8, 8, 1, 1, true, true)

static int tma_match(const SuperDesc &sd, const SearchParams *host_params, int nparams, bool *need_raw) {
    const SearchParams &P0 = host_params[0];
    for (int i = 0; i < nparams; i++) {
        const SearchParams &P = host_params[i];
        auto okType = [](int t) { return t == 3 || t == 4 || t == 5; };
        if ((P.dctmode != 0 && P.dctmode != 5) || P.tryMany || !okType(P.search) || !okType(P.search_coarse) || P.pel > 2)
            return -1;
        if (P.blkW != P0.blkW || P.blkH != P0.blkH || P.chroma != P0.chroma || P.bps != P0.bps || P.pel != P0.pel || P.dctmode != P0.dctmode)
            return -1;
    }
    // the winner's vector travels as two 16-bit halves of one shuffle
    if ((sd.lv[0][0].padded_w << P0.logPel) >= 32768 || (sd.lv[0][0].padded_h << P0.logPel) >= 32768)
        return -1;
    const int bw = P0.blkW, bh = P0.blkH, bps = P0.bps, lx = sd.logxR, ly = sd.logyR;
    // source blocks that are not 4-byte aligned in some level need the raw staging area (same test as K.srcAligned)
    *need_raw = false;
    for (int i = 0; i < nparams; i++)
        for (int l = 0; l < host_params[i].levels; l++) {
            const int stepX = bw - host_params[i].ovX, c = P0.chroma ? 1 : 0;
            if ((((sd.lv[0][l].hpad * bps) | (stepX * bps) | (sd.lv[c][l].hpad * bps) | ((stepX >> (c ? lx : 0)) * bps)) & 3) != 0)
                *need_raw = true;
        }
    const bool chroma = P0.chroma, satd = P0.dctmode == 5;
#define MATCH(ID_, BPS_, BW_, BH_, LX_, LY_, CH_, SATD_)                                                                      \
    if (bps == BPS_ && bw == BW_ && bh == BH_ && chroma == CH_ && satd == SATD_ && (!CH_ || (lx == LX_ && ly == LY_))) \
        return ID_;
    TMA_CONFIGS(MATCH)
#undef MATCH
    return -1;
}

} // namespace

// Launches the TMA kernel when every parameter set matches one compiled configuration.
int launch_search_tma(Ctx *ctx, const SuperDesc &sd, const mvb200_super_geom &geom, const uint8_t *frames_base, int nframes, const SearchParams *host_params, int nparams,
                      const SuperDesc *dev_sd, const SearchParams *dev_params, const SearchJob *dev_jobs, int njobs, int *hist, int hist_ints, bool *handled) {
    *handled = false;
    if (!frames_base || nframes <= 0)
        return 0;
    bool need_raw = false;
    const int id = tma_match(sd, host_params, nparams, &need_raw);
    if (id < 0)
        return 0;
    if (!get_encode_fn()) // (the driver is older than the tensor-map API: the register-blocked kernel takes the job)
        return 0;
    *handled = true;
#define TRY(ID_, BPS_, BW_, BH_, LX_, LY_, CH_, SATD_) \
    if (id == ID_)                                     \
        return launch_cfg<BPS_, BW_, BH_, LX_, LY_, CH_, SATD_>(ctx, sd, frames_base, nframes, dev_sd, dev_params, dev_jobs, njobs, hist, hist_ints, geom, host_params[0].pel, need_raw, handled);
    TMA_CONFIGS(TRY)
#undef TRY
    return 0;
}

namespace {
template <int BPS, int BW, int BH, int LXR, int LYR, bool CHROMA, bool SATD_>
bool recalc_launch(Ctx *ctx, int n, int njobs, const SuperDesc *dev_sd, const RecalcParams *dev_params, const RecalcJob *dev_jobs) {
    if constexpr (Cfg<BPS, BW, BH, LXR, LYR, CHROMA, SATD_>::SRC_SMEM) {
        return false; // (16-bit 16x16: the chain reads its source rows from shared memory; the generic kernel takes these)
    } else {
        k_recalc_tma<BPS, BW, BH, LXR, LYR, CHROMA, SATD_><<<dim3((n + 3) / 4, njobs), 128, 0, ctx->stream>>>(dev_sd, dev_params, dev_jobs);
        return true;
    }
}
} // namespace

// Recalculate on the chain's candidate layout (k_recalc_tma) when the new grid is one of the compiled configurations; *handled
// stays false otherwise and launch_recalc runs the generic kernel.
int launch_recalc_tma(Ctx *ctx, const SuperDesc &sd, const RecalcParams &P, const SuperDesc *dev_sd, const RecalcParams *dev_params, const RecalcJob *dev_jobs, int njobs, bool *handled) {
    *handled = false;
    auto okType = [](int t) { return t == 3 || t == 4 || t == 5; };
    if ((P.dctmode != 0 && P.dctmode != 5) || !okType(P.search) || P.logPel > 1)
        return 0;
    if ((sd.lv[0][0].padded_w << P.logPel) >= 32768 || (sd.lv[0][0].padded_h << P.logPel) >= 32768)
        return 0;
    const int bps = P.bps, bw = P.blkW, bh = P.blkH, lx = sd.logxR, ly = sd.logyR;
    const bool chroma = P.chroma != 0, satd = P.dctmode == 5;
    const int n = P.nBlkX * P.nBlkY;
    // Parity first: the configurations whose Recalculate output has been compared with the reference ON THE GPU (goldens
    // c1_640x360_p8_blk8_pel1, blk16_ov8_pel2_p8_r3, divide2_blk16_esa_neg_badrange: 8-bit 8x8 4:2:0; noise_badblocks_blk8 and the
    // 4K digests of bench.py --config c4: the same with SATD) run here by default; the other compiled configurations (16x16, 16-bit
    // 8x8, gray) share every line of code but no GPU run has pinned them yet: MVB200_RECALC_TMA_ALL=1 enables them.
    const char *all = getenv("MVB200_RECALC_TMA_ALL");
    const bool every = all && all[0] == '1';
#define RECALC(ID_, BPS_, BW_, BH_, LX_, LY_, CH_, SATD_)                                                                                                       \
    if (!*handled && (every || ID_ == 2 || ID_ == 7) && bps == BPS_ && bw == BW_ && bh == BH_ && chroma == CH_ && satd == SATD_ && (!CH_ || (lx == LX_ && ly == LY_))) \
        *handled = recalc_launch<BPS_, BW_, BH_, LX_, LY_, CH_, SATD_>(ctx, n, njobs, dev_sd, dev_params, dev_jobs);
    TMA_CONFIGS(RECALC)
#undef RECALC
    if (*handled) {
        ctx->launches++;
        MVB_CUDA_CHECK(cudaGetLastError());
    }
    return 0;
}

// Pure host: chains (CTAs of one warp) per SM the TMA kernel runs for these parameter sets -- the smaller of the register
// cap it is compiled for (__launch_bounds__) and what 227 KB of shared memory per SM hold (1 KB is reserved per CTA) --
// or 0 when the kernel does not apply.
int tma_plan(const SuperDesc &sd, const SearchParams *host_params, int nparams, int *smem_per_chain) {
    bool need_raw = false;
    const int id = tma_match(sd, host_params, nparams, &need_raw);
    *smem_per_chain = 0;
#define PLAN(ID_, BPS_, BW_, BH_, LX_, LY_, CH_, SATD_)                           \
    if (id == ID_) {                                                              \
        using C = Cfg<BPS_, BW_, BH_, LX_, LY_, CH_, SATD_>;                      \
        *smem_per_chain = need_raw ? C::SMEM_RAW : C::SMEM;                       \
        const int by_smem = (227 * 1024) / (*smem_per_chain + 1024);              \
        return by_smem < C::MINB ? by_smem : C::MINB;                             \
    }
    TMA_CONFIGS(PLAN)
#undef PLAN
    return 0;
}

} // namespace mvb200
