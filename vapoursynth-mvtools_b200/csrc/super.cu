// mv.Super on the GPU: builds the padded, multi-level, sub-pel "super" frames of a
// whole batch of source frames, in the reference's exact memory layout
// (MVFrame.cpp:1229-1247 / MVSuper.c:78-103), so any stage's bytes can be handed
// back to VapourSynth or diffed against the oracle without repacking.
//
// All kernels here are streaming integer stencils: one thread per output sample,
// consecutive threads on consecutive samples of a row (coalesced), grid.z = frame.
// They are HBM-bound; the pyramid levels and the three refine passes are separate
// launches in the reference's order because each consumes the previous one's padded
// output (reduce l -> pad l+1 -> reduce l+1 ...; pad 0 -> refine).
//
//   k_fill     <- mvfFillPlane            MVFrame.cpp:1367-1371
//   k_reduce   <- RB2F_C / RB2Filtered / RB2BilinearFiltered / RB2Quadratic / RB2Cubic
//                                         MVFrame.cpp:576-1014 (vertical pass at double
//                                         width + in-place horizontal pass, fused; the
//                                         stale right half of the intermediate is
//                                         reproduced because it is part of the frame)
//   k_pad      <- PadReferenceFrame       MVFrame.cpp:1264-1318
//   k_refine_* <- Horizontal/Vertical/DiagonalBilinear :508-572, *Bicubic :1115-1176,
//                 *Wiener :1019-1111, Average2 :1180-1197, driver mvpRefine :1386-1527
#include "engine.h"

#include <cstdlib>

namespace mvb200 {

namespace {

template <typename T>
__device__ __forceinline__ int ld(const T *p, long long i) {
    return (int)p[i];
}

// ---- fill ---------------------------------------------------------------
template <typename T>
__global__ void k_fill(const uint8_t *__restrict__ src, long long src_pitch, long long src_frame, uint8_t *__restrict__ dst,
                       long long dst_pitch, long long dst_frame, int w, int h) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h)
        return;
    const T *s = (const T *)(src + (long long)blockIdx.z * src_frame + (long long)y * src_pitch);
    T *d = (T *)(dst + (long long)blockIdx.z * dst_frame + (long long)y * dst_pitch);
    d[x] = s[x];
}

// one 32-bit word (4 / 2 samples) per thread; rows of both clips start 4-byte aligned
__global__ void k_fill_w(const uint8_t *__restrict__ src, long long src_pitch, long long src_frame, uint8_t *__restrict__ dst, long long dst_pitch, long long dst_frame,
                         int wwords, int h) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= wwords || y >= h)
        return;
    const unsigned *s = (const unsigned *)(src + (long long)blockIdx.z * src_frame + (long long)y * src_pitch);
    unsigned *d = (unsigned *)(dst + (long long)blockIdx.z * dst_frame + (long long)y * dst_pitch);
    d[x] = __ldg(s + x);
}

// ---- reduce -------------------------------------------------------------
// vertical tap of the separable reduce filters at intermediate column x, output row y
template <typename T>
__device__ __forceinline__ int reduce_v(const T *s, long long sp, int x, int y, int h, int rfilter) {
    const T *r = s + (long long)(2 * y) * sp + x;
    bool edge = (rfilter == 1) ? (y == 0) : (y == 0 || y >= max(h - 1, 1));
    if (edge)
        return (ld(r, 0) + ld(r, sp) + 1) / 2;
    if (rfilter == 1)
        return (ld(r, -sp) + ld(r, 0) * 2 + ld(r, sp) + 2) / 4;
    if (rfilter == 2)
        return (ld(r, -sp) + (ld(r, 0) + ld(r, sp)) * 3 + ld(r, sp * 2) + 4) / 8;
    int m0 = ld(r, -sp * 2), m1 = ld(r, -sp), m2 = ld(r, 0), m3 = ld(r, sp), m4 = ld(r, sp * 2), m5 = ld(r, sp * 3);
    if (rfilter == 3) {
        m2 = (m2 + m3) * 22;
        m1 = (m1 + m4) * 9;
        m0 += m5 + m2 + m1 + 32;
        return m0 >> 6;
    }
    m2 = (m2 + m3) * 10;
    m1 = (m1 + m4) * 5;
    m0 += m5 + m2 + m1 + 16;
    return m0 >> 5;
}

template <typename T>
__global__ void k_reduce(const uint8_t *__restrict__ src_base, uint8_t *__restrict__ dst_base, long long pitch_bytes,
                         long long frame_bytes, int w, int h, int rfilter) {
    // src_base / dst_base: interior origin of level l / l+1 in frame 0.  pitch in bytes.
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    int wide = rfilter == 0 ? w : 2 * w;
    if (x >= wide || y >= h)
        return;
    long long sp = pitch_bytes / (long long)sizeof(T);
    const T *s = (const T *)(src_base + (long long)blockIdx.z * frame_bytes);
    T *d = (T *)(dst_base + (long long)blockIdx.z * frame_bytes) + (long long)y * sp;
    int v;
    if (rfilter == 0) {
        const T *r = s + (long long)(2 * y) * sp + 2 * x;
        v = (ld(r, 0) + ld(r, 1) + ld(r, sp + 1) + ld(r, sp) + 2) / 4;
    } else if (x >= w) {
        v = reduce_v(s, sp, x, y, h, rfilter); // stale intermediate the in-place pass leaves behind
    } else if (x == 0) {
        v = (reduce_v(s, sp, 0, y, h, rfilter) + reduce_v(s, sp, 1, y, h, rfilter) + 1) / 2;
    } else if (rfilter == 1) {
        v = (reduce_v(s, sp, 2 * x - 1, y, h, rfilter) + reduce_v(s, sp, 2 * x, y, h, rfilter) * 2 + reduce_v(s, sp, 2 * x + 1, y, h, rfilter) + 2) / 4;
    } else if (x >= max(w - 1, 1)) {
        v = (reduce_v(s, sp, 2 * x, y, h, rfilter) + reduce_v(s, sp, 2 * x + 1, y, h, rfilter) + 1) / 2;
    } else if (rfilter == 2) {
        v = (reduce_v(s, sp, 2 * x - 1, y, h, rfilter) + (reduce_v(s, sp, 2 * x, y, h, rfilter) + reduce_v(s, sp, 2 * x + 1, y, h, rfilter)) * 3 +
             reduce_v(s, sp, 2 * x + 2, y, h, rfilter) + 4) / 8;
    } else {
        int m0 = reduce_v(s, sp, 2 * x - 2, y, h, rfilter), m1 = reduce_v(s, sp, 2 * x - 1, y, h, rfilter);
        int m2 = reduce_v(s, sp, 2 * x, y, h, rfilter), m3 = reduce_v(s, sp, 2 * x + 1, y, h, rfilter);
        int m4 = reduce_v(s, sp, 2 * x + 2, y, h, rfilter), m5 = reduce_v(s, sp, 2 * x + 3, y, h, rfilter);
        if (rfilter == 3) {
            m2 = (m2 + m3) * 22;
            m1 = (m1 + m4) * 9;
            m0 += m5 + m2 + m1 + 32;
            v = m0 >> 6;
        } else {
            m2 = (m2 + m3) * 10;
            m1 = (m1 + m4) * 5;
            m0 += m5 + m2 + m1 + 16;
            v = m0 >> 5;
        }
    }
    d[x] = (T)v;
}

// ---- pad ----------------------------------------------------------------
template <typename T>
__global__ void k_pad(uint8_t *__restrict__ base, long long pitch_bytes, long long frame_bytes, int hpad, int vpad, int w, int h) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w + 2 * hpad || y >= h + 2 * vpad)
        return;
    bool interior = (x >= hpad && x < hpad + w && y >= vpad && y < vpad + h);
    if (interior)
        return;
    long long sp = pitch_bytes / (long long)sizeof(T);
    T *p = (T *)(base + (long long)blockIdx.z * frame_bytes);
    int sx = min(max(x - hpad, 0), w - 1) + hpad;
    int sy = min(max(y - vpad, 0), h - 1) + vpad;
    p[(long long)y * sp + x] = p[(long long)sy * sp + sx];
}

// Border-only, one 32-bit word per thread.  grid.y covers the 2*vpad full-width border rows first, then the
// h interior rows whose 2*hpad side columns are written.  Requires hpad and w to be whole words.
template <typename T>
__global__ void k_pad_w(uint8_t *__restrict__ base, long long pitch_bytes, long long frame_bytes, int hpad, int vpad, int w, int h) {
    constexpr int SPW = 4 / (int)sizeof(T);
    const int pww = (w + 2 * hpad) / SPW, hpw = hpad / SPW;
    int xw = blockIdx.x * blockDim.x + threadIdx.x;
    int r = blockIdx.y * blockDim.y + threadIdx.y;
    int y;
    if (r < 2 * vpad) {
        y = r < vpad ? r : h + r; // rows [0, vpad) and [vpad + h, 2 vpad + h)
        if (xw >= pww)
            return;
    } else {
        y = r - 2 * vpad + vpad;
        if (y >= vpad + h || xw >= 2 * hpw)
            return;
        xw = xw < hpw ? xw : xw - hpw + (hpad + w) / SPW; // left words, then right words
    }
    uint8_t *fb = base + (long long)blockIdx.z * frame_bytes;
    const int sy = min(max(y - vpad, 0), h - 1) + vpad;
    const T *srow = (const T *)(fb + (long long)sy * pitch_bytes);
    const int x = xw * SPW;
    unsigned v;
    if (x >= hpad && x + SPW <= hpad + w) {
        v = *(const unsigned *)(srow + x);
    } else { // all samples of the word replicate the nearest edge sample
        const unsigned e = (unsigned)srow[x < hpad ? hpad : hpad + w - 1];
        v = sizeof(T) == 1 ? e * 0x01010101u : e * 0x00010001u;
    }
    *(unsigned *)(fb + (long long)y * pitch_bytes + (long long)x * (int)sizeof(T)) = v;
}

// ---- refine -------------------------------------------------------------
template <typename T>
__device__ __forceinline__ int wiener6(int m0, int m1, int m2, int m3, int m4, int m5, int pixelMax) {
    m2 = (m2 + m3) * 4;
    m2 -= m1 + m4;
    m2 *= 5;
    m0 += m5 + m2 + 16;
    m0 >>= 5;
    return max(0, min(m0, pixelMax));
}

// dir: 0 = horizontal (step 1), 1 = vertical (step pitch)
template <typename T>
__global__ void k_refine_hv(uint8_t *__restrict__ dst_base, const uint8_t *__restrict__ src_base, long long pitch_bytes,
                            long long frame_bytes, int w, int h, int bits, int sharp, int dir) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h)
        return;
    long long sp = pitch_bytes / (long long)sizeof(T);
    const T *s = (const T *)(src_base + (long long)blockIdx.z * frame_bytes) + (long long)y * sp + x;
    T *d = (T *)(dst_base + (long long)blockIdx.z * frame_bytes) + (long long)y * sp + x;
    long long st = dir ? sp : 1;
    int i = dir ? y : x, n = dir ? h : w;
    int pixelMax = (1 << bits) - 1;
    int mode; // 0 copy, 1 average, 2 filter
    if (i == n - 1)
        mode = 0;
    else if (sharp == 0)
        mode = 1;
    else if (sharp == 1)
        mode = (i >= 1 && i < n - 3) ? 2 : 1;
    else
        mode = (i >= 2 && i < n - 4) ? 2 : 1;
    int v;
    if (mode == 0)
        v = ld(s, 0);
    else if (mode == 1)
        v = (ld(s, 0) + ld(s, st) + 1) >> 1;
    else if (sharp == 1)
        v = min(pixelMax, max(0, (-ld(s, -st) - ld(s, st * 2) + (ld(s, 0) + ld(s, st)) * 9 + 8) >> 4));
    else
        v = wiener6<T>(ld(s, -st * 2), ld(s, -st), ld(s, 0), ld(s, st), ld(s, st * 2), ld(s, st * 3), pixelMax);
    *d = (T)v;
}

// One 32-bit word of output (4 / 2 samples) per thread; same arithmetic as k_refine_hv.
template <typename T>
__global__ void k_refine_hv_w(uint8_t *__restrict__ dst_base, const uint8_t *__restrict__ src_base, long long pitch_bytes, long long frame_bytes, int w, int h, int bits,
                              int sharp, int dir) {
    constexpr int SPW = 4 / (int)sizeof(T), MASK = sizeof(T) == 1 ? 0xff : 0xffff, SH = 8 * (int)sizeof(T);
    const int xw = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int x = xw * SPW;
    if (x >= w || y >= h)
        return;
    const uint8_t *srow = src_base + (long long)blockIdx.z * frame_bytes + (long long)y * pitch_bytes;
    unsigned *d = (unsigned *)(dst_base + (long long)blockIdx.z * frame_bytes + (long long)y * pitch_bytes) + xw;
    const int pixelMax = (1 << bits) - 1;
    int t[6][SPW]; // the six taps (positions -2..+3) of each output sample
    if (dir == 0) {
        // samples x-SPW*? .. : words xw-2/SPW.. ; gather samples x-2 .. x+SPW+2 from whole words
        constexpr int NW = sizeof(T) == 1 ? 3 : 4; // words covering [x - 2 (rounded down to a word), x + SPW + 3)
        const int w0 = sizeof(T) == 1 ? xw - 1 : xw - 1;
        int smp[NW * SPW];
#pragma unroll
        for (int k = 0; k < NW; k++) {
            const int wi = w0 + k;
            const bool inb = wi >= 0 && wi * SPW < w; // samples outside the row are never used by a filtered output
            const unsigned v = inb ? __ldg((const unsigned *)srow + wi) : 0u;
#pragma unroll
            for (int b = 0; b < SPW; b++)
                smp[k * SPW + b] = (int)((v >> (SH * b)) & MASK);
        }
#pragma unroll
        for (int j = 0; j < SPW; j++)
#pragma unroll
            for (int k = 0; k < 6; k++)
                t[k][j] = smp[SPW + j + k - 2];
    } else {
#pragma unroll
        for (int k = 0; k < 6; k++) {
            const int yy = y + k - 2;
            const bool inb = yy >= 0 && yy < h;
            const unsigned v = inb ? __ldg((const unsigned *)(srow + (long long)(k - 2) * pitch_bytes) + xw) : 0u;
#pragma unroll
            for (int b = 0; b < SPW; b++)
                t[k][b] = (int)((v >> (SH * b)) & MASK);
        }
    }
    unsigned out = 0;
#pragma unroll
    for (int j = 0; j < SPW; j++) {
        const int i = dir ? y : x + j, n = dir ? h : w;
        int mode; // 0 copy, 1 average, 2 filter
        if (i >= n - 1)
            mode = 0;
        else if (sharp == 0)
            mode = 1;
        else if (sharp == 1)
            mode = (i >= 1 && i < n - 3) ? 2 : 1;
        else
            mode = (i >= 2 && i < n - 4) ? 2 : 1;
        int v;
        if (mode == 0)
            v = t[2][j];
        else if (mode == 1)
            v = (t[2][j] + t[3][j] + 1) >> 1;
        else if (sharp == 1)
            v = min(pixelMax, max(0, (-t[1][j] - t[4][j] + (t[2][j] + t[3][j]) * 9 + 8) >> 4));
        else
            v = wiener6<T>(t[0][j], t[1][j], t[2][j], t[3][j], t[4][j], t[5][j], pixelMax);
        out |= (unsigned)v << (SH * j);
    }
    *d = out;
}

// ---- level 0 of a pel-2 super frame in ONE pass -------------------------------------------------------------------------
// mvfFillPlane + PadReferenceFrame + mvpRefine (MVFrame.cpp:1367-1371, 1264-1318, 1386-1494) for level 0: the reference
// copies the source plane into the frame, replicates its borders, then filters the padded plane three times (H from P0,
// V from P0, D from V -- or the bilinear diagonal from P0).  Every sample of the four sub-planes is a pure function of
// the SOURCE plane (a padded sample is the clamped source sample), so one CTA stages a (TH + 5) x (TW + 8) source tile
// in shared memory once and writes the tile of all four sub-planes: the source is read once and nothing is read back
// from the super frame (the separate passes read P0 three times and V once, after a memset of the whole frame).
// One 32-bit word (4 / 2 samples) per thread and step; rows of the super frame start 4-byte aligned and W * bps, hpad * bps
// are whole words (checked by the caller).  The threads of the last tile column also clear the row tails [W * bps, pitch).
template <int SHARP>
__device__ __forceinline__ int refine_tap(int t0, int t1, int t2, int t3, int t4, int t5, int i, int n, int pixelMax) {
    if (i >= n - 1)
        return t2;
    const int avg = (t2 + t3 + 1) >> 1;
    if (SHARP == 0)
        return avg;
    if (SHARP == 1) // HorizontalBicubic / VerticalBicubic, MVFrame.cpp:1115-1176
        return (i >= 1 && i < n - 3) ? min(pixelMax, max(0, (-t1 - t4 + (t2 + t3) * 9 + 8) >> 4)) : avg;
    if (!(i >= 2 && i < n - 4)) // Wiener, MVFrame.cpp:1019-1111
        return avg;
    int m2 = (t2 + t3) * 4;
    m2 -= t1 + t4;
    m2 *= 5;
    const int m0 = (t0 + t5 + m2 + 16) >> 5;
    return max(0, min(m0, pixelMax));
}

template <typename T, int SHARP>
__global__ void __launch_bounds__(256) k_super0_pel2(const uint8_t *__restrict__ src_base, long long src_pitch, long long src_frame, uint8_t *__restrict__ lvl0, long long pitch,
                                                     long long frame_bytes, long long sub_bytes, int w, int h, int hpad, int vpad, int bits) {
    constexpr int SPW = 4 / (int)sizeof(T), SH = 8 * (int)sizeof(T);
    constexpr unsigned MASK = sizeof(T) == 1 ? 0xffu : 0xffffu;
    constexpr int TW = 64, TH = 32;           // samples x rows of one output tile (per sub-plane)
    constexpr int PW = (TW + 8) / SPW;        // words of a staged row: samples x0 - 4 .. x0 + TW + 3
    constexpr int PR = TH + 5;                // staged rows: y0 - 2 .. y0 + TH + 2
    __shared__ unsigned sP[PR][PW + 1];
    __shared__ unsigned sV[TH][PW + 1];
    const int W = w + 2 * hpad, Hh = h + 2 * vpad;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const int pixelMax = (1 << bits) - 1;
    const uint8_t *src = src_base + (long long)blockIdx.z * src_frame;
    uint8_t *out = lvl0 + (long long)blockIdx.z * frame_bytes;
    const int tid = threadIdx.x;

    // ---- phase A: the source tile, borders replicated by clamping the coordinates (also beyond the padded plane: never used)
    for (int k = tid; k < PR * PW; k += 256) {
        const int r = k / PW, cw = k - r * PW;
        const int y = min(max(y0 - 2 + r - vpad, 0), h - 1);
        const int x = x0 - 4 + cw * SPW - hpad; // source column of the word's first sample
        const uint8_t *row = src + (long long)y * src_pitch;
        unsigned v;
        if (x >= 0 && x + SPW <= w) {
            v = __ldg((const unsigned *)(row + (long long)x * (int)sizeof(T)));
        } else {
            v = 0;
#pragma unroll
            for (int j = 0; j < SPW; j++)
                v |= (unsigned)((const T *)row)[min(max(x + j, 0), w - 1)] << (SH * j);
        }
        sP[r][cw] = v;
    }
    __syncthreads();
    // ---- phase B: V over the whole staged width (D filters it horizontally), rows y0 .. y0 + TH - 1
    for (int k = tid; k < TH * PW; k += 256) {
        const int r = k / PW, cw = k - r * PW;
        const int y = y0 + r;
        unsigned t[6];
#pragma unroll
        for (int q = 0; q < 6; q++)
            t[q] = sP[r + q][cw];
        unsigned v = 0;
#pragma unroll
        for (int j = 0; j < SPW; j++) {
            const int a0 = (t[0] >> (SH * j)) & MASK, a1 = (t[1] >> (SH * j)) & MASK, a2 = (t[2] >> (SH * j)) & MASK, a3 = (t[3] >> (SH * j)) & MASK, a4 = (t[4] >> (SH * j)) & MASK,
                      a5 = (t[5] >> (SH * j)) & MASK;
            v |= (unsigned)refine_tap<SHARP>(a0, a1, a2, a3, a4, a5, y, Hh, pixelMax) << (SH * j);
        }
        sV[r][cw] = v;
        const int x = x0 - 4 + cw * SPW;
        if (cw >= 4 / SPW && cw < PW - 4 / SPW && x < W && y < Hh)
            *(unsigned *)(out + 2 * sub_bytes + (long long)y * pitch + (long long)x * (int)sizeof(T)) = v;
    }
    __syncthreads();
    // ---- phase C: P0 (the padded plane itself), H and D.  A thread makes one word of each: the three words around it (12 / 6
    // samples: columns c - SPW .. c + 2 SPW - 1) hold every tap of its samples
    constexpr int OW = TW / SPW;
    for (int k = tid; k < TH * OW; k += 256) {
        const int r = k / OW, ow = k - r * OW;
        const int y = y0 + r, x = x0 + ow * SPW;
        if (x >= W || y >= Hh)
            continue;
        const int cw = 4 / SPW + ow; // staged word of the output word
        constexpr int NS = 3 * SPW + (SPW == 2 ? 2 : 0); // u8: words cw-1..cw+1 (12 samples); u16: words cw-1..cw+2 (8 samples): taps reach c + SPW + 2
        constexpr int NWD = SPW == 2 ? 4 : 3;
        int sp[NS], sv[NS], sn[NS]; // the row of P0, of V, and -- bilinear diagonal only -- the next row of P0
#pragma unroll
        for (int q = 0; q < NWD; q++) {
            const unsigned wp = sP[r + 2][cw - 1 + q];
            const unsigned wv = SHARP == 0 ? 0u : sV[r][cw - 1 + q];
            const unsigned wn = SHARP == 0 ? sP[r + 3][cw - 1 + q] : 0u;
#pragma unroll
            for (int j = 0; j < SPW; j++) {
                sp[q * SPW + j] = (int)((wp >> (SH * j)) & MASK);
                sv[q * SPW + j] = (int)((wv >> (SH * j)) & MASK);
                sn[q * SPW + j] = (int)((wn >> (SH * j)) & MASK);
            }
        }
        unsigned vh = 0, vd = 0;
#pragma unroll
        for (int j = 0; j < SPW; j++) {
            const int c = SPW + j; // index of this sample in sp[] / sv[]
            const int hval = refine_tap<SHARP>(sp[c - 2], sp[c - 1], sp[c], sp[c + 1], sp[c + 2], sp[c + 3], x + j, W, pixelMax);
            int dval;
            if (SHARP == 0) { // DiagonalBilinear, MVFrame.cpp:536-572
                const int p00 = sp[c], p01 = sp[c + 1], p10 = sn[c], p11 = sn[c + 1];
                if (y < Hh - 1)
                    dval = (x + j < W - 1) ? (p00 + p01 + p10 + p11 + 2) >> 2 : (p00 + p10 + 1) >> 1;
                else
                    dval = (x + j < W - 1) ? (p00 + p01 + 1) >> 1 : p00;
            } else {
                dval = refine_tap<SHARP>(sv[c - 2], sv[c - 1], sv[c], sv[c + 1], sv[c + 2], sv[c + 3], x + j, W, pixelMax);
            }
            vh |= (unsigned)hval << (SH * j);
            vd |= (unsigned)dval << (SH * j);
        }
        const long long o = (long long)y * pitch + (long long)x * (int)sizeof(T);
        *(unsigned *)(out + o) = sP[r + 2][cw];
        *(unsigned *)(out + sub_bytes + o) = vh;
        *(unsigned *)(out + 3 * sub_bytes + o) = vd;
    }
    // ---- the row tails of the four sub-planes stay zero (MVSuper.c:75 clears the whole frame)
    if (x0 + TW >= W) {
        const int tail_words = (int)((pitch - (long long)W * (int)sizeof(T)) / 4);
        const int rows = min(TH, Hh - y0);
        for (int k = tid; k < 4 * rows * tail_words; k += 256) {
            const int sp = k / (rows * tail_words), rem = k - sp * rows * tail_words;
            const int r = rem / tail_words, cw = rem - r * tail_words;
            *(unsigned *)(out + sp * sub_bytes + (long long)(y0 + r) * pitch + (long long)W * (int)sizeof(T) + 4ll * cw) = 0u;
        }
    }
}

template <typename T>
__global__ void k_refine_diag(uint8_t *__restrict__ dst_base, const uint8_t *__restrict__ src_base, long long pitch_bytes,
                              long long frame_bytes, int w, int h) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h)
        return;
    long long sp = pitch_bytes / (long long)sizeof(T);
    const T *s = (const T *)(src_base + (long long)blockIdx.z * frame_bytes) + (long long)y * sp + x;
    T *d = (T *)(dst_base + (long long)blockIdx.z * frame_bytes) + (long long)y * sp + x;
    int v;
    if (y < h - 1) {
        if (x < w - 1)
            v = (ld(s, 0) + ld(s, 1) + ld(s, sp) + ld(s, sp + 1) + 2) >> 2;
        else
            v = (ld(s, 0) + ld(s, sp) + 1) >> 1;
    } else {
        if (x < w - 1)
            v = (ld(s, 0) + ld(s, 1) + 1) >> 1;
        else
            v = ld(s, 0);
    }
    *d = (T)v;
}

template <typename T>
__global__ void k_avg2(uint8_t *__restrict__ dst_base, const uint8_t *__restrict__ a_base, const uint8_t *__restrict__ b_base,
                       long long pitch_bytes, long long frame_bytes, int w, int h, int last_column) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w + last_column || y >= h)
        return;
    long long sp = pitch_bytes / (long long)sizeof(T);
    long long o = (long long)y * sp + x;
    const T *a = (const T *)(a_base + (long long)blockIdx.z * frame_bytes);
    const T *b = (const T *)(b_base + (long long)blockIdx.z * frame_bytes);
    T *d = (T *)(dst_base + (long long)blockIdx.z * frame_bytes);
    // last_column: the column after the w averaged ones, as the reference's AVX2 Average2 leaves it in an 8-bit clip --
    // avg(the 0 after plane a's row, b's sample); see the launch site
    d[o] = (T)(((x < w ? ld(a, o) : 0) + ld(b, o) + 1) >> 1);
}

// mvpRefineExtPel2 / Pel4, MVFrame.cpp:1530-1614: sub-plane i = (sx, sy) of level 0 is the (sx, sy) phase of a
// pel-times upsized clip; blockIdx.z = sub-plane index - 1 (the frame loop is on the host: per-frame entry point)
template <typename T>
__global__ void k_refine_ext(uint8_t *__restrict__ level0, long long pitch_bytes, long long sub_bytes, long long dst_off_bytes, const uint8_t *__restrict__ pel_plane,
                             long long pel_pitch, int w, int h, int pel) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h)
        return;
    const int i = blockIdx.z + 1, sx = i % pel, sy = i / pel;
    T *d = (T *)(level0 + (long long)i * sub_bytes + dst_off_bytes + (long long)y * pitch_bytes);
    const T *s = (const T *)(pel_plane + (long long)(y * pel + sy) * pel_pitch);
    d[x] = s[x * pel + sx];
}

inline dim3 grid2d(int w, int h, int n, dim3 blk) { return dim3((w + blk.x - 1) / blk.x, (h + blk.y - 1) / blk.y, n); }

template <typename T>
int super_typed(Ctx *ctx, const mvb200_super_params *p, const SuperDesc &sd, const ClipDesc &cd, const uint8_t *dev_src, uint8_t *dev_super, int n, const PelClip *pelclip) {
    cudaStream_t st = ctx->stream;
    const dim3 blk(64, 4);
    const int bps = sd.bps;
    // MVSuper.c:75 -- the whole super frame starts zeroed (the gaps between levels stay zero).  A plane whose level 0 is
    // written by the fused kernel (every byte of its rows, tails included) only needs its upper levels cleared.
    const bool no_fused = getenv("MVB200_NO_FUSED_SUPER") != nullptr; // (A/B switch: the separate passes)
    auto fused_plane = [&](int plane) {
        const mvb200_level_geom &g = sd.lv[plane][0];
        return !no_fused && p->pel == 2 && !pelclip && (p->modeyuv & (1 << plane)) && (g.width * bps) % 4 == 0 && (g.hpad * bps) % 4 == 0 && g.width >= 4 && g.height >= 1;
    };
    {
        bool all_fused = true;
        for (int plane = 0; plane < sd.nplanes; plane++)
            all_fused = all_fused && fused_plane(plane);
        if (!all_fused) {
            MVB_CUDA_CHECK(cudaMemsetAsync(dev_super, 0, (size_t)sd.frame_bytes * n, st));
        } else {
            for (int plane = 0; plane < sd.nplanes; plane++) {
                const long long begin = sd.offset[plane] + (long long)sd.lv[plane][0].padded_h * 4 * sd.pitch[plane]; // after the four sub-planes of level 0
                const long long end = plane + 1 < sd.nplanes ? sd.offset[plane + 1] : sd.frame_bytes;
                if (end > begin)
                    MVB_CUDA_CHECK(cudaMemset2DAsync(dev_super + begin, (size_t)sd.frame_bytes, 0, (size_t)(end - begin), (size_t)n, st));
            }
        }
        ctx->launches++;
    }
    for (int plane = 0; plane < sd.nplanes; plane++) {
        if (!(p->modeyuv & (1 << plane)))
            continue;
        const long long pitch = sd.pitch[plane];
        uint8_t *pl = dev_super + sd.offset[plane];
        auto level_base = [&](int l) { return pl + (long long)sd.lv[plane][l].row0 * pitch; };
        auto interior = [&](int l) { return level_base(l) + (long long)sd.lv[plane][l].vpad * pitch + (long long)sd.lv[plane][l].hpad * bps; };
        const mvb200_level_geom &g0 = sd.lv[plane][0];
        const bool words0 = (g0.width * bps) % 4 == 0 && (g0.hpad * bps) % 4 == 0; // level 0: whole 32-bit words per row segment
        const bool fused0 = fused_plane(plane);
        if (fused0) {
            const dim3 g((g0.padded_w + 63) / 64, (g0.padded_h + 31) / 32, n);
            const long long sub = pitch * g0.padded_h;
#define MVB_SUPER0(SHARP_)                                                                                                                                                         \
    k_super0_pel2<T, SHARP_><<<g, 256, 0, st>>>(dev_src + cd.offset[plane], cd.pitch[plane], cd.frame_bytes, level_base(0), pitch, sd.frame_bytes, sub, g0.width, g0.height, g0.hpad, \
                                                g0.vpad, sd.bits)
            if (p->sharp == 0)
                MVB_SUPER0(0);
            else if (p->sharp == 1)
                MVB_SUPER0(1);
            else
                MVB_SUPER0(2);
#undef MVB_SUPER0
        } else if (words0)
            k_fill_w<<<grid2d(g0.width * bps / 4, g0.height, n, blk), blk, 0, st>>>(dev_src + cd.offset[plane], cd.pitch[plane], cd.frame_bytes, interior(0), pitch, sd.frame_bytes,
                                                                                  g0.width * bps / 4, g0.height);
        else
            k_fill<T><<<grid2d(g0.width, g0.height, n, blk), blk, 0, st>>>(dev_src + cd.offset[plane], cd.pitch[plane], cd.frame_bytes, interior(0), pitch,
                                                                          sd.frame_bytes, g0.width, g0.height);
        ctx->launches++;
        for (int l = 0; l < sd.levels - 1; l++) {
            const mvb200_level_geom &g1 = sd.lv[plane][l + 1];
            int wide = p->rfilter == 0 ? g1.width : 2 * g1.width;
            k_reduce<T><<<grid2d(wide, g1.height, n, blk), blk, 0, st>>>(interior(l), interior(l + 1), pitch, sd.frame_bytes, g1.width, g1.height, p->rfilter);
            k_pad<T><<<grid2d(g1.padded_w, g1.padded_h, n, blk), blk, 0, st>>>(level_base(l + 1), pitch, sd.frame_bytes, g1.hpad, g1.vpad, g1.width, g1.height);
            ctx->launches += 2;
        }
        if (fused0)
            continue; // level 0 is complete: padded plane and the three half-pel planes
        if (words0)
            k_pad_w<T><<<grid2d(g0.padded_w * bps / 4, 2 * g0.vpad + g0.height, n, blk), blk, 0, st>>>(level_base(0), pitch, sd.frame_bytes, g0.hpad, g0.vpad, g0.width, g0.height);
        else
            k_pad<T><<<grid2d(g0.padded_w, g0.padded_h, n, blk), blk, 0, st>>>(level_base(0), pitch, sd.frame_bytes, g0.hpad, g0.vpad, g0.width, g0.height);
        ctx->launches++;
        if (p->pel > 1 && pelclip) { // mvpRefineExt, MVFrame.cpp:1530-1631 (n == 1)
            const long long sub = pitch * g0.padded_h;
            const long long off = pelclip->padded ? 0 : (long long)g0.vpad * pitch + (long long)g0.hpad * bps;
            dim3 g((g0.width + blk.x - 1) / blk.x, (g0.height + blk.y - 1) / blk.y, p->pel * p->pel - 1);
            k_refine_ext<T><<<g, blk, 0, st>>>(level_base(0), pitch, sub, off, pelclip->dev + pelclip->offset[plane], pelclip->pitch[plane], g0.width, g0.height, p->pel);
            ctx->launches++;
            if (!pelclip->padded)
                for (int i = 1; i < p->pel * p->pel; i++) {
                    k_pad<T><<<grid2d(g0.padded_w, g0.padded_h, 1, blk), blk, 0, st>>>(level_base(0) + sub * i, pitch, sd.frame_bytes, g0.hpad, g0.vpad, g0.width, g0.height);
                    ctx->launches++;
                }
        } else if (p->pel > 1) { // mvpRefine, MVFrame.cpp:1386-1527
            const long long sub = pitch * g0.padded_h;
            auto P = [&](int i) { return level_base(0) + sub * i; };
            const int w = g0.padded_w, h = g0.padded_h;
            const int iH = p->pel == 2 ? 1 : 2, iV = p->pel == 2 ? 2 : 8, iD = p->pel == 2 ? 3 : 10;
            dim3 g = grid2d(w, h, n, blk);
            const bool words = (w * bps) % 4 == 0;
            dim3 gw = grid2d(w * bps / 4, h, n, blk);
            auto refine_hv = [&](int di, int si, int dir) {
                if (words)
                    k_refine_hv_w<T><<<gw, blk, 0, st>>>(P(di), P(si), pitch, sd.frame_bytes, w, h, sd.bits, p->sharp, dir);
                else
                    k_refine_hv<T><<<g, blk, 0, st>>>(P(di), P(si), pitch, sd.frame_bytes, w, h, sd.bits, p->sharp, dir);
            };
            refine_hv(iH, 0, 0);
            refine_hv(iV, 0, 1);
            if (p->sharp == 0)
                k_refine_diag<T><<<g, blk, 0, st>>>(P(iD), P(0), pitch, sd.frame_bytes, w, h);
            else
                refine_hv(iD, iV, 0);
            ctx->launches += 3;
            if (p->pel == 4) {
                // The 3/4-pel planes average w - 1 columns (MVFrame.cpp:1519-1526).  In an 8-bit clip the reference's default
                // path (opt=1: Average2_avx2, MVFrame_AVX2.cpp:19-33, 32-byte steps) also writes the last column unless
                // w - 1 is a multiple of 32: avg(the zeroed stride slack after plane a's row, b's last sample).  The C loop
                // (16-bit clips) leaves it at the frame's initial 0.
                auto avg = [&](int d, const uint8_t *a, const uint8_t *b, int ww, int hh) {
                    const int last = (bps == 1 && ww == w - 1 && (w - 1) % 32 != 0) ? 1 : 0;
                    k_avg2<T><<<grid2d(ww + last, hh, n, blk), blk, 0, st>>>(P(d), a, b, pitch, sd.frame_bytes, ww, hh, last);
                    ctx->launches++;
                };
                avg(1, P(0), P(2), w, h);
                avg(9, P(8), P(10), w, h);
                avg(4, P(0), P(8), w, h);
                avg(6, P(2), P(10), w, h);
                avg(5, P(4), P(6), w, h);
                avg(3, P(0) + bps, P(2), w - 1, h);
                avg(11, P(8) + bps, P(10), w - 1, h);
                avg(12, P(0) + pitch, P(8), w, h - 1);
                avg(14, P(2) + pitch, P(10), w, h - 1);
                avg(13, P(12), P(14), w, h);
                avg(7, P(4) + bps, P(6), w - 1, h);
                avg(15, P(12) + bps, P(14), w - 1, h);
            }
        }
    }
    MVB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

} // namespace

int launch_super(Ctx *ctx, const mvb200_super_params *p, const SuperDesc &sd, const ClipDesc &cd, const uint8_t *dev_src, uint8_t *dev_super, int n, const PelClip *pelclip) {
    if (n <= 0)
        return 0;
    if (pelclip && n != 1)
        return set_error("mvtools_b200: a pelclip is supported by the per-frame entry point only");
    if (sd.bps == 1)
        return super_typed<uint8_t>(ctx, p, sd, cd, dev_src, dev_super, n, pelclip);
    return super_typed<uint16_t>(ctx, p, sd, cd, dev_src, dev_super, n, pelclip);
}

} // namespace mvb200
