// Internal header of libmvtools_b200.so (not installed; the public ABI is
// include/mvtools_b200.h).
#pragma once

#include "../../include/mvtools_b200.h"

#include <cstddef>
#include <cstdint>

#ifdef __CUDACC__
#include <cuda_runtime.h>
#endif

namespace mvb200 {

// sets the thread-local message returned by mvb200_last_error(); returns 1
int set_error(const char *fmt, ...) __attribute__((format(printf, 1, 2)));

int plane_dim_luma(int src, int level, int ratio, int pad);

constexpr int kDevMaxLevels = 16;

// Device-side description of where every (plane, level) of a super frame lives.
// Frames are `frame_bytes` apart; plane p starts `offset[p]` bytes into a frame and
// has row pitch `pitch[p]` bytes (mvb200_super_layout).
struct SuperDesc {
    int nplanes, levels, bps, bits;
    int xRatioUV, yRatioUV, logxR, logyR;
    long long pitch[3], offset[3], frame_bytes;
    mvb200_level_geom lv[3][kDevMaxLevels];
};

// Source / output clip layout on the device (mvb200_clip_layout).
struct ClipDesc {
    int nplanes, bps, bits;
    int w[3], h[3];
    long long pitch[3], offset[3], frame_bytes;
};

void make_super_desc(const mvb200_super_params *sp, SuperDesc *d);
void make_clip_desc(const mvb200_super_params *sp, ClipDesc *d);

#ifdef __CUDACC__

// Parameters of one mv.Analyse instance as the search kernel consumes them.
struct SearchParams {
    int blkW, blkH, ovX, ovY, pel, logPel, levels; // levels = nLvCount
    int chroma, bps, bits;
    int search, search_coarse, nSearchParam, nPelSearch;
    int nLambda, lsad, pnew, plevel, global, pglobal, pzero;
    long long badSAD;
    int badrange, meander, tryMany, dctmode, divide;
    int nBlkX[kDevMaxLevels], nBlkY[kDevMaxLevels];
    int level_off[kDevMaxLevels]; // byte offset of level l's size header inside the blob
    int vectors_size;
    long long verybigSAD;
    // dct 1..4 (DCTFFTW.cpp): cosine tables (blkW^2 then blkH^2 floats) and the per-chain scratch
    // (blkW*blkH floats of row-pass results + blkW*blkH ints holding the source block's DCT image)
    const float *dct_cos;
    uint8_t *dct_scratch;
    long long dct_scratch_stride;
};

// One search chain = one (analyse instance, frame): walked by one warp.
struct SearchJob {
    const uint8_t *src; // super frame n (device)
    const uint8_t *ref; // super frame nref, or nullptr (-> default vectors, validity 0)
    uint8_t *out;       // blob
    int param;          // index into the SearchParams array
    int field_shift;
    int src_index, ref_index; // frame indices relative to the super-clip base (TMA coordinates)
};

// mv.Recalculate: parameters and one job per frame
struct RecalcParams {
    int blkW, blkH, ovX, ovY, logPel, chroma, bps, search, nSearchParam, nLambda, pnew, smooth, dctmode, divide;
    long long thSAD, verybigSAD;
    int nBlkX, nBlkY, vectors_size;
    int oBlkX, oBlkY, oBlkW, oBlkH, oOvX, oOvY, oLogPel; // geometry of the OLD vector field (its finest level)
    int bits;
    const float *dct_cos; // as in SearchParams; the scratch is per (block, frame) here
    uint8_t *dct_scratch;
    long long dct_scratch_stride;
};
struct RecalcJob {
    const uint8_t *src, *ref; // super frames (ref may be null)
    const uint8_t *old;       // old blob
    uint8_t *out;             // new blob
    int field_shift;
};

struct Ctx;

// kernels' host launchers (each returns 0 or sets the error)
// one frame of a pel-times upsized clip on the device (mv.Super's pelclip argument)
struct PelClip {
    const uint8_t *dev;
    long long pitch[3], offset[3];
    int padded; // the pel clip already carries the padding (MVSuper.c:246-248)
};
int launch_super(Ctx *ctx, const mvb200_super_params *p, const SuperDesc &sd, const ClipDesc &cd, const uint8_t *dev_src, uint8_t *dev_super, int n, const PelClip *pelclip = nullptr);
// frames_base / nframes describe the resident super clip the jobs point into (needed for the TMA descriptors)
int launch_search(Ctx *ctx, const SuperDesc &sd, const mvb200_super_geom &geom, const uint8_t *frames_base, int nframes, const SearchParams *host_params, int nparams, const SearchJob *host_jobs, int njobs);
int launch_search_tma(Ctx *ctx, const SuperDesc &sd, const mvb200_super_geom &geom, const uint8_t *frames_base, int nframes, const SearchParams *host_params, int nparams,
                      const SuperDesc *dev_sd, const SearchParams *dev_params, const SearchJob *dev_jobs, int njobs, int *hist, int hist_ints, bool *handled);
int launch_search_fast(Ctx *ctx, const SuperDesc &sd, const SearchParams *host_params, int nparams, const SuperDesc *dev_sd, const SearchParams *dev_params,
                       const SearchJob *dev_jobs, int njobs, int *hist, int hist_ints, bool *handled);
// pure host (mvb200_search_plan_query): chains per SM of the TMA kernel for these parameter sets (0: it does not apply) and
// whether the register-blocked kernel applies
int tma_plan(const SuperDesc &sd, const SearchParams *host_params, int nparams, int *smem_per_chain);
bool fast_applies(const SuperDesc &sd, const SearchParams *host_params, int nparams);
int launch_recalc(Ctx *ctx, const SuperDesc &sd, const RecalcParams &P, const RecalcJob *host_jobs, int njobs);
int launch_recalc_tma(Ctx *ctx, const SuperDesc &sd, const RecalcParams &P, const SuperDesc *dev_sd, const RecalcParams *dev_params, const RecalcJob *dev_jobs, int njobs, bool *handled);
int launch_compensate(Ctx *ctx, const mvb200_compensate_params *cp, const SuperDesc &sd, const ClipDesc &cd, const uint8_t *dev_src_super, const uint8_t *dev_ref_super, const uint8_t *dev_vectors, int field_shift, uint8_t *dev_dst);
int launch_compensate_batch(Ctx *ctx, const mvb200_compensate_params *cp, const SuperDesc &sd, const ClipDesc &cd, const uint8_t *dev_super, int first, int n, const uint8_t *dev_vectors, long long vectors_stride, uint8_t *dev_dst);
int launch_degrain(Ctx *ctx, const mvb200_degrain_params *dp, const SuperDesc &sd, const ClipDesc &cd, const uint8_t *dev_src, const uint8_t *dev_super, int num_frames, int first, int n, const uint8_t *const *dev_vectors, long long vectors_stride, uint8_t *dev_dst);

constexpr int kClipChunks = 4;

struct Ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    long long launches = 0;
    // growable device scratch buffers.  Every user owns its slot numbers (a shared slot means one call's data under another's
    // cached pointers): 0, 1, 7 search.cu (jobs / histogram / vectors); 2, 3, 4 degrain.cu; 2, 6 compensate.cu; 5 the per-frame and
    // windowed entry points of engine.cu; 8-14 clip_pipeline (8/9 source clip, 10/11 output, 12 super clip, 13/14 vectors);
    // 16, 17, 18, 19 flow.cu (tables -- resident across calls --, small fields, staging, windows); 20/21 clip_pipeline's double
    // buffered super clip of a vector-clip window; 15, 22, 23 free
    void *scratch[24] = {};
    size_t scratch_bytes[24] = {};
    // pinned host staging
    void *pinned = nullptr;
    size_t pinned_bytes = 0;
    int sm_count = 0;
    // the pinned staging buffer is reused from call to call: an event marks the last upload out of it,
    // so that the next writer waits for that copy only (not for the kernels queued behind it)
    cudaEvent_t staging_ev = nullptr;
    // overlap windows of the last Degrain configuration, resident on the device
    int win_key[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    short *win_dev[2] = {nullptr, nullptr};
    bool win_valid = false;
    // ... and of the last Compensate configuration
    int cwin_key[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    short *cwin_dev[2] = {nullptr, nullptr};
    bool cwin_valid = false;
    // SimpleResize tables of the last Mask / Flow configuration (flow.cu), resident on the device
    int rz_key[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int *rz_dev = nullptr;
    bool rz_valid = false;
    // overlap windows of the last BlockFPS configuration
    int bfw_key[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    short *bfw_dev[2] = {nullptr, nullptr};
    bool bfw_valid = false;
    // host-buffer clip pipeline (mvb200_degrain_clip_host_async): copy streams + double-buffered clips
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_compute[2] = {nullptr, nullptr}, ev_d2h[2] = {nullptr, nullptr};
    cudaEvent_t ev_piece_in[4] = {}, ev_piece_out[4] = {}; // a segment's copies move in kClipChunks pieces
    long long clip_calls = 0;

    void *get_scratch(int slot, size_t bytes); // nullptr + error on failure
    void *get_pinned(size_t bytes);
    int staging_wait();   // before the host writes the pinned staging buffer
    int staging_posted(); // after the upload out of it was enqueued on `stream`
};

// Upload of a small host-built table (overlap windows, resize tables) into device scratch, complete on return.  Through the
// context's stream: a synchronous cudaMemcpy from pageable memory may return before its DMA has reached the device, and a
// kernel on this NON-BLOCKING stream is not ordered behind the legacy stream's copy -- that was the rare wrong first frame of a
// configuration with large tables (blk 128x64: 147 KB of windows).  -DMVB_SYNC_TABLE_UPLOAD rebuilds the old behaviour for
// tools/stress_tables.py.
inline cudaError_t upload_table(Ctx *c, void *dev, const void *host, size_t bytes) {
#ifdef MVB_SYNC_TABLE_UPLOAD
    return cudaMemcpy(dev, host, bytes, cudaMemcpyHostToDevice);
#else
    cudaError_t e = cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, c->stream);
    return e != cudaSuccess ? e : cudaStreamSynchronize(c->stream);
#endif
}

#define MVB_CUDA_CHECK(call)                                                                          \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return ::mvb200::set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorName(e_), __FILE__, \
                                       __LINE__, cudaGetErrorString(e_));                             \
    } while (0)

#endif // __CUDACC__

} // namespace mvb200

#ifdef __CUDACC__
struct mvb200_ctx { // the opaque handle of the C ABI
    mvb200::Ctx c;
};
#endif
