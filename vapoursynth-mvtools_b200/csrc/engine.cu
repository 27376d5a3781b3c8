// libmvtools_b200.so: context, device layouts and the C-ABI entry points declared in
// include/mvtools_b200.h (the compute half; the parameter derivation half is
// host_logic.cpp).  There is no CPU path in this library: every compute entry point
// needs a CUDA device of compute capability 10.x and fails loudly otherwise.
#include "engine.h"

#include <algorithm>
#include <functional>

#include <unistd.h>

#include <cstdio>
#include <cstring>
#include <vector>

namespace mvb200 {

static long long round_up(long long v, long long a) { return (v + a - 1) / a * a; }

void make_clip_desc(const mvb200_super_params *sp, ClipDesc *d) {
    memset(d, 0, sizeof(*d));
    d->nplanes = sp->nplanes;
    d->bits = sp->bits;
    d->bps = (sp->bits + 7) / 8;
    long long off = 0;
    for (int p = 0; p < sp->nplanes; p++) {
        d->w[p] = p ? sp->width >> sp->ssw : sp->width;
        d->h[p] = p ? sp->height >> sp->ssh : sp->height;
        d->pitch[p] = round_up((long long)d->w[p] * d->bps, 256);
        d->offset[p] = off;
        off += d->pitch[p] * d->h[p];
    }
    d->frame_bytes = round_up(off, 256);
}

void make_super_desc(const mvb200_super_params *sp, SuperDesc *d) {
    memset(d, 0, sizeof(*d));
    mvb200_super_geom g;
    mvb200_super_geometry(sp, &g);
    d->nplanes = sp->nplanes;
    d->levels = sp->levels;
    d->bits = sp->bits;
    d->bps = (sp->bits + 7) / 8;
    d->xRatioUV = sp->xRatioUV;
    d->yRatioUV = sp->yRatioUV;
    d->logxR = sp->ssw;
    d->logyR = sp->ssh;
    long long off = 0;
    for (int p = 0; p < sp->nplanes; p++) {
        // + 4 bytes of slack so the unaligned-word loads of the search may touch the word
        // after the last sample of a row; pitch multiple of 256
        d->pitch[p] = round_up((long long)g.plane_w[p] * d->bps + 4, 256);
        d->offset[p] = off;
        off += d->pitch[p] * g.plane_h[p];
        for (int l = 0; l < sp->levels && l < kDevMaxLevels; l++)
            d->lv[p][l] = g.lv[p][l];
    }
    d->frame_bytes = round_up(off, 256);
}

void *Ctx::get_scratch(int slot, size_t bytes) {
    if (bytes <= scratch_bytes[slot])
        return scratch[slot];
    // device memory may still be in use by queued kernels or copies: drain first
    cudaStreamSynchronize(stream);
    if (copy_in)
        cudaStreamSynchronize(copy_in);
    if (copy_out)
        cudaStreamSynchronize(copy_out);
    if (scratch[slot])
        cudaFree(scratch[slot]);
    scratch[slot] = nullptr;
    scratch_bytes[slot] = 0;
    size_t want = bytes + std::min<size_t>(bytes / 4, (size_t)256 << 20) + 4096; // (growth slack, bounded: clips of a 4K batch are tens of GB)
    void *p = nullptr;
    if (cudaMalloc(&p, want) != cudaSuccess) {
        set_error("mvtools_b200: out of device memory (%zu bytes of scratch)", want);
        return nullptr;
    }
    if (cudaMemsetAsync(p, 0, want, stream) != cudaSuccess) {
        cudaFree(p);
        set_error("mvtools_b200: cudaMemset failed");
        return nullptr;
    }
    scratch[slot] = p;
    scratch_bytes[slot] = want;
    return p;
}

void *Ctx::get_pinned(size_t bytes) {
    if (bytes <= pinned_bytes)
        return pinned;
    cudaStreamSynchronize(stream);
    if (pinned)
        cudaFreeHost(pinned);
    pinned = nullptr;
    pinned_bytes = 0;
    size_t want = bytes + bytes / 4 + 4096;
    if (cudaMallocHost(&pinned, want) != cudaSuccess) {
        pinned = nullptr;
        set_error("mvtools_b200: cannot allocate %zu bytes of pinned host memory", want);
        return nullptr;
    }
    pinned_bytes = want;
    return pinned;
}

int Ctx::staging_wait() {
    if (staging_ev)
        MVB_CUDA_CHECK(cudaEventSynchronize(staging_ev));
    return 0;
}

int Ctx::staging_posted() {
    if (!staging_ev)
        MVB_CUDA_CHECK(cudaEventCreateWithFlags(&staging_ev, cudaEventDisableTiming));
    MVB_CUDA_CHECK(cudaEventRecord(staging_ev, stream));
    return 0;
}

constexpr int kMaxFramesPerLaunch = 32768;

static int check_levels(const mvb200_super_params *sp) {
    if (sp->levels > kDevMaxLevels)
        return set_error("mvtools_b200: more than %d pyramid levels are not supported on the device", kDevMaxLevels);
    return 0;
}

static void make_search_params(const mvb200_analyse_params *ap, SearchParams *P) {
    memset(P, 0, sizeof(*P));
    const mvb200_analysis_data &ad = ap->ad;
    P->blkW = ad.nBlkSizeX;
    P->blkH = ad.nBlkSizeY;
    P->ovX = ad.nOverlapX;
    P->ovY = ad.nOverlapY;
    P->pel = ad.nPel;
    P->logPel = ad.nPel == 4 ? 2 : (ad.nPel == 2 ? 1 : 0);
    P->levels = ad.nLvCount;
    P->chroma = ap->chroma;
    P->bits = ad.bitsPerSample;
    P->bps = (ad.bitsPerSample + 7) / 8;
    P->search = ap->search;
    P->search_coarse = ap->search_coarse;
    P->nSearchParam = ap->nSearchParam;
    P->nPelSearch = ap->nPelSearch;
    P->nLambda = ap->nLambda;
    P->lsad = ap->lsad;
    P->pnew = ap->pnew;
    P->plevel = ap->plevel;
    P->global = ap->global;
    P->pglobal = ap->pglobal;
    P->pzero = ap->pzero;
    P->badSAD = ap->badSAD;
    P->badrange = ap->badrange;
    P->meander = ap->meander;
    P->tryMany = ap->tryMany;
    P->dctmode = ap->dctmode;
    P->divide = ap->divide;
    P->vectors_size = ap->vectors_size;
    P->verybigSAD = (long long)ad.nBlkSizeX * ad.nBlkSizeY * (1LL << ad.bitsPerSample);
    int off = 8;
    for (int l = ad.nLvCount - 1; l >= 0; l--) { // coarsest first (GroupOfPlanes.c:76-85)
        P->nBlkX[l] = ap->nBlkX[l];
        P->nBlkY[l] = ap->nBlkY[l];
        P->level_off[l] = off;
        off += 4 + ap->nBlkX[l] * ap->nBlkY[l] * 16;
    }
}

static int check_analyse(const mvb200_analyse_params *ap, const mvb200_super_params *sp) {
    if (ap->ad.nLvCount > kDevMaxLevels)
        return set_error("Analyse: more than %d levels are not supported on the device", kDevMaxLevels);
    if (ap->super_pel != sp->pel || ap->super_hpad != sp->hpad || ap->super_vpad != sp->vpad || ap->ad.bitsPerSample != sp->bits)
        return set_error("Analyse: parameters were initialised for a different super clip");
    return check_levels(sp);
}

// mvanalyseGetFrame's choice of the reference frame (MVAnalyse.c:84-109,127-133)
static int ref_frame_of(const mvb200_analyse_params *ap, int n) {
    int nref;
    if (ap->ad.nDeltaFrame > 0)
        nref = n + (ap->ad.isBackward ? ap->ad.nDeltaFrame : -ap->ad.nDeltaFrame);
    else
        nref = -ap->ad.nDeltaFrame;
    return (nref >= 0 && nref < ap->num_frames) ? nref : -1;
}

// field shift with an explicit tff (MVAnalyse.c:146-179); frame props are the plugin's business
static int field_shift_of(const mvb200_analyse_params *ap, int n, int nref) {
    if (!(ap->fields && ap->ad.nPel > 1 && (ap->ad.nDeltaFrame % 2)) || !ap->tff_exists)
        return 0;
    int src_top = ap->tff ^ (n % 2), ref_top = ap->tff ^ (nref % 2);
    return (src_top && !ref_top) ? ap->ad.nPel / 2 : ((ref_top && !src_top) ? -(ap->ad.nPel / 2) : 0);
}

// the same two rules for the consumers of a vector clip (MVRecalculate.c:76-83,168-174; MVCompensate.c:80-87,188-223),
// with an explicit tff; `odd` = the parity test each filter uses (nDeltaFrame % 2, or (nref - n) % 2 in Compensate)
static int ref_frame_of_vd(const mvb200_analysis_data &vd, int num_frames, int n) {
    const int nref = vd.nDeltaFrame > 0 ? n + (vd.isBackward ? vd.nDeltaFrame : -vd.nDeltaFrame) : -vd.nDeltaFrame;
    return (nref >= 0 && nref < num_frames) ? nref : -1;
}
static int field_shift_tff(int fields, int tff_exists, int tff, int pel, bool odd, int n, int nref) {
    if (!(fields && pel > 1 && odd) || !tff_exists)
        return 0;
    const int src_top = tff ^ (n % 2), ref_top = tff ^ (nref % 2);
    return (src_top && !ref_top) ? pel / 2 : ((ref_top && !src_top) ? -(pel / 2) : 0);
}

// mvb200_recalculate_params -> the kernel's parameter block (dct tables are attached by launch_recalc)
static void make_recalc_params(const mvb200_recalculate_params *rp, const mvb200_super_params *sp, RecalcParams *out) {
    RecalcParams &P = *out;
    memset(&P, 0, sizeof(P));
    const mvb200_analysis_data &ad = rp->ad, &od = rp->old;
    P.blkW = ad.nBlkSizeX;
    P.blkH = ad.nBlkSizeY;
    P.ovX = ad.nOverlapX;
    P.ovY = ad.nOverlapY;
    P.logPel = ad.nPel == 4 ? 2 : (ad.nPel == 2 ? 1 : 0);
    P.chroma = rp->chroma;
    P.bps = (ad.bitsPerSample + 7) / 8;
    P.search = rp->search;
    P.nSearchParam = rp->nSearchParam;
    P.nLambda = rp->nLambda;
    P.pnew = rp->pnew;
    P.smooth = rp->smooth;
    P.dctmode = rp->dctmode;
    P.bits = sp->bits;
    P.dct_cos = nullptr;
    P.dct_scratch = nullptr;
    P.dct_scratch_stride = 0;
    P.divide = rp->divide;
    P.thSAD = rp->thSAD;
    P.verybigSAD = (long long)ad.nBlkSizeX * ad.nBlkSizeY * (1LL << ad.bitsPerSample);
    P.nBlkX = ad.nBlkX;
    P.nBlkY = ad.nBlkY;
    P.vectors_size = rp->vectors_size;
    P.oBlkX = od.nBlkX;
    P.oBlkY = od.nBlkY;
    P.oBlkW = od.nBlkSizeX;
    P.oBlkH = od.nBlkSizeY;
    P.oOvX = od.nOverlapX;
    P.oOvY = od.nOverlapY;
    P.oLogPel = od.nPel == 4 ? 2 : (od.nPel == 2 ? 1 : 0);
}

} // namespace mvb200

using namespace mvb200;


extern "C" {

int mvb200_ctx_create(int device, mvb200_ctx **out) {
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0)
        return set_error("mvtools_b200: no CUDA device available (%s); this engine has no CPU fallback", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= count)
        return set_error("mvtools_b200: CUDA device %d does not exist (%d visible)", device, count);
    cudaDeviceProp prop;
    MVB_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return set_error("mvtools_b200: device %d (%s, sm_%d%d) is not a Blackwell sm_100 part; kernels are built for sm_100a only", device, prop.name, prop.major, prop.minor);
    MVB_CUDA_CHECK(cudaSetDevice(device));
    mvb200_ctx *ctx = new mvb200_ctx;
    ctx->c.device = device;
    ctx->c.sm_count = prop.multiProcessorCount;
    cudaError_t se = cudaStreamCreateWithFlags(&ctx->c.stream, cudaStreamNonBlocking);
    if (se != cudaSuccess) {
        delete ctx;
        return set_error("mvtools_b200: cudaStreamCreate failed: %s", cudaGetErrorString(se));
    }
    *out = ctx;
    return 0;
}

int mvb200_device_count(void) {
    int count = 0;
    const cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        set_error("mvtools_b200: no CUDA device available (%s)", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return 0;
    }
    int usable = 0;
    for (int d = 0; d < count; d++) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10)
            usable = d + 1; // (device ordinals are used as they are: count up to the last usable one)
    }
    return usable;
}

int mvb200_window_plan_query(const mvb200_analyse_params *ap, const mvb200_super_params *sp, int nvec, int lookahead, int64_t src_frame_bytes, int64_t out_frame_bytes, int ndevices,
                             int contexts_per_device, int64_t pinned_budget, int64_t device_memory, mvb200_window_plan *out) {
    memset(out, 0, sizeof(*out));
    mvb200_search_plan plan;
    if (mvb200_search_plan_query(ap, sp, 0, &plan))
        return 1;
    if (nvec < 1)
        nvec = 1;
    if (ndevices < 1)
        ndevices = 1;
    if (pinned_budget <= 0) {
        const int64_t phys = (int64_t)sysconf(_SC_PHYS_PAGES) * (int64_t)sysconf(_SC_PAGE_SIZE);
        const int64_t lo = (int64_t)2 << 30, hi = (int64_t)32 << 30;
        pinned_budget = phys / 8 < lo ? lo : (phys / 8 > hi ? hi : phys / 8);
    }
    if (device_memory <= 0)
        device_memory = (int64_t)180 << 30;
    const int64_t wave = plan.chains_per_wave > 0 ? plan.chains_per_wave : 1;
    int frames0 = (int)((wave + nvec - 1) / nvec); // one window fills the chain slots once
    if (frames0 > 768) // (a single vector clip per frame would ask for 1776 frames: bound the latency of a window)
        frames0 = 768;
    mvb200_frame_layout sl;
    mvb200_super_layout(sp, &sl);
    if (out_frame_bytes <= 0)
        out_frame_bytes = src_frame_bytes;
    // device bytes of one context working on windows of fr frames: source / output clips double buffered, the super clip (twice
    // when the output frames are super frames), the vectors, a quarter of growth slack, the searches' histograms
    const int64_t per_frame_dev = (2 * src_frame_bytes + 2 * out_frame_bytes + sl.frame_bytes + (int64_t)nvec * (ap->vectors_size + 256)) * 5 / 4;
    const int64_t fixed_dev = wave * 65536;
    const int64_t per_frame_pinned = src_frame_bytes + out_frame_bytes > 0 ? src_frame_bytes + out_frame_bytes : 1;
    const int64_t dev_limit = device_memory * 3 / 5;
    // frames per window when `per` contexts per device work side by side: within the device limit and the pinned budget (one
    // window per context, one being served, one recycled)
    auto frames_for = [&](int per, int want) -> int {
        int64_t fr = want;
        const int64_t by_dev = (dev_limit / per - fixed_dev) / (per_frame_dev > 0 ? per_frame_dev : 1);
        const int64_t by_pin = pinned_budget / (((int64_t)ndevices * per + 2) * per_frame_pinned);
        fr = fr < by_dev ? fr : by_dev;
        fr = fr < by_pin ? fr : by_pin;
        return (int)(fr < 0 ? 0 : fr);
    };
    int per, frames;
    if (lookahead > 0) { // the window is given: only the number of contexts is adapted to the device
        per = contexts_per_device > 0 ? contexts_per_device : 2;
        frames = lookahead;
        while (per > 1 && per * ((int64_t)frames * per_frame_dev + fixed_dev) > dev_limit)
            per--;
    } else if (contexts_per_device > 0) { // the number of contexts is given (reduced only if the device cannot hold them)
        per = contexts_per_device;
        while (per > 1 && frames_for(per, frames0) < 16)
            per--;
        frames = frames_for(per, frames0);
    } else {
        // automatic: two waves of chains in flight per device are what pays (the second hides the first one's tail and the other
        // stages); a configuration whose window cannot be a whole wave (4K: a window is the pinned budget) gets
        // more, smaller windows on more contexts -- in flight = contexts x window, under the same budgets.  At most 4.
        per = 1;
        frames = frames_for(1, frames0);
        int64_t best = -1;
        for (int c = 1; c <= 4; c++) {
            const int fr = frames_for(c, frames0);
            if (fr < 16 && c > 1)
                break;
            int64_t score = (int64_t)c * fr * nvec;
            score = score < 2 * wave ? score : 2 * wave;
            if (score > best + best / 16) { // (a further context must buy more than 6 % more chains in flight)
                best = score;
                per = c;
                frames = fr;
            }
        }
    }
    if (frames < 16)
        frames = 16;
    out->frames = frames;
    out->contexts_per_device = per;
    out->pinned_budget = pinned_budget;
    return 0;
}

int64_t mvb200_device_memory(int device) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess)
        return 0;
    return (int64_t)prop.totalGlobalMem;
}

void mvb200_ctx_destroy(mvb200_ctx *ctx) {
    if (!ctx)
        return;
    cudaSetDevice(ctx->c.device);
    cudaStreamSynchronize(ctx->c.stream);
    for (auto &p : ctx->c.scratch)
        if (p)
            cudaFree(p);
    if (ctx->c.pinned)
        cudaFreeHost(ctx->c.pinned);
    if (ctx->c.copy_in) {
        cudaStreamSynchronize(ctx->c.copy_in);
        cudaStreamDestroy(ctx->c.copy_in);
    }
    if (ctx->c.copy_out) {
        cudaStreamSynchronize(ctx->c.copy_out);
        cudaStreamDestroy(ctx->c.copy_out);
    }
    for (int i = 0; i < 2; i++) {
        if (ctx->c.ev_h2d[i])
            cudaEventDestroy(ctx->c.ev_h2d[i]);
        if (ctx->c.ev_compute[i])
            cudaEventDestroy(ctx->c.ev_compute[i]);
        if (ctx->c.ev_d2h[i])
            cudaEventDestroy(ctx->c.ev_d2h[i]);
    }
    for (int i = 0; i < kClipChunks; i++) {
        if (ctx->c.ev_piece_in[i])
            cudaEventDestroy(ctx->c.ev_piece_in[i]);
        if (ctx->c.ev_piece_out[i])
            cudaEventDestroy(ctx->c.ev_piece_out[i]);
    }
    if (ctx->c.staging_ev)
        cudaEventDestroy(ctx->c.staging_ev);
    cudaStreamDestroy(ctx->c.stream);
    delete ctx;
}

int mvb200_ctx_synchronize(mvb200_ctx *ctx) {
    MVB_CUDA_CHECK(cudaSetDevice(ctx->c.device));
    MVB_CUDA_CHECK(cudaStreamSynchronize(ctx->c.stream));
    if (ctx->c.copy_in)
        MVB_CUDA_CHECK(cudaStreamSynchronize(ctx->c.copy_in));
    if (ctx->c.copy_out)
        MVB_CUDA_CHECK(cudaStreamSynchronize(ctx->c.copy_out));
    return 0;
}

void *mvb200_ctx_stream(mvb200_ctx *ctx) { return (void *)ctx->c.stream; }

} // extern "C"

struct mvb200_event {
    int device;
    int n;
    cudaEvent_t ev[3];
};

extern "C" {

int mvb200_ctx_record(mvb200_ctx *ctx, mvb200_event **out) {
    *out = nullptr;
    Ctx &c = ctx->c;
    MVB_CUDA_CHECK(cudaSetDevice(c.device));
    mvb200_event *e = new mvb200_event{c.device, 0, {nullptr, nullptr, nullptr}};
    cudaStream_t streams[3] = {c.stream, c.copy_in, c.copy_out};
    for (cudaStream_t s : streams) {
        if (!s)
            continue;
        cudaError_t err = cudaEventCreateWithFlags(&e->ev[e->n], cudaEventDisableTiming);
        if (err == cudaSuccess) {
            e->n++;
            err = cudaEventRecord(e->ev[e->n - 1], s);
        }
        if (err != cudaSuccess) {
            mvb200_event_destroy(e);
            return set_error("mvtools_b200: cannot record an event (%s)", cudaGetErrorString(err));
        }
    }
    *out = e;
    return 0;
}

int mvb200_event_wait(mvb200_event *ev) {
    if (!ev)
        return 0;
    for (int i = 0; i < ev->n; i++)
        MVB_CUDA_CHECK(cudaEventSynchronize(ev->ev[i]));
    return 0;
}

int mvb200_event_query(mvb200_event *ev) {
    if (!ev)
        return 0;
    for (int i = 0; i < ev->n; i++) {
        const cudaError_t e = cudaEventQuery(ev->ev[i]);
        if (e == cudaErrorNotReady)
            return 1;
        if (e != cudaSuccess) {
            set_error("mvtools_b200: CUDA error %s behind an event (%s)", cudaGetErrorName(e), cudaGetErrorString(e));
            return -1;
        }
    }
    return 0;
}

void mvb200_event_destroy(mvb200_event *ev) {
    if (!ev)
        return;
    for (int i = 0; i < ev->n; i++)
        cudaEventDestroy(ev->ev[i]);
    delete ev;
}

int mvb200_search_plan_query(const mvb200_analyse_params *ap, const mvb200_super_params *sp, int sm_count, mvb200_search_plan *out) {
    memset(out, 0, sizeof(*out));
    if (check_analyse(ap, sp))
        return 1;
    SuperDesc sd;
    make_super_desc(sp, &sd);
    SearchParams P;
    make_search_params(ap, &P);
    const int sms = sm_count > 0 ? sm_count : 148;
    const char *force = getenv("MVB200_FORCE_GENERIC");
    const char *notma = getenv("MVB200_NO_TMA");
    const bool generic_only = force && force[0] == '1';
    int smem = 0;
    const int tma = (generic_only || (notma && notma[0] == '1')) ? 0 : tma_plan(sd, &P, 1, &smem);
    if (tma > 0) {
        out->kernel = 2;
        out->chains_per_sm = tma;
        out->smem_per_chain = smem;
    } else {
        // one-warp CTAs without shared memory: the register file (<= 255 registers per thread) holds 8 of them per SM
        // at the least, 16 at the register counts these kernels compile to
        out->kernel = (!generic_only && fast_applies(sd, &P, 1)) ? 1 : 0;
        out->chains_per_sm = 16;
    }
    out->chains_per_wave = out->chains_per_sm * sms;
    return 0;
}

int64_t mvb200_ctx_launch_count(mvb200_ctx *ctx) { return ctx->c.launches; }

void mvb200_clip_layout(const mvb200_super_params *p, mvb200_frame_layout *l) {
    ClipDesc d;
    make_clip_desc(p, &d);
    memset(l, 0, sizeof(*l));
    l->nplanes = d.nplanes;
    l->bytes_per_sample = d.bps;
    for (int i = 0; i < d.nplanes; i++) {
        l->plane_w[i] = d.w[i];
        l->plane_h[i] = d.h[i];
        l->pitch[i] = d.pitch[i];
        l->offset[i] = d.offset[i];
    }
    l->frame_bytes = d.frame_bytes;
}

void mvb200_super_layout(const mvb200_super_params *p, mvb200_frame_layout *l) {
    SuperDesc d;
    make_super_desc(p, &d);
    mvb200_super_geom g;
    mvb200_super_geometry(p, &g);
    memset(l, 0, sizeof(*l));
    l->nplanes = d.nplanes;
    l->bytes_per_sample = d.bps;
    for (int i = 0; i < d.nplanes; i++) {
        l->plane_w[i] = g.plane_w[i];
        l->plane_h[i] = g.plane_h[i];
        l->pitch[i] = d.pitch[i];
        l->offset[i] = d.offset[i];
    }
    l->frame_bytes = d.frame_bytes;
}

// ------------------------------------------------------------- batch path

int mvb200_super_batch(mvb200_ctx *ctx, const mvb200_super_params *p, const void *dev_src, void *dev_super, int n) {
    if (check_levels(p))
        return 1;
    MVB_CUDA_CHECK(cudaSetDevice(ctx->c.device));
    SuperDesc sd;
    ClipDesc cd;
    make_super_desc(p, &sd);
    make_clip_desc(p, &cd);
    // the frame index travels in grid.z / grid.y of the kernels (<= 65535): longer ranges go in pieces
    for (int f0 = 0; f0 < n; f0 += kMaxFramesPerLaunch) {
        const int c = n - f0 < kMaxFramesPerLaunch ? n - f0 : kMaxFramesPerLaunch;
        if (launch_super(&ctx->c, p, sd, cd, (const uint8_t *)dev_src + (long long)f0 * cd.frame_bytes, (uint8_t *)dev_super + (long long)f0 * sd.frame_bytes, c))
            return 1;
    }
    return 0;
}

int mvb200_analyse_multi_batch(mvb200_ctx *ctx, const mvb200_analyse_params *const *ap, int nap, const mvb200_super_params *sp, const void *dev_super, int first, int n, void *const *dev_vectors, int64_t vectors_stride) {
    MVB_CUDA_CHECK(cudaSetDevice(ctx->c.device));
    SuperDesc sd;
    make_super_desc(sp, &sd);
    std::vector<SearchParams> params((size_t)nap);
    std::vector<SearchJob> jobs;
    if (nap <= 0 || n <= 0)
        return 0;
    jobs.reserve((size_t)nap * n);
    for (int i = 0; i < nap; i++) {
        if (check_analyse(ap[i], sp))
            return 1;
        // the instances describe ONE resident super clip: its length sizes the tensor maps and bounds every frame index
        if (ap[i]->num_frames != ap[0]->num_frames)
            return set_error("Analyse: the instances of one launch disagree on the clip length (%d / %d frames)", ap[0]->num_frames, ap[i]->num_frames);
        if (first < 0 || first + n > ap[i]->num_frames)
            return set_error("Analyse: frames [%d, %d) are outside the clip of %d frames", first, first + n, ap[i]->num_frames);
        if (vectors_stride < ap[i]->vectors_size)
            return set_error("Analyse: vectors_stride (%lld) is smaller than the blob (%d bytes)", (long long)vectors_stride, ap[i]->vectors_size);
        make_search_params(ap[i], &params[i]);
    }
    // frame-major order so that chains touching the same super frames are scheduled together
    for (int f = 0; f < n; f++)
        for (int i = 0; i < nap; i++) {
            int fn = first + f;
            int nref = ref_frame_of(ap[i], fn);
            SearchJob j;
            j.src = (const uint8_t *)dev_super + (long long)fn * sd.frame_bytes;
            j.ref = nref >= 0 ? (const uint8_t *)dev_super + (long long)nref * sd.frame_bytes : nullptr;
            j.out = (uint8_t *)dev_vectors[i] + (long long)f * vectors_stride;
            j.param = i;
            j.field_shift = nref >= 0 ? field_shift_of(ap[i], fn, nref) : 0;
            j.src_index = fn;
            j.ref_index = nref >= 0 ? nref : 0;
            jobs.push_back(j);
        }
    mvb200_super_geom geom;
    mvb200_super_geometry(sp, &geom);
    return launch_search(&ctx->c, sd, geom, (const uint8_t *)dev_super, nap > 0 ? ap[0]->num_frames : 0, params.data(), nap, jobs.data(), (int)jobs.size());
}

int mvb200_analyse_batch(mvb200_ctx *ctx, const mvb200_analyse_params *ap, const mvb200_super_params *sp, const void *dev_super, int first, int n, void *dev_vectors, int64_t vectors_stride) {
    const mvb200_analyse_params *aps[1] = {ap};
    void *outs[1] = {dev_vectors};
    return mvb200_analyse_multi_batch(ctx, aps, 1, sp, dev_super, first, n, outs, vectors_stride);
}

int mvb200_degrain_batch(mvb200_ctx *ctx, const mvb200_degrain_params *dp, const mvb200_super_params *sp, const void *dev_src, const void *dev_super, int num_frames, int first, int n, const void *const *dev_vectors, int64_t vectors_stride, void *dev_dst) {
    if (check_levels(sp))
        return 1;
    MVB_CUDA_CHECK(cudaSetDevice(ctx->c.device));
    SuperDesc sd;
    ClipDesc cd;
    make_super_desc(sp, &sd);
    make_clip_desc(sp, &cd);
    for (int f0 = 0; f0 < n; f0 += kMaxFramesPerLaunch) { // (grid.y / grid.z carry the frame index)
        const int c = n - f0 < kMaxFramesPerLaunch ? n - f0 : kMaxFramesPerLaunch;
        const uint8_t *vec[2 * MVB200_MAX_RADIUS];
        for (int r = 0; r < 2 * dp->radius; r++)
            vec[r] = (const uint8_t *)dev_vectors[r] + (long long)f0 * vectors_stride;
        if (launch_degrain(&ctx->c, dp, sd, cd, (const uint8_t *)dev_src, (const uint8_t *)dev_super, num_frames, first + f0, c, vec, vectors_stride, (uint8_t *)dev_dst + (long long)f0 * cd.frame_bytes))
            return 1;
    }
    return 0;
}

// mv.Recalculate / mv.Compensate over frames [first, first + n) of a device-resident super clip (the counterparts of
// mvb200_analyse_multi_batch / mvb200_degrain_batch): the old vector blobs are `old_stride` apart, the new ones
// `vectors_stride`; reference frames and field shifts follow the filters' own rules with an explicit tff.
int mvb200_recalculate_batch(mvb200_ctx *ctx, const mvb200_recalculate_params *rp, const mvb200_super_params *sp, const void *dev_super, int first, int n,
                             const void *dev_old_vectors, int64_t old_stride, void *dev_vectors, int64_t vectors_stride) {
    if (check_levels(sp))
        return 1;
    if (n <= 0)
        return 0;
    if (first < 0 || first + n > rp->num_frames)
        return set_error("Recalculate: frames [%d, %d) are outside the clip of %d frames", first, first + n, rp->num_frames);
    if (vectors_stride < rp->vectors_size)
        return set_error("Recalculate: vectors_stride (%lld) is smaller than the blob (%d bytes)", (long long)vectors_stride, rp->vectors_size);
    Ctx *c = &ctx->c;
    MVB_CUDA_CHECK(cudaSetDevice(c->device));
    SuperDesc sd;
    make_super_desc(sp, &sd);
    RecalcParams P;
    make_recalc_params(rp, sp, &P);
    if (n > kMaxFramesPerLaunch) { // (grid.y carries the frame index)
        for (int f0 = 0; f0 < n; f0 += kMaxFramesPerLaunch)
            if (mvb200_recalculate_batch(ctx, rp, sp, dev_super, first + f0, n - f0 < kMaxFramesPerLaunch ? n - f0 : kMaxFramesPerLaunch, (const uint8_t *)dev_old_vectors + (long long)f0 * old_stride, old_stride,
                                         (uint8_t *)dev_vectors + (long long)f0 * vectors_stride, vectors_stride))
                return 1;
        return 0;
    }
    std::vector<RecalcJob> jobs((size_t)n);
    for (int f = 0; f < n; f++) {
        const int fn = first + f, nref = ref_frame_of_vd(rp->ad, rp->num_frames, fn);
        RecalcJob &j = jobs[(size_t)f];
        j.src = (const uint8_t *)dev_super + (long long)fn * sd.frame_bytes;
        j.ref = nref >= 0 ? (const uint8_t *)dev_super + (long long)nref * sd.frame_bytes : nullptr;
        j.old = (const uint8_t *)dev_old_vectors + (long long)f * old_stride;
        j.out = (uint8_t *)dev_vectors + (long long)f * vectors_stride;
        j.field_shift = nref >= 0 ? field_shift_tff(rp->fields, rp->tff_exists, rp->tff, rp->ad.nPel, rp->ad.nDeltaFrame % 2 != 0, fn, nref) : 0;
    }
    return launch_recalc(c, sd, P, jobs.data(), n);
}

int mvb200_compensate_batch(mvb200_ctx *ctx, const mvb200_compensate_params *cp, const mvb200_super_params *sp, const void *dev_super, int first, int n,
                            const void *dev_vectors, int64_t vectors_stride, void *dev_dst) {
    if (check_levels(sp))
        return 1;
    if (n <= 0)
        return 0;
    if (first < 0 || first + n > cp->num_frames)
        return set_error("Compensate: frames [%d, %d) are outside the clip of %d frames", first, first + n, cp->num_frames);
    Ctx *c = &ctx->c;
    MVB_CUDA_CHECK(cudaSetDevice(c->device));
    SuperDesc sd;
    ClipDesc cd;
    make_super_desc(sp, &sd);
    make_clip_desc(sp, &cd);
    if (!cp->fields) { // no field shifts: the frames of the batch in one launch per plane (chunks of kMaxFramesPerLaunch: grid.z)
        for (int f0 = 0; f0 < n; f0 += kMaxFramesPerLaunch) {
            const int m = std::min(kMaxFramesPerLaunch, n - f0);
            if (launch_compensate_batch(c, cp, sd, cd, (const uint8_t *)dev_super, first + f0, m, (const uint8_t *)dev_vectors + (long long)f0 * vectors_stride, vectors_stride,
                                        (uint8_t *)dev_dst + (long long)f0 * cd.frame_bytes))
                return 1;
        }
        return 0;
    }
    for (int f = 0; f < n; f++) {
        const int fn = first + f, nref = ref_frame_of_vd(cp->vd, cp->num_frames, fn);
        const uint8_t *src = (const uint8_t *)dev_super + (long long)fn * sd.frame_bytes;
        const uint8_t *ref = nref >= 0 ? (const uint8_t *)dev_super + (long long)nref * sd.frame_bytes : nullptr;
        const int fs = nref >= 0 ? field_shift_tff(cp->fields, cp->tff_exists, cp->tff, cp->vd.nPel, (nref - fn) % 2 != 0, fn, nref) : 0;
        if (launch_compensate(c, cp, sd, cd, src, ref, (const uint8_t *)dev_vectors + (long long)f * vectors_stride, fs, (uint8_t *)dev_dst + (long long)f * cd.frame_bytes))
            return 1;
    }
    return 0;
}

// ------------------------------------------------ host-buffer clip pipeline
// Super + nap x Analyse + DegrainN over a whole segment of n frames in HOST memory (clip layout, frames
// frame_bytes apart; pinned memory for truly asynchronous copies); frames [first, first + count) are denoised
// into host_dst (count frames), the others only serve as references (the halo of a sharded clip).  Asynchronous: the call returns when the
// work is enqueued; host_dst is valid after mvb200_ctx_synchronize().  Consecutive calls overlap: the H2D copy
// of segment k+1 and the D2H copy of segment k-1 run on their own streams beside the kernels of segment k
// (source / output clips are double buffered on the device).
int mvb200_degrain_clip_host_async(mvb200_ctx *ctx, const mvb200_super_params *sp, const mvb200_analyse_params *const *ap, int nap, const mvb200_degrain_params *dp,
                                   const void *host_src, void *host_dst, int n, int first, int count) {
    return mvb200_degrain_recalc_clip_host_async(ctx, sp, ap, nap, nullptr, 0, dp, host_src, host_dst, n, first, count);
}

// What the fused clip calls share (clip_pipeline below): a segment of n source frames in HOST memory goes to the device in
// pieces on a copy stream, every frame gets its super frame, `search` enqueues the searches (and Recalculate passes) over the
// super clip, `consume(o0, o1)` enqueues the consumer of output frames [o0, o1) of the range into the (double buffered) device
// output, `download(o0, o1)` enqueues that piece's copies to the host on the other copy stream.  Asynchronous; consecutive calls
// on one context overlap: the H2D copy of segment k+1 and the D2H copy of segment k-1 run beside the kernels of segment k.
struct ClipWork {
    size_t dst_bytes = 0;                                  // device output buffer for the whole range
    std::function<int(uint8_t *dsrc, uint8_t *dsup)> search;
    std::function<int(uint8_t *dsrc, uint8_t *dsup, uint8_t *ddst, int o0, int o1)> consume;
    std::function<int(const uint8_t *ddst, int o0, int o1, cudaStream_t copy_out)> download;
    bool super_is_output = false; // download reads the super clip itself (a vector clip's frames): the super clip is double buffered too
};

static int clip_pipeline(mvb200_ctx *ctx, const mvb200_super_params *sp, const void *host_src, int n, int count, const ClipWork &work) {
    Ctx &c = ctx->c;
    MVB_CUDA_CHECK(cudaSetDevice(c.device));
    if (!c.copy_in) {
        MVB_CUDA_CHECK(cudaStreamCreateWithFlags(&c.copy_in, cudaStreamNonBlocking));
        MVB_CUDA_CHECK(cudaStreamCreateWithFlags(&c.copy_out, cudaStreamNonBlocking));
        for (int i = 0; i < kClipChunks; i++) {
            MVB_CUDA_CHECK(cudaEventCreateWithFlags(&c.ev_piece_in[i], cudaEventDisableTiming));
            MVB_CUDA_CHECK(cudaEventCreateWithFlags(&c.ev_piece_out[i], cudaEventDisableTiming));
        }
        for (int i = 0; i < 2; i++) {
            MVB_CUDA_CHECK(cudaEventCreateWithFlags(&c.ev_h2d[i], cudaEventDisableTiming));
            MVB_CUDA_CHECK(cudaEventCreateWithFlags(&c.ev_compute[i], cudaEventDisableTiming));
            MVB_CUDA_CHECK(cudaEventCreateWithFlags(&c.ev_d2h[i], cudaEventDisableTiming));
        }
    }
    SuperDesc sd;
    ClipDesc cd;
    make_super_desc(sp, &sd);
    make_clip_desc(sp, &cd);
    const int slot = (int)(c.clip_calls & 1);
    const bool reuse = c.clip_calls >= 2; // this slot's buffers were used by the call before the previous one
    c.clip_calls++;
    const size_t clip_bytes = (size_t)cd.frame_bytes * n;
    const int sup_slot = work.super_is_output ? 20 + slot : 12; // (scratch slots: 0-7 kernels' own, 8-14 this pipeline, 16-19 flow.cu, 20-21 here)
    void *before[3] = {c.scratch[8 + slot], c.scratch[10 + slot], c.scratch[sup_slot]};
    uint8_t *dsrc = (uint8_t *)c.get_scratch(8 + slot, clip_bytes);
    uint8_t *ddst = (uint8_t *)c.get_scratch(10 + slot, work.dst_bytes ? work.dst_bytes : 256);
    uint8_t *dsup = (uint8_t *)c.get_scratch(sup_slot, (size_t)sd.frame_bytes * n);
    if (!dsrc || !ddst || !dsup)
        return 1;
    if (before[0] != dsrc || before[1] != ddst || before[2] != dsup)
        MVB_CUDA_CHECK(cudaStreamSynchronize(c.stream)); // fresh buffers are cleared on the compute stream: finish that before the copy stream writes
    // The segment moves in up to kClipChunks pieces so that only a piece's copy -- not the whole segment's -- is exposed
    // before the first kernel and after the last one: Super of piece i runs while piece i+1 is still arriving, and the
    // output frames of piece i leave while piece i+1 is being computed.  (Between consecutive calls the copies already
    // overlap the neighbouring segments' kernels.)
    const int nin = n >= 8 * kClipChunks ? kClipChunks : 1, nout = count >= 8 * kClipChunks ? kClipChunks : 1;
    // H2D: once the kernels that last read this slot's source clip are done
    if (reuse)
        MVB_CUDA_CHECK(cudaStreamWaitEvent(c.copy_in, c.ev_compute[slot], 0));
    // compute: after the D2H that last read this slot's output (and, piece by piece, after this segment's H2D)
    if (reuse)
        MVB_CUDA_CHECK(cudaStreamWaitEvent(c.stream, c.ev_d2h[slot], 0));
    for (int i = 0; i < nin; i++) {
        const int f0 = (int)((long long)n * i / nin), f1 = (int)((long long)n * (i + 1) / nin);
        MVB_CUDA_CHECK(cudaMemcpyAsync(dsrc + (size_t)cd.frame_bytes * f0, (const uint8_t *)host_src + (size_t)cd.frame_bytes * f0, (size_t)cd.frame_bytes * (f1 - f0),
                                       cudaMemcpyHostToDevice, c.copy_in));
        MVB_CUDA_CHECK(cudaEventRecord(c.ev_piece_in[i], c.copy_in));
        MVB_CUDA_CHECK(cudaStreamWaitEvent(c.stream, c.ev_piece_in[i], 0));
        // every frame of the segment gets its super frame: the frames outside the output range are references
        if (mvb200_super_batch(ctx, sp, dsrc + (size_t)cd.frame_bytes * f0, dsup + (size_t)sd.frame_bytes * f0, f1 - f0))
            return 1;
    }
    MVB_CUDA_CHECK(cudaEventRecord(c.ev_h2d[slot], c.copy_in));
    if (work.search && work.search(dsrc, dsup))
        return 1;
    for (int i = 0; i < nout; i++) {
        const int o0 = (int)((long long)count * i / nout), o1 = (int)((long long)count * (i + 1) / nout);
        if (work.consume && work.consume(dsrc, dsup, ddst, o0, o1))
            return 1;
        MVB_CUDA_CHECK(cudaEventRecord(i == nout - 1 ? c.ev_compute[slot] : c.ev_piece_out[i], c.stream));
        // D2H of the piece
        MVB_CUDA_CHECK(cudaStreamWaitEvent(c.copy_out, i == nout - 1 ? c.ev_compute[slot] : c.ev_piece_out[i], 0));
        if (work.download(work.super_is_output ? dsup : ddst, o0, o1, c.copy_out))
            return 1;
    }
    MVB_CUDA_CHECK(cudaEventRecord(c.ev_d2h[slot], c.copy_out));
    return 0;
}

// The same with mv.Recalculate between the searches and the blend: rp[i * nrp + k] (k = 0 .. nrp - 1, NULL = none) are the
// refinement passes applied, in order, to vector clip i (gopRecalculateMVs, MVRecalculate.c:68-254).
int mvb200_degrain_recalc_clip_host_async(mvb200_ctx *ctx, const mvb200_super_params *sp, const mvb200_analyse_params *const *ap, int nap, const mvb200_recalculate_params *const *rp, int nrp,
                                          const mvb200_degrain_params *dp, const void *host_src, void *host_dst, int n, int first, int count) {
    if (check_levels(sp))
        return 1;
    if (n <= 0 || count <= 0)
        return 0;
    if (first < 0 || first + count > n)
        return set_error("mvtools_b200: output range [%d, %d) is outside the segment of %d frames", first, first + count, n);
    if (nap != dp->radius * 2)
        return set_error("mvtools_b200: Degrain%d needs %d vector clips, %d given", dp->radius, dp->radius * 2, nap);
    Ctx &c = ctx->c;
    MVB_CUDA_CHECK(cudaSetDevice(c.device));
    ClipDesc cd;
    make_clip_desc(sp, &cd);
    int64_t vstride = 0;
    bool any_recalc = false;
    for (int i = 0; i < nap; i++) {
        vstride = std::max<int64_t>(vstride, ((int64_t)ap[i]->vectors_size + 255) / 256 * 256);
        for (int k = 0; rp && k < nrp; k++)
            if (rp[i * nrp + k]) {
                vstride = std::max<int64_t>(vstride, ((int64_t)rp[i * nrp + k]->vectors_size + 255) / 256 * 256);
                any_recalc = true;
            }
    }
    uint8_t *dvec = (uint8_t *)c.get_scratch(13, (size_t)vstride * count * nap);
    uint8_t *dvec2 = any_recalc ? (uint8_t *)c.get_scratch(14, (size_t)vstride * count) : nullptr; // one vector clip being refined
    if (!dvec || (any_recalc && !dvec2))
        return 1;
    std::vector<void *> outs((size_t)nap);
    for (int i = 0; i < nap; i++)
        outs[i] = dvec + (size_t)vstride * count * i;
    ClipWork work;
    work.dst_bytes = (size_t)cd.frame_bytes * count;
    work.search = [&](uint8_t *, uint8_t *dsup) -> int {
        if (mvb200_analyse_multi_batch(ctx, ap, nap, sp, dsup, first, count, outs.data(), vstride))
            return 1;
        for (int i = 0; any_recalc && i < nap; i++)
            for (int k = 0; k < nrp; k++) {
                if (!rp[i * nrp + k])
                    continue;
                mvb200_recalculate_params q = *rp[i * nrp + k]; // the segment is the clip as far as the engine is concerned
                q.num_frames = n;
                if (mvb200_recalculate_batch(ctx, &q, sp, dsup, first, count, outs[i], vstride, dvec2, vstride))
                    return 1;
                MVB_CUDA_CHECK(cudaMemcpyAsync(outs[i], dvec2, (size_t)vstride * count, cudaMemcpyDeviceToDevice, c.stream));
            }
        return 0;
    };
    work.consume = [&](uint8_t *dsrc, uint8_t *dsup, uint8_t *ddst, int o0, int o1) -> int {
        std::vector<const void *> vin((size_t)nap);
        for (int k = 0; k < nap; k++)
            vin[k] = (const uint8_t *)outs[k] + (size_t)vstride * o0;
        return mvb200_degrain_batch(ctx, dp, sp, dsrc, dsup, n, first + o0, o1 - o0, vin.data(), vstride, ddst + (size_t)cd.frame_bytes * o0);
    };
    work.download = [&](const uint8_t *ddst, int o0, int o1, cudaStream_t s) -> int {
        MVB_CUDA_CHECK(cudaMemcpyAsync((uint8_t *)host_dst + (size_t)cd.frame_bytes * o0, ddst + (size_t)cd.frame_bytes * o0, (size_t)cd.frame_bytes * (o1 - o0), cudaMemcpyDeviceToHost, s));
        return 0;
    };
    return clip_pipeline(ctx, sp, host_src, n, count, work);
}

void *mvb200_pinned_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) { // portable: the caller's thread may have another current device
        set_error("mvtools_b200: cannot allocate %zu bytes of pinned host memory", bytes);
        return nullptr;
    }
    return p;
}

void mvb200_pinned_free(void *p) {
    if (p)
        cudaFreeHost(p);
}

int mvb200_compensate_clip_host(mvb200_ctx *ctx, const mvb200_super_params *sp, const mvb200_analyse_params *ap, const mvb200_compensate_params *cp, const void *host_src, int src_first, int n, int out_first, int out_count,
                                void *host_dst) {
    return mvb200_compensate_recalc_clip_host(ctx, sp, ap, nullptr, 0, cp, host_src, src_first, n, out_first, out_count, host_dst);
}

// ... with nrp mv.Recalculate passes (rp[0 .. nrp - 1], in order) between the search and the compensation
int mvb200_compensate_recalc_clip_host(mvb200_ctx *ctx, const mvb200_super_params *sp, const mvb200_analyse_params *ap, const mvb200_recalculate_params *const *rp, int nrp, const mvb200_compensate_params *cp,
                                       const void *host_src, int src_first, int n, int out_first, int out_count, void *host_dst) {
    if (mvb200_compensate_recalc_clip_host_async(ctx, sp, ap, rp, nrp, cp, host_src, src_first, n, out_first, out_count, host_dst))
        return 1;
    return mvb200_ctx_synchronize(ctx);
}

// Asynchronous (like mvb200_degrain_recalc_clip_host_async): returns when the work is enqueued; host_dst is valid after
// mvb200_ctx_synchronize() or an event recorded after the call; consecutive calls on one context overlap their copies.
int mvb200_compensate_recalc_clip_host_async(mvb200_ctx *ctx, const mvb200_super_params *sp, const mvb200_analyse_params *ap, const mvb200_recalculate_params *const *rp, int nrp,
                                             const mvb200_compensate_params *cp, const void *host_src, int src_first, int n, int out_first, int out_count, void *host_dst) {
    if (check_levels(sp) || check_analyse(ap, sp))
        return 1;
    if (n <= 0 || out_count <= 0)
        return 0;
    if (out_first < src_first || out_first + out_count > src_first + n)
        return set_error("Compensate: output frames [%d, %d) are outside the segment [%d, %d)", out_first, out_first + out_count, src_first, src_first + n);
    if (ap->ad.nDeltaFrame <= 0 || ap->fields || cp->fields)
        return set_error("Compensate: the clip pipeline handles relative references without field shifts only");
    Ctx *c = &ctx->c;
    MVB_CUDA_CHECK(cudaSetDevice(c->device));
    ClipDesc cd;
    make_clip_desc(sp, &cd);
    mvb200_analyse_params a = *ap; // the segment is the clip as far as the engine is concerned
    a.num_frames = n;
    mvb200_compensate_params q = *cp;
    q.num_frames = n;
    int64_t vstride = ((int64_t)a.vectors_size + 255) / 256 * 256;
    for (int k = 0; rp && k < nrp; k++)
        if (rp[k])
            vstride = std::max<int64_t>(vstride, ((int64_t)rp[k]->vectors_size + 255) / 256 * 256);
    uint8_t *dvec = (uint8_t *)c->get_scratch(13, (size_t)vstride * out_count);
    uint8_t *dvec2 = (uint8_t *)c->get_scratch(14, (size_t)vstride * out_count);
    if (!dvec || !dvec2)
        return 1;
    const int first = out_first - src_first;
    ClipWork work;
    work.dst_bytes = (size_t)cd.frame_bytes * out_count;
    work.search = [&](uint8_t *, uint8_t *dsup) -> int {
        if (mvb200_analyse_batch(ctx, &a, sp, dsup, first, out_count, dvec, vstride))
            return 1;
        for (int k = 0; rp && k < nrp; k++) {
            if (!rp[k])
                continue;
            mvb200_recalculate_params r = *rp[k];
            r.num_frames = n;
            if (mvb200_recalculate_batch(ctx, &r, sp, dsup, first, out_count, dvec, vstride, dvec2, vstride))
                return 1;
            std::swap(dvec, dvec2);
        }
        return 0;
    };
    work.consume = [&](uint8_t *, uint8_t *dsup, uint8_t *ddst, int o0, int o1) -> int {
        return mvb200_compensate_batch(ctx, &q, sp, dsup, first + o0, o1 - o0, dvec + (size_t)vstride * o0, vstride, ddst + (size_t)cd.frame_bytes * o0);
    };
    work.download = [&](const uint8_t *ddst, int o0, int o1, cudaStream_t s) -> int {
        MVB_CUDA_CHECK(cudaMemcpyAsync((uint8_t *)host_dst + (size_t)cd.frame_bytes * o0, ddst + (size_t)cd.frame_bytes * o0, (size_t)cd.frame_bytes * (o1 - o0), cudaMemcpyDeviceToHost, s));
        return 0;
    };
    return clip_pipeline(ctx, sp, host_src, n, out_count, work);
}

// A window of a vector clip from source frames: Super of the segment, Analyse of [out_first, out_first + out_count); the super
// frames and the blobs of the output range go back to the host.
int mvb200_analyse_clip_host(mvb200_ctx *ctx, const mvb200_super_params *sp, const mvb200_analyse_params *ap, const void *host_src, int src_first, int n, int out_first, int out_count,
                             void *host_super, void *host_vectors, int64_t vectors_stride) {
    if (mvb200_analyse_clip_host_async(ctx, sp, ap, host_src, src_first, n, out_first, out_count, host_super, host_vectors, vectors_stride))
        return 1;
    return mvb200_ctx_synchronize(ctx);
}

int mvb200_analyse_clip_host_async(mvb200_ctx *ctx, const mvb200_super_params *sp, const mvb200_analyse_params *ap, const void *host_src, int src_first, int n, int out_first, int out_count,
                                   void *host_super, void *host_vectors, int64_t vectors_stride) {
    if (check_levels(sp) || check_analyse(ap, sp))
        return 1;
    if (n <= 0 || out_count <= 0)
        return 0;
    if (out_first < src_first || out_first + out_count > src_first + n)
        return set_error("Analyse: output frames [%d, %d) are outside the segment [%d, %d)", out_first, out_first + out_count, src_first, src_first + n);
    if (ap->ad.nDeltaFrame <= 0 || ap->fields)
        return set_error("Analyse: the clip pipeline handles relative references without field shifts only");
    if (vectors_stride < ap->vectors_size)
        return set_error("Analyse: vectors_stride (%lld) is smaller than the blob (%d bytes)", (long long)vectors_stride, ap->vectors_size);
    {   // inside the real clip the frames the vectors refer to must be inside the segment (else: silent default vectors)
        const int d = ap->ad.nDeltaFrame, realN = ap->num_frames;
        const int need0 = ap->ad.isBackward ? out_first : (out_first - d > 0 ? out_first - d : 0);
        const int need1 = ap->ad.isBackward ? (out_first + out_count + d < realN ? out_first + out_count + d : realN) : out_first + out_count;
        if (need0 < src_first || need1 > src_first + n)
            return set_error("Analyse: the segment [%d, %d) does not hold the frames [%d, %d) that the vectors of output frames [%d, %d) refer to", src_first, src_first + n, need0, need1, out_first,
                             out_first + out_count);
    }
    Ctx *c = &ctx->c;
    MVB_CUDA_CHECK(cudaSetDevice(c->device));
    SuperDesc sd;
    make_super_desc(sp, &sd);
    mvb200_analyse_params a = *ap; // the segment is the clip as far as the engine is concerned
    a.num_frames = n;
    const int64_t vstride = ((int64_t)a.vectors_size + 255) / 256 * 256;
    const int first = out_first - src_first;
    ClipWork work;
    work.super_is_output = true;                    // the output frames ARE the super frames: nothing to consume,
    work.dst_bytes = (size_t)vstride * out_count;   // the double-buffered output holds the blobs
    uint8_t *dvec_cur = nullptr;
    work.search = [&](uint8_t *, uint8_t *) -> int { return 0; };
    work.consume = [&](uint8_t *, uint8_t *dsup, uint8_t *ddst, int o0, int o1) -> int {
        if (o0 != 0) // (the search is one launch over the whole range: it runs before the first piece leaves)
            return 0;
        dvec_cur = ddst;
        return mvb200_analyse_batch(ctx, &a, sp, dsup, first, out_count, ddst, vstride);
    };
    work.download = [&](const uint8_t *dsup, int o0, int o1, cudaStream_t s) -> int {
        MVB_CUDA_CHECK(cudaMemcpyAsync((uint8_t *)host_super + (size_t)sd.frame_bytes * o0, dsup + (size_t)sd.frame_bytes * (first + o0), (size_t)sd.frame_bytes * (o1 - o0), cudaMemcpyDeviceToHost, s));
        MVB_CUDA_CHECK(cudaMemcpy2DAsync((uint8_t *)host_vectors + (size_t)vectors_stride * o0, (size_t)vectors_stride, dvec_cur + (size_t)vstride * o0, (size_t)vstride, (size_t)a.vectors_size,
                                         (size_t)(o1 - o0), cudaMemcpyDeviceToHost, s));
        return 0;
    };
    return clip_pipeline(ctx, sp, host_src, n, out_count, work);
}

// ------------------------------------------------------- host-buffer path

static int upload_planes(Ctx *c, uint8_t *dev_frame, const long long *pitch, const long long *offset, const int *w, const int *h, int bps, int nplanes, const uint8_t *const src[3], const ptrdiff_t stride[3]) {
    for (int p = 0; p < nplanes; p++) {
        if (!src[p])
            continue;
        MVB_CUDA_CHECK(cudaMemcpy2DAsync(dev_frame + offset[p], (size_t)pitch[p], src[p], (size_t)stride[p], (size_t)w[p] * bps, (size_t)h[p], cudaMemcpyHostToDevice, c->stream));
    }
    return 0;
}

static int download_planes(Ctx *c, const uint8_t *dev_frame, const long long *pitch, const long long *offset, const int *w, const int *h, int bps, int nplanes, uint8_t *const dst[3], const ptrdiff_t stride[3]) {
    for (int p = 0; p < nplanes; p++)
        MVB_CUDA_CHECK(cudaMemcpy2DAsync(dst[p], (size_t)stride[p], dev_frame + offset[p], (size_t)pitch[p], (size_t)w[p] * bps, (size_t)h[p], cudaMemcpyDeviceToHost, c->stream));
    return 0;
}

static void super_plane_dims(const mvb200_super_params *sp, int w[3], int h[3]) {
    for (int p = 0; p < 3; p++) {
        w[p] = p ? sp->super_width >> sp->ssw : sp->super_width;
        h[p] = p ? sp->super_height >> sp->ssh : sp->super_height;
    }
}

int mvb200_super_frame(mvb200_ctx *ctx, const mvb200_super_params *p, const uint8_t *const src[3], const ptrdiff_t src_stride[3], uint8_t *const dst[3], const ptrdiff_t dst_stride[3]) {
    if (check_levels(p))
        return 1;
    Ctx *c = &ctx->c;
    MVB_CUDA_CHECK(cudaSetDevice(c->device));
    SuperDesc sd;
    ClipDesc cd;
    make_super_desc(p, &sd);
    make_clip_desc(p, &cd);
    uint8_t *buf = (uint8_t *)c->get_scratch(5, (size_t)(cd.frame_bytes + sd.frame_bytes));
    if (!buf)
        return 1;
    uint8_t *dsrc = buf, *dsup = buf + cd.frame_bytes;
    if (upload_planes(c, dsrc, cd.pitch, cd.offset, cd.w, cd.h, cd.bps, cd.nplanes, src, src_stride))
        return 1;
    if (launch_super(c, p, sd, cd, dsrc, dsup, 1))
        return 1;
    int w[3], h[3];
    super_plane_dims(p, w, h);
    if (download_planes(c, dsup, sd.pitch, sd.offset, w, h, sd.bps, sd.nplanes, dst, dst_stride))
        return 1;
    MVB_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return 0;
}

// mvsuperGetFrame with a pelclip (MVSuper.c:91-104)
int mvb200_super_frame_pel(mvb200_ctx *ctx, const mvb200_super_params *p, const uint8_t *const src[3], const ptrdiff_t src_stride[3], const uint8_t *const pel[3],
                           const ptrdiff_t pel_stride[3], int pel_padded, uint8_t *const dst[3], const ptrdiff_t dst_stride[3]) {
    if (check_levels(p))
        return 1;
    if (p->pel < 2) // MVSuper.c:240: the pelclip is ignored for pel 1
        return mvb200_super_frame(ctx, p, src, src_stride, dst, dst_stride);
    Ctx *c = &ctx->c;
    MVB_CUDA_CHECK(cudaSetDevice(c->device));
    SuperDesc sd;
    ClipDesc cd;
    make_super_desc(p, &sd);
    make_clip_desc(p, &cd);
    PelClip pc;
    int pw[3], ph[3];
    long long pel_bytes = 0;
    for (int i = 0; i < 3; i++) {
        const int lw = (pel_padded ? p->width + 2 * p->hpad : p->width) * p->pel, lh = (pel_padded ? p->height + 2 * p->vpad : p->height) * p->pel;
        pw[i] = i ? lw >> p->ssw : lw; // the pel clip has the input clip's format (MVSuper.c:232)
        ph[i] = i ? lh >> p->ssh : lh;
        pc.pitch[i] = round_up((long long)pw[i] * cd.bps, 256);
        pc.offset[i] = pel_bytes;
        if (i < cd.nplanes)
            pel_bytes += pc.pitch[i] * ph[i];
    }
    pc.padded = pel_padded;
    uint8_t *buf = (uint8_t *)c->get_scratch(5, (size_t)(cd.frame_bytes + sd.frame_bytes + pel_bytes));
    if (!buf)
        return 1;
    uint8_t *dsrc = buf, *dsup = buf + cd.frame_bytes, *dpel = dsup + sd.frame_bytes;
    pc.dev = dpel;
    if (upload_planes(c, dsrc, cd.pitch, cd.offset, cd.w, cd.h, cd.bps, cd.nplanes, src, src_stride))
        return 1;
    if (upload_planes(c, dpel, pc.pitch, pc.offset, pw, ph, cd.bps, cd.nplanes, pel, pel_stride))
        return 1;
    if (launch_super(c, p, sd, cd, dsrc, dsup, 1, &pc))
        return 1;
    int w[3], h[3];
    super_plane_dims(p, w, h);
    if (download_planes(c, dsup, sd.pitch, sd.offset, w, h, sd.bps, sd.nplanes, dst, dst_stride))
        return 1;
    MVB_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return 0;
}

int mvb200_analyse_frame(mvb200_ctx *ctx, const mvb200_analyse_params *ap, const mvb200_super_params *sp, const uint8_t *const src_super[3], const ptrdiff_t src_stride[3], const uint8_t *const ref_super[3], const ptrdiff_t ref_stride[3], int field_shift, uint8_t *vectors_out) {
    if (check_analyse(ap, sp))
        return 1;
    Ctx *c = &ctx->c;
    MVB_CUDA_CHECK(cudaSetDevice(c->device));
    SuperDesc sd;
    make_super_desc(sp, &sd);
    size_t blob = ((size_t)ap->vectors_size + 255) & ~(size_t)255;
    uint8_t *buf = (uint8_t *)c->get_scratch(5, (size_t)sd.frame_bytes * 2 + blob);
    if (!buf)
        return 1;
    int w[3], h[3];
    super_plane_dims(sp, w, h);
    int nplanes_used = ap->chroma ? sd.nplanes : 1;
    if (upload_planes(c, buf, sd.pitch, sd.offset, w, h, sd.bps, nplanes_used, src_super, src_stride))
        return 1;
    if (ref_super && upload_planes(c, buf + sd.frame_bytes, sd.pitch, sd.offset, w, h, sd.bps, nplanes_used, ref_super, ref_stride))
        return 1;
    SearchParams P;
    make_search_params(ap, &P);
    SearchJob j;
    j.src = buf;
    j.ref = ref_super ? buf + sd.frame_bytes : nullptr;
    j.out = buf + sd.frame_bytes * 2;
    j.param = 0;
    j.field_shift = field_shift;
    j.src_index = 0;
    j.ref_index = 1;
    mvb200_super_geom geom;
    mvb200_super_geometry(sp, &geom);
    if (launch_search(c, sd, geom, buf, 2, &P, 1, &j, 1))
        return 1;
    MVB_CUDA_CHECK(cudaMemcpyAsync(vectors_out, j.out, (size_t)ap->vectors_size, cudaMemcpyDeviceToHost, c->stream));
    MVB_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return 0;
}

int mvb200_analyse_frames(mvb200_ctx *ctx, const mvb200_analyse_params *ap, const mvb200_super_params *sp, int nframes, const uint8_t *const *frames, const ptrdiff_t *strides, int njobs, const int32_t *src_index, const int32_t *ref_index,
                          const int32_t *field_shift, uint8_t *const *vectors_out) {
    if (check_analyse(ap, sp))
        return 1;
    if (njobs <= 0)
        return 0;
    if (nframes <= 0)
        return set_error("Analyse: no frames given");
    for (int k = 0; k < njobs; k++)
        if (src_index[k] < 0 || src_index[k] >= nframes || ref_index[k] >= nframes)
            return set_error("Analyse: job %d refers to a frame outside the %d given", k, nframes);
    Ctx *c = &ctx->c;
    MVB_CUDA_CHECK(cudaSetDevice(c->device));
    SuperDesc sd;
    make_super_desc(sp, &sd);
    const size_t blob = ((size_t)ap->vectors_size + 255) & ~(size_t)255;
    uint8_t *buf = (uint8_t *)c->get_scratch(5, (size_t)sd.frame_bytes * nframes + blob * njobs);
    if (!buf)
        return 1;
    int w[3], h[3];
    super_plane_dims(sp, w, h);
    const int nplanes_used = ap->chroma ? sd.nplanes : 1;
    for (int i = 0; i < nframes; i++)
        if (upload_planes(c, buf + (size_t)i * sd.frame_bytes, sd.pitch, sd.offset, w, h, sd.bps, nplanes_used, frames + 3 * i, strides + 3 * i))
            return 1;
    uint8_t *blobs = buf + (size_t)sd.frame_bytes * nframes;
    SearchParams P;
    make_search_params(ap, &P);
    std::vector<SearchJob> jobs((size_t)njobs);
    for (int k = 0; k < njobs; k++) {
        SearchJob &j = jobs[k];
        j.src = buf + (size_t)src_index[k] * sd.frame_bytes;
        j.ref = ref_index[k] >= 0 ? buf + (size_t)ref_index[k] * sd.frame_bytes : nullptr;
        j.out = blobs + (size_t)k * blob;
        j.param = 0;
        j.field_shift = ref_index[k] >= 0 ? field_shift[k] : 0;
        j.src_index = src_index[k];
        j.ref_index = ref_index[k] >= 0 ? ref_index[k] : 0;
    }
    mvb200_super_geom geom;
    mvb200_super_geometry(sp, &geom);
    if (launch_search(c, sd, geom, buf, nframes, &P, 1, jobs.data(), njobs))
        return 1;
    for (int k = 0; k < njobs; k++)
        MVB_CUDA_CHECK(cudaMemcpyAsync(vectors_out[k], jobs[k].out, (size_t)ap->vectors_size, cudaMemcpyDeviceToHost, c->stream));
    MVB_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return 0;
}

int mvb200_degrain_frame(mvb200_ctx *ctx, const mvb200_degrain_params *dp, const mvb200_super_params *sp, const uint8_t *const src[3], const ptrdiff_t src_stride[3], const uint8_t *const *const ref_super, const ptrdiff_t *ref_stride, const uint8_t *const *vectors, uint8_t *const dst[3], const ptrdiff_t dst_stride[3]) {
    if (check_levels(sp))
        return 1;
    Ctx *c = &ctx->c;
    MVB_CUDA_CHECK(cudaSetDevice(c->device));
    SuperDesc sd;
    ClipDesc cd;
    make_super_desc(sp, &sd);
    make_clip_desc(sp, &cd);
    const int R2 = dp->radius * 2;
    int maxdelta = 0;
    for (int r = 0; r < R2; r++)
        maxdelta = dp->vd[r].nDeltaFrame > maxdelta ? dp->vd[r].nDeltaFrame : maxdelta;
    // a virtual clip of 2*maxdelta+1 super frames with the current frame in the middle, so the
    // batch kernel's "frame n +- delta" addressing finds each reference where it expects it
    const int vframes = 2 * maxdelta + 1;
    // blob size: walk the level headers of the first blob (sizes are stored in the blob itself)
    size_t blob_bytes = (size_t)(*(const int32_t *)vectors[0]);
    size_t blob_stride = (blob_bytes + 255) & ~(size_t)255;
    size_t need = (size_t)sd.frame_bytes * vframes + (size_t)cd.frame_bytes * 2 + blob_stride * R2;
    uint8_t *buf = (uint8_t *)c->get_scratch(5, need);
    if (!buf)
        return 1;
    uint8_t *dsup = buf, *dsrc = buf + (size_t)sd.frame_bytes * vframes, *ddst = dsrc + cd.frame_bytes, *dvec = ddst + cd.frame_bytes;
    int w[3], h[3];
    super_plane_dims(sp, w, h);
    if (upload_planes(c, dsrc, cd.pitch, cd.offset, cd.w, cd.h, cd.bps, cd.nplanes, src, src_stride))
        return 1;
    const void *dv[2 * MVB200_MAX_RADIUS];
    for (int r = 0; r < R2; r++) {
        if ((size_t)(*(const int32_t *)vectors[r]) != blob_bytes)
            return set_error("Degrain: vector blobs of different sizes");
        MVB_CUDA_CHECK(cudaMemcpyAsync(dvec + blob_stride * r, vectors[r], blob_bytes, cudaMemcpyHostToDevice, c->stream));
        dv[r] = dvec + blob_stride * r;
        const uint8_t *const *rp = ref_super + r * 3;
        if (rp[0]) {
            int idx = maxdelta + (dp->vd[r].isBackward ? dp->vd[r].nDeltaFrame : -dp->vd[r].nDeltaFrame);
            if (upload_planes(c, dsup + (size_t)sd.frame_bytes * idx, sd.pitch, sd.offset, w, h, sd.bps, sd.nplanes, rp, ref_stride + r * 3))
                return 1;
        }
    }
    // `first` = maxdelta in the virtual clip; the source frame pointer is biased accordingly
    if (launch_degrain(c, dp, sd, cd, dsrc - (long long)maxdelta * cd.frame_bytes, dsup, vframes, maxdelta, 1, (const uint8_t *const *)dv, (long long)blob_stride, ddst))
        return 1;
    if (download_planes(c, ddst, cd.pitch, cd.offset, cd.w, cd.h, cd.bps, cd.nplanes, dst, dst_stride))
        return 1;
    MVB_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return 0;
}

int mvb200_compensate_frame(mvb200_ctx *ctx, const mvb200_compensate_params *cp, const mvb200_super_params *sp, const uint8_t *const src_super[3], const ptrdiff_t src_stride[3], const uint8_t *const ref_super[3], const ptrdiff_t ref_stride[3], const uint8_t *vectors, int field_shift, uint8_t *const dst[3], const ptrdiff_t dst_stride[3]) {
    if (check_levels(sp))
        return 1;
    Ctx *c = &ctx->c;
    MVB_CUDA_CHECK(cudaSetDevice(c->device));
    SuperDesc sd;
    ClipDesc cd;
    make_super_desc(sp, &sd);
    make_clip_desc(sp, &cd);
    const size_t blob_bytes = (size_t)(*(const int32_t *)vectors), blob_stride = (blob_bytes + 255) & ~(size_t)255;
    uint8_t *buf = (uint8_t *)c->get_scratch(5, (size_t)sd.frame_bytes * 2 + (size_t)cd.frame_bytes + blob_stride);
    if (!buf)
        return 1;
    uint8_t *dsrc = buf, *dref = buf + sd.frame_bytes, *ddst = dref + sd.frame_bytes, *dvec = ddst + cd.frame_bytes;
    int w[3], h[3];
    super_plane_dims(sp, w, h);
    const int num_planes = (cp->super_modeyuv & (MVB200_UPLANE | MVB200_VPLANE)) ? sd.nplanes : 1;
    if (upload_planes(c, dsrc, sd.pitch, sd.offset, w, h, sd.bps, num_planes, src_super, src_stride))
        return 1;
    if (ref_super && upload_planes(c, dref, sd.pitch, sd.offset, w, h, sd.bps, num_planes, ref_super, ref_stride))
        return 1;
    MVB_CUDA_CHECK(cudaMemcpyAsync(dvec, vectors, blob_bytes, cudaMemcpyHostToDevice, c->stream));
    if (launch_compensate(c, cp, sd, cd, dsrc, ref_super ? dref : nullptr, dvec, field_shift, ddst))
        return 1;
    // only the planes of the super clip's modeyuv are produced (the reference leaves the others untouched)
    if (download_planes(c, ddst, cd.pitch, cd.offset, cd.w, cd.h, cd.bps, num_planes, dst, dst_stride))
        return 1;
    MVB_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return 0;
}

int mvb200_recalculate_frame(mvb200_ctx *ctx, const mvb200_recalculate_params *rp, const mvb200_super_params *sp, const uint8_t *const src_super[3], const ptrdiff_t src_stride[3], const uint8_t *const ref_super[3], const ptrdiff_t ref_stride[3], const uint8_t *old_vectors, int field_shift, uint8_t *vectors_out) {
    if (check_levels(sp))
        return 1;
    Ctx *c = &ctx->c;
    MVB_CUDA_CHECK(cudaSetDevice(c->device));
    SuperDesc sd;
    make_super_desc(sp, &sd);
    const size_t old_bytes = (size_t)(*(const int32_t *)old_vectors), old_stride = (old_bytes + 255) & ~(size_t)255;
    const size_t new_stride = ((size_t)rp->vectors_size + 255) & ~(size_t)255;
    uint8_t *buf = (uint8_t *)c->get_scratch(5, (size_t)sd.frame_bytes * 2 + old_stride + new_stride);
    if (!buf)
        return 1;
    int w[3], h[3];
    super_plane_dims(sp, w, h);
    const int nplanes_used = rp->chroma ? sd.nplanes : 1;
    if (upload_planes(c, buf, sd.pitch, sd.offset, w, h, sd.bps, nplanes_used, src_super, src_stride))
        return 1;
    if (ref_super && upload_planes(c, buf + sd.frame_bytes, sd.pitch, sd.offset, w, h, sd.bps, nplanes_used, ref_super, ref_stride))
        return 1;
    uint8_t *dold = buf + sd.frame_bytes * 2, *dnew = dold + old_stride;
    MVB_CUDA_CHECK(cudaMemcpyAsync(dold, old_vectors, old_bytes, cudaMemcpyHostToDevice, c->stream));
    RecalcParams P;
    make_recalc_params(rp, sp, &P);
    RecalcJob j;
    j.src = buf;
    j.ref = ref_super ? buf + sd.frame_bytes : nullptr;
    j.old = dold;
    j.out = dnew;
    j.field_shift = field_shift;
    if (launch_recalc(c, sd, P, &j, 1))
        return 1;
    MVB_CUDA_CHECK(cudaMemcpyAsync(vectors_out, dnew, (size_t)rp->vectors_size, cudaMemcpyDeviceToHost, c->stream));
    MVB_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return 0;
}

} // extern "C"
