// libmvtools.so -- the drop-in VapourSynth plugin for the hot path.
//
// Exports exactly what the reference's module exports (EntryPoint.c:28-50):
// VapourSynthPluginInit2, which configures the plugin as "com.nodame.mvtools" / namespace
// "mv" and registers Super, Analyse, Recalculate, Compensate, Degrain1..6, Mask, Finest, Flow, FlowInter, FlowFPS, FlowBlur, BlockFPS and SCDetection with the reference's
// byte-identical argument strings, so existing .vpy scripts load unchanged.  Every filter keeps the reference's
// create-time validation (messages included, they come from mvb200_*_init), frame
// properties (Super_*, MVTools_MVAnalysisData, MVTools_vectors) and request pattern; the
// pixel work is done by libmvtools_b200.so on the GPU through the C ABI in
// include/mvtools_b200.h.  There is no CPU code path here: if the CUDA context cannot be
// created the filter's frame request fails with that error.
//
// VapourSynth calls getFrame concurrently (fmParallel).  The reference spends every host core on each frame; a search on the GPU
// is one warp per frame and vector clip, so this plugin works in WINDOWS of frames instead (fusedGetFrame below): consumers fed by
// this plugin's own Super / Analyse / Recalculate nodes produce a window of output frames with one engine call from the source
// frames, keep the next windows in flight on two engine contexts per GPU and on every visible GPU, and serve the frame server's
// threads from the windows' pinned buffers.  Everything else is evaluated node by node on engine context 0 (mv.Analyse with its
// own look-ahead window), each engine context guarded by its own mutex (a context is not re-entrant).
#include <VapourSynth4.h>
#include <VSHelper4.h>

#include <mvtools_b200.h>

#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include <unistd.h>

namespace {

// Engine contexts (own CUDA streams and scratch each), every one guarded by its own mutex.  Slot 0 does the per-frame work;
// the fused windows go round the whole list (frames are independent units: a window carries its own halo).
//   MVB200_DEVICES=0,2,3  the devices to use        MVB200_DEVICE=1  a single device (what a multi-process launcher sets per rank)
//   neither: every device the engine can run on
// Each device gets MVB200_CONTEXTS_PER_DEVICE contexts (default 2): consecutive windows then run CONCURRENTLY on one GPU -- one
// window's copies, Super and DegrainN kernels beside the other's search, and the next window's search chains take the slots the
// moment the slow chains of the previous one leave them.  Steady state of the 1080p Degrain3 script over 60 windows
// (profiles/r2_plugin_long_runs.txt): 1298 / 1493 / 1491 frames/s with 1 / 2 / 4 contexts -- two are enough, more only cost
// device memory.  A device named several times in MVB200_DEVICES gets exactly that many contexts.
constexpr int kMaxDevices = 16;
struct EngineSlot {
    std::mutex m; // serialises this context's calls (the engine context is not re-entrant)
    mvb200_ctx *ctx = nullptr;
    int device = 0;
    bool failed = false;
    std::string err;
};
EngineSlot g_engines[kMaxDevices * 8];
std::once_flag g_devices_once;
int g_ndev = 1;      // engine slots in all
int g_ndevices = 1;  // distinct positions of the device list (slots are laid out context-major: [context][device])
int g_perdev = 1;    // contexts per device the slots are laid out for (8 unless the device list repeats devices)
int g_perdev_asked = 0; // MVB200_CONTEXTS_PER_DEVICE (0: unset -- the window plan chooses, mvb200_window_plan_query)

void listDevices() {
    int n = 0;
    if (const char *list = getenv("MVB200_DEVICES")) {
        for (const char *p = list; *p && n < kMaxDevices;) {
            char *end = nullptr;
            const long d = strtol(p, &end, 10);
            if (end == p)
                break;
            g_engines[n++].device = (int)d;
            p = *end == ',' ? end + 1 : end;
        }
    } else if (const char *one = getenv("MVB200_DEVICE")) {
        g_engines[n++].device = atoi(one);
    } else {
        const int count = mvb200_device_count();
        for (int d = 0; d < count && n < kMaxDevices; d++)
            g_engines[n++].device = d;
    }
    if (n == 0)
        g_engines[n++].device = 0; // (no usable device: creating the context reports why)
    g_ndevices = n;
    int per = 8; // slots for up to 8 contexts per device; a filter uses as many as its window plan says
    if (const char *c = getenv("MVB200_CONTEXTS_PER_DEVICE")) {
        g_perdev_asked = atoi(c) < 1 ? 1 : (atoi(c) > 8 ? 8 : atoi(c));
        per = g_perdev_asked;
    }
    if (getenv("MVB200_DEVICES")) { // an explicit list says how many contexts each device gets
        bool repeated = false;
        for (int i = 0; i < n; i++)
            for (int j = 0; j < i; j++)
                repeated = repeated || g_engines[i].device == g_engines[j].device;
        if (repeated)
            per = g_perdev_asked = 1;
    }
    for (int c = 1; c < per; c++)
        for (int d = 0; d < n; d++)
            g_engines[c * n + d].device = g_engines[d].device;
    g_perdev = per;
    g_ndev = n * per;
}

int deviceCount() {
    std::call_once(g_devices_once, listDevices);
    return g_ndev;
}

// The context of slot `which` (created on first use), with the slot's mutex to hold around every call on it;
// returns nullptr and fills err on failure.  Call with the slot's mutex held.
mvb200_ctx *engineOf(EngineSlot &e, std::string &err) {
    if (!e.ctx && !e.failed) {
        if (mvb200_ctx_create(e.device, &e.ctx)) {
            e.err = mvb200_last_error();
            e.ctx = nullptr;
            e.failed = true;
        }
    }
    if (!e.ctx)
        err = e.err;
    return e.ctx;
}

// the per-frame paths: device slot 0
std::mutex &g_engine_mutex = g_engines[0].m;
mvb200_ctx *engine(std::string &err) {
    deviceCount();
    return engineOf(g_engines[0], err);
}

// The plugin's own Super / Analyse nodes, by node pointer (API 4: a VSNode* IS the node, every reference to it has the
// same value).  mv.DegrainN looks its inputs up here to recognise the canonical script
//     sup = Super(clip); v.. = Analyse(sup, ..); DegrainN(clip, sup, v..)
// whose whole chain one engine call can run (see DegrainData::fused).
enum { kNodeSuper = 1, kNodeAnalyse = 2, kNodeRecalc = 3 };
struct NodeInfo {
    int kind;
    void *data;
};
std::mutex g_registry_mutex;
std::map<VSNode *, NodeInfo> g_registry;

// the node createVideoFilter just put into `out` (no reference is kept)
VSNode *registerSelf(const VSAPI *vsapi, VSMap *out, int kind, void *data) {
    VSNode *self = vsapi->mapGetNode(out, "clip", 0, nullptr);
    if (!self)
        return nullptr;
    vsapi->freeNode(self);
    std::lock_guard<std::mutex> lock(g_registry_mutex);
    g_registry[self] = NodeInfo{kind, data};
    return self;
}

void unregisterSelf(VSNode *self) {
    std::lock_guard<std::mutex> lock(g_registry_mutex);
    g_registry.erase(self);
}

// MVB200_LOOKAHEAD: frames per engine call.  Unset = automatic (-1: the window that fills the GPU's search-chain slots once,
// see fusedWindowFrames / analyseWindowFrames); 0 or 1 = off (one frame per call, the reference's own granularity); N = N frames.
int lookaheadFromEnv() {
    const char *la = getenv("MVB200_LOOKAHEAD");
    if (!la || !*la)
        return -1;
    const int v = atoi(la);
    return v < 2 ? 0 : (v > 4096 ? 4096 : v);
}

// Pinned host memory one fused filter instance may hold for its windows (MVB200_PINNED_MB; default: an eighth of the
// machine's memory, between 2 and 32 GiB).
int64_t pinnedBudget() {
    const char *mb = getenv("MVB200_PINNED_MB");
    if (mb && atoll(mb) > 0)
        return atoll(mb) << 20;
    const int64_t phys = (int64_t)sysconf(_SC_PHYS_PAGES) * (int64_t)sysconf(_SC_PAGE_SIZE);
    const int64_t lo = (int64_t)2 << 30, hi = (int64_t)32 << 30;
    const int64_t b = phys / 8;
    return b < lo ? lo : (b > hi ? hi : b);
}

// ---------------------------------------------------------------- fused windows
// Shared by the fused paths of DegrainN / Compensate / Flow* / BlockFPS: when a filter's inputs are this plugin's own
// Super / Analyse nodes wired as in the canonical script, the Super / Analyse nodes are never asked for a frame.  The
// output clip is cut into ALIGNED windows of B frames; a window is produced by ONE engine call over SOURCE frames (a
// *_clip_host entry point: Super, the searches and the consumer all run on the device).  A search is a strict serial chain
// (one warp per frame and vector clip), so B is chosen to fill the GPU's chain slots once (mvb200_search_plan_query), and
// the windows are pipelined: the request that first touches window k also gathers the source frames of window k + 1 and
// enqueues it before it waits for its own frame, so the GPU works on the next window while the frame server's threads
// collect the current one (DegrainN: the engine call is asynchronous; the copies of neighbouring windows overlap the
// kernels).  No request ever waits for another REQUEST to be scheduled -- a request that finds its window not yet
// enqueued holds the window's source frames itself and packs / enqueues it (cooperatively with the other requests in
// the same position) -- so a frame server with few threads cannot deadlock; requests only wait for the GPU.
struct PinnedPool { // window buffers are recycled (process-wide: a script's filters, and a re-created script, share them): page-locking gigabytes per window would cost more than the copies
    std::mutex m;
    std::vector<std::pair<void *, size_t>> free_;
    ~PinnedPool() {
        for (auto &b : free_)
            mvb200_pinned_free(b.first);
    }
    // best fit: a window's source and output buffers differ in size (halo frames; a vector clip's super frames), and a first
    // fit would hand the large ones to the small requests and then page-lock new large ones window after window
    void *get(size_t bytes, size_t *got) {
        {
            std::lock_guard<std::mutex> lock(m);
            size_t best = free_.size(), held = 0;
            for (size_t i = 0; i < free_.size(); i++) {
                held += free_[i].second;
                if (free_[i].second >= bytes && (best == free_.size() || free_[i].second < free_[best].second))
                    best = i;
            }
            if (best != free_.size()) {
                void *p = free_[best].first;
                *got = free_[best].second;
                free_.erase(free_.begin() + (ptrdiff_t)best);
                return p;
            }
            // nothing fits: make room within the budget by giving back the smallest idle buffers
            while (!free_.empty() && (int64_t)(held + bytes) > pinnedBudget()) {
                size_t smallest = 0;
                for (size_t i = 1; i < free_.size(); i++)
                    if (free_[i].second < free_[smallest].second)
                        smallest = i;
                held -= free_[smallest].second;
                mvb200_pinned_free(free_[smallest].first);
                free_.erase(free_.begin() + (ptrdiff_t)smallest);
            }
        }
        *got = bytes;
        return mvb200_pinned_alloc(bytes);
    }
    void put(void *p, size_t bytes) {
        if (!p)
            return;
        std::lock_guard<std::mutex> lock(m);
        size_t held = bytes;
        for (auto &b : free_)
            held += b.second;
        if ((int64_t)held > pinnedBudget()) { // (more idle buffers than any filter may hold: give it back)
            mvb200_pinned_free(p);
            return;
        }
        free_.emplace_back(p, bytes);
    }
};

std::shared_ptr<PinnedPool> sharedPinnedPool() {
    static std::shared_ptr<PinnedPool> pool = std::make_shared<PinnedPool>();
    return pool;
}

struct FusedWindow {
    enum State { Idle = 0, Packing, Launching, Running, Ready, Failed };
    int w0 = 0, w1 = 0, s0 = 0, s1 = 0; // output frames [w0, w1) are made of source frames [s0, s1)
    std::mutex m;
    std::condition_variable cv;
    std::atomic<int> state{Idle};
    bool prefetchClaimed = false; // (under FusedCache::m) some request already gathers this window's source frames ahead of time
    std::atomic<int> packNext{0}, packDone{0};
    std::vector<const VSFrame *> srcFrames; // references to the window's source frames, released one by one as they are packed
    const VSAPI *vsapi = nullptr;
    std::string err;
    void *hostSrc = nullptr, *hostDst = nullptr; // in the clip layout, frame_bytes apart
    size_t srcBytes = 0, dstBytes = 0;
    int slot = 0;               // which engine context computes this window
    std::vector<int32_t> flags; // the engine's "usable" flag per output frame, where the filter has one
    mvb200_event *done = nullptr;
    std::shared_ptr<PinnedPool> pool;
    ~FusedWindow() {
        for (const VSFrame *fr : srcFrames) // (a window that was published but never packed)
            if (fr && vsapi)
                vsapi->freeFrame(fr);
        if (done) {
            mvb200_event_wait(done); // the GPU may still be writing hostDst
            mvb200_event_destroy(done);
        }
        if (pool) {
            pool->put(hostSrc, srcBytes);
            pool->put(hostDst, dstBytes);
        }
    }
};

// What one frame request keeps between its two activations (frameData[0]): its own window (so that it survives the
// cache's eviction while the request waits for frames) and the windows whose source frames this request holds.
struct FusedRequest {
    std::shared_ptr<FusedWindow> win;
    std::vector<std::shared_ptr<FusedWindow>> hold;
};

// MVB200_STATS=1: where the host threads of a fused filter spent their time, printed to stderr when the filter is freed
// (nanoseconds summed over threads; "wait" is time blocked on the GPU, everything else is host work).
struct FusedStats {
    std::atomic<int64_t> request{0}, collect{0}, pack{0}, launch{0}, waitOwn{0}, waitEvent{0}, serve{0}, frames{0}, windows{0};
    const bool on = getenv("MVB200_STATS") != nullptr;
    const int every = on ? atoi(getenv("MVB200_STATS")) : 0; // > 1: also a line every that many windows (what happened since the last one)
    int64_t last[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, lastT = 0;
    std::mutex m;
    static int64_t now() { return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
    void print(const char *name, int window, bool delta) {
        std::lock_guard<std::mutex> lock(m);
        const int64_t cur[9] = {request.load(), collect.load(), pack.load(), launch.load(), waitOwn.load(), waitEvent.load(), serve.load(), frames.load(), windows.load()};
        const int64_t t = now();
        double v[9];
        for (int i = 0; i < 9; i++)
            v[i] = (double)(cur[i] - (delta ? last[i] : 0));
        if (delta)
            fprintf(stderr, "[mvtools-b200] %s: +%.0f frames, +%.0f windows of %d in %.1f ms; ", name, v[7], v[8], window, lastT ? (t - lastT) / 1e6 : 0.0);
        else
            fprintf(stderr, "[mvtools-b200] %s: %.0f frames in %.0f windows of %d; ", name, v[7], v[8], window);
        fprintf(stderr, "host ms (summed over threads): request %.1f, collect %.1f, pack %.1f, launch %.1f, wait for packing / enqueueing %.1f, wait for the GPU %.1f, serve %.1f\n", v[0] / 1e6, v[1] / 1e6,
                v[2] / 1e6, v[3] / 1e6, v[4] / 1e6, v[5] / 1e6, v[6] / 1e6);
        if (delta) {
            memcpy(last, cur, sizeof(last));
            lastT = t;
        }
    }
};
struct StatTimer {
    std::atomic<int64_t> *acc;
    int64_t t0;
    StatTimer(FusedStats &st, std::atomic<int64_t> &a) : acc(st.on ? &a : nullptr), t0(st.on ? FusedStats::now() : 0) {}
    void stop() {
        if (acc)
            acc->fetch_add(FusedStats::now() - t0);
        acc = nullptr;
    }
    ~StatTimer() { stop(); }
};

struct FusedCache {
    FusedStats stats;
    std::atomic<const char *> name{""}; // (every request stores the same literal)
    ~FusedCache() {
        if (stats.on && stats.frames.load())
            stats.print(name.load(), window, false);
    }
    std::mutex m;                                           // guards `windows` and the windows' prefetchClaimed
    std::map<int, std::shared_ptr<FusedWindow>> windows;    // by window index (frame / window)
    std::vector<std::shared_ptr<FusedWindow>> packing;      // windows whose source frames wait to be copied into their pinned buffers
    std::shared_ptr<PinnedPool> pool = sharedPinnedPool();
    mvb200_frame_layout cl;
    // the output frames' layout and bytes per frame in a window's output buffer when they differ from the source clip's (a
    // vector clip's frames are super frames); outExtra more bytes per frame follow the frames (the vector blobs)
    mvb200_frame_layout out{};
    size_t outFrameBytes = 0, outExtra = 0;
    int window = 0;
    int nslots = 1; // engine contexts this filter's windows go round (set with the window: fusedWindowFrames)
    int depth = 1;  // windows enqueued ahead of the one being served: one per context
};

// Frames per window: MVB200_LOOKAHEAD if given, else what fills the search kernel's chain slots once with nvec vector
// clips per frame (at least 16).  Also decides how many engine contexts the filter's windows go round: every device, and as
// many contexts per device as keep the windows' device buffers (source / output clips double buffered, super clip, vectors)
// within 45 % of the device's memory -- and the pinned host windows (one per context, one being served, one recycled)
// within the pinned-memory budget.
int fusedWindowFrames(FusedCache *f, int la, const mvb200_analyse_params *ap, const mvb200_super_params *sp, int nvec) {
    deviceCount();
    const char *mb = getenv("MVB200_PINNED_MB");
    mvb200_window_plan plan;
    if (mvb200_window_plan_query(ap, sp, nvec, la, f->cl.frame_bytes, f->outFrameBytes ? (int64_t)(f->outFrameBytes + f->outExtra) : f->cl.frame_bytes, g_ndevices, g_perdev_asked,
                                 mb && atoll(mb) > 0 ? atoll(mb) << 20 : 0, mvb200_device_memory(g_engines[0].device), &plan)) {
        f->nslots = f->depth = 1;
        return la > 0 ? la : 64;
    }
    f->nslots = g_ndevices * plan.contexts_per_device;
    f->depth = f->nslots;
    return plan.frames;
}

// One activation of a fused filter.  propFrame: the source frame the output frame takes its properties from (always
// requested); segment(w0, w1, &s0, &s1): the source frames the output window [w0, w1) needs; run(ctx, hostSrc, s0, s1, w0,
// w1, hostDst, flags): ENQUEUES the engine call (source frames [s0, s1) in the clip layout -> output frames [w0, w1); flags:
// one "usable" flag per output frame, zeroed) -- it may return before the work is done, completion is an engine event;
// planeMask: bit p set = plane p comes from the engine, else it is a copy of the source frame's.  alsoFrame (-1: none):
// one more source frame the caller reads after this returns (requested here so that no frame is requested twice);
// flagOut: the frame's "usable" flag.
// post (optional): called with the finished output frame and the window's output buffer, to attach what is not pixels.
template <class Segment, class Run>
const VSFrame *fusedGetFrame(const char *filter, FusedCache *f, VSNode *clip, const VSVideoInfo *outvi, int n, int propFrame, unsigned planeMask, int activationReason, void **frameData, VSFrameContext *frameCtx,
                             VSCore *core, const VSAPI *vsapi, Segment segment, Run run, int alsoFrame = -1, int32_t *flagOut = nullptr,
                             const std::function<void(VSFrame *, const uint8_t *hostDst, int w0, int w1)> &post = nullptr) {
    const int B = f->window, N = outvi->numFrames;
    if (activationReason == arInitial) {
        StatTimer tRequest(f->stats, f->stats.request);
        f->name.store(filter, std::memory_order_relaxed);
        vsapi->requestFrameFilter(propFrame, clip, frameCtx);
        if (alsoFrame >= 0 && alsoFrame != propFrame)
            vsapi->requestFrameFilter(alsoFrame, clip, frameCtx);
        FusedRequest *rq = new FusedRequest;
        {
            std::lock_guard<std::mutex> lock(f->m);
            const int k = n / B;
            auto get = [&](int idx) -> std::shared_ptr<FusedWindow> & {
                std::shared_ptr<FusedWindow> &w = f->windows[idx];
                if (!w) {
                    w = std::make_shared<FusedWindow>();
                    w->pool = f->pool;
                    w->slot = idx % f->nslots;
                    w->w0 = idx * B;
                    w->w1 = w->w0 + B < N ? w->w0 + B : N;
                    segment(w->w0, w->w1, &w->s0, &w->s1);
                }
                return w;
            };
            rq->win = get(k);
            if (rq->win->state.load() < FusedWindow::Running) // not enqueued yet: this request may have to do it
                rq->hold.push_back(rq->win);
            for (int a = 1; a <= f->depth && (int64_t)(k + a) * B < N; a++) { // ... and the windows after it, once
                std::shared_ptr<FusedWindow> &w = get(k + a);
                if (w->state.load() == FusedWindow::Idle && !w->prefetchClaimed) {
                    w->prefetchClaimed = true;
                    rq->hold.push_back(w);
                }
            }
            // keep the window being served, the ones ahead of it and one behind; requests in flight keep their own alive
            for (auto it = f->windows.begin(); it != f->windows.end();)
                it = (it->first < k - 1 || it->first > k + f->depth + 1) ? f->windows.erase(it) : std::next(it);
        }
        int requestedEnd = -1; // the windows are in ascending order: their segments overlap by the halo at most
        for (auto &w : rq->hold)
            for (int s = w->s0 > requestedEnd ? w->s0 : requestedEnd; s < w->s1; s++) {
                if (s != propFrame && s != alsoFrame)
                    vsapi->requestFrameFilter(s, clip, frameCtx);
                requestedEnd = s + 1;
            }
        frameData[0] = rq;
        return nullptr;
    }
    std::unique_ptr<FusedRequest> rq((FusedRequest *)frameData[0]); // (released on every exit, arError included)
    frameData[0] = nullptr;
    if (!rq)
        return nullptr;
    if (activationReason != arAllFramesReady) {
        std::lock_guard<std::mutex> lock(f->m); // a failed request's look-ahead claims go back
        for (auto &w : rq->hold)
            if (w != rq->win && w->state.load() == FusedWindow::Idle)
                w->prefetchClaimed = false;
        return nullptr;
    }
    const size_t fb = (size_t)f->cl.frame_bytes;
    // (1) every window this request holds the source frames of and nobody has taken yet: take references to its frames and
    // publish it for packing (the frames then belong to the window, not to this request)
    for (auto &w : rq->hold) {
        std::unique_lock<std::mutex> wl(w->m);
        if (w->state.load() != FusedWindow::Idle)
            continue;
        StatTimer tCollect(f->stats, f->stats.collect);
        size_t gs = 0, gd = 0;
        w->hostSrc = w->pool->get(fb * (size_t)(w->s1 - w->s0), &gs);
        w->srcBytes = gs;
        w->hostDst = w->hostSrc ? w->pool->get(((f->outFrameBytes ? f->outFrameBytes : fb) + f->outExtra) * (size_t)(w->w1 - w->w0), &gd) : nullptr;
        w->dstBytes = gd;
        w->flags.assign((size_t)(w->w1 - w->w0), 0);
        if (!w->hostSrc || !w->hostDst) {
            w->err = mvb200_last_error();
            w->state.store(FusedWindow::Failed);
            w->cv.notify_all();
            continue;
        }
        const int count = w->s1 - w->s0;
        w->vsapi = vsapi;
        w->srcFrames.resize((size_t)count);
        for (int i = 0; i < count; i++)
            w->srcFrames[(size_t)i] = vsapi->getFrameFilter(w->s0 + i, clip, frameCtx);
        w->state.store(FusedWindow::Packing);
        wl.unlock();
        std::lock_guard<std::mutex> lock(f->m);
        f->packing.push_back(w);
    }
    // (2) pack: EVERY request of this filter copies frames of whatever windows wait for it into their pinned buffers (a window
    // is a gigabyte: one thread would need longer for it than the GPU for the window before); the thread that packs a window's
    // last frame enqueues it on its engine context
    auto packSome = [&]() -> bool {
        std::shared_ptr<FusedWindow> w;
        {
            std::lock_guard<std::mutex> lock(f->m);
            for (auto it = f->packing.begin(); it != f->packing.end();) {
                if ((*it)->packNext.load() >= (*it)->s1 - (*it)->s0) {
                    it = f->packing.erase(it); // every frame is taken
                } else {
                    w = *it;
                    break;
                }
            }
        }
        if (!w)
            return false;
        const int count = w->s1 - w->s0;
        const int i = w->packNext.fetch_add(1);
        if (i >= count)
            return true;
        StatTimer tPack(f->stats, f->stats.pack);
        const VSFrame *fr = w->srcFrames[(size_t)i];
        for (int p = 0; p < f->cl.nplanes; p++)
            vsh::bitblt((uint8_t *)w->hostSrc + fb * (size_t)i + f->cl.offset[p], (ptrdiff_t)f->cl.pitch[p], vsapi->getReadPtr(fr, p), vsapi->getStride(fr, p),
                        (size_t)f->cl.plane_w[p] * f->cl.bytes_per_sample, (size_t)f->cl.plane_h[p]);
        w->srcFrames[(size_t)i] = nullptr;
        vsapi->freeFrame(fr);
        if (w->packDone.fetch_add(1) + 1 < count)
            return true;
        w->state.store(FusedWindow::Launching);
        tPack.stop();
        StatTimer tLaunch(f->stats, f->stats.launch);
        if (f->stats.on && (f->stats.windows.fetch_add(1) + 1) % (f->stats.every > 1 ? f->stats.every : (1 << 30)) == 0)
            f->stats.print(filter, f->window, true);
        std::string msg;
        bool ok = false;
        mvb200_event *ev = nullptr;
        {
            EngineSlot &es = g_engines[w->slot];
            std::lock_guard<std::mutex> elock(es.m);
            mvb200_ctx *ctx = engineOf(es, msg);
            if (ctx) {
                ok = run(ctx, w->hostSrc, w->s0, w->s1, w->w0, w->w1, w->hostDst, w->flags.data()) == 0 && mvb200_ctx_record(ctx, &ev) == 0;
                if (!ok)
                    msg = mvb200_last_error();
            }
        }
        std::lock_guard<std::mutex> wl(w->m);
        w->done = ev;
        w->err = msg;
        w->state.store(ok ? FusedWindow::Running : FusedWindow::Failed);
        w->cv.notify_all();
        return true;
    };
    while (packSome()) {
    }
    // (3) this request's own window may still be in other threads' hands (its last frames being packed, or being enqueued)
    {
        StatTimer tWaitOwn(f->stats, f->stats.waitOwn);
        FusedWindow *own = rq->win.get();
        std::unique_lock<std::mutex> wl(own->m);
        while (own->state.load() == FusedWindow::Packing || own->state.load() == FusedWindow::Launching) {
            wl.unlock();
            const bool worked = packSome();
            wl.lock();
            if (!worked && (own->state.load() == FusedWindow::Packing || own->state.load() == FusedWindow::Launching))
                own->cv.wait_for(wl, std::chrono::milliseconds(1));
        }
    }
    const VSFrame *src = vsapi->getFrameFilter(propFrame, clip, frameCtx);
    FusedWindow *w = rq->win.get();
    std::string msg;
    int st = w->state.load();
    if (st == FusedWindow::Running) {
        StatTimer tWaitEvent(f->stats, f->stats.waitEvent);
        if (mvb200_event_wait(w->done)) { // (any number of threads may wait for the same event)
            std::lock_guard<std::mutex> wl(w->m);
            w->err = mvb200_last_error();
            w->state.store(FusedWindow::Failed);
        } else {
            int expect = FusedWindow::Running;
            w->state.compare_exchange_strong(expect, FusedWindow::Ready);
        }
        st = w->state.load();
    }
    if (st != FusedWindow::Ready) {
        {
            std::lock_guard<std::mutex> wl(w->m);
            msg = st == FusedWindow::Failed ? w->err : "internal: window neither enqueued nor held";
        }
        vsapi->setFilterError((std::string(filter) + ": " + msg).c_str(), frameCtx);
        vsapi->freeFrame(src);
        return nullptr;
    }
    StatTimer tServe(f->stats, f->stats.serve);
    f->stats.frames.fetch_add(1);
    if (flagOut)
        *flagOut = w->flags[(size_t)(n - w->w0)];
    const mvb200_frame_layout &ol = f->outFrameBytes ? f->out : f->cl;
    const uint8_t *bytes = (const uint8_t *)w->hostDst + (f->outFrameBytes ? f->outFrameBytes : fb) * (size_t)(n - w->w0);
    VSFrame *dst = vsapi->newVideoFrame(&outvi->format, outvi->width, outvi->height, src, core);
    for (int p = 0; p < outvi->format.numPlanes; p++) {
        const size_t row = (size_t)ol.plane_w[p] * ol.bytes_per_sample, rows = (size_t)ol.plane_h[p];
        if (planeMask & (1u << p))
            vsh::bitblt(vsapi->getWritePtr(dst, p), vsapi->getStride(dst, p), bytes + ol.offset[p], (ptrdiff_t)ol.pitch[p], row, rows);
        else
            vsh::bitblt(vsapi->getWritePtr(dst, p), vsapi->getStride(dst, p), vsapi->getReadPtr(src, p), vsapi->getStride(src, p), row, rows);
    }
    if (post)
        post(dst, (const uint8_t *)w->hostDst, w->w0, w->w1);
    vsapi->freeFrame(src);
    return dst;
}

int argInt(const VSAPI *vsapi, const VSMap *in, const char *name) {
    int err = 0;
    int v = vsapi->mapGetIntSaturated(in, name, 0, &err);
    return err ? MVB200_UNSET : v;
}

int64_t argInt64(const VSAPI *vsapi, const VSMap *in, const char *name) {
    int err = 0;
    int64_t v = vsapi->mapGetInt(in, name, 0, &err);
    return err ? MVB200_UNSET64 : v;
}

// The MVTools_vectors blob of a vector frame (GroupOfPlanes.c:76-85: its first int is its own size in bytes).  *err is set when
// the property is missing or when it is shorter than it claims to be -- the engine trusts that size.
const uint8_t *vectorBlob(const VSAPI *vsapi, const VSFrame *f, int *err) {
    int e = 0;
    const VSMap *props = vsapi->getFramePropertiesRO(f);
    const uint8_t *v = (const uint8_t *)vsapi->mapGetData(props, "MVTools_vectors", 0, &e);
    if (!e) {
        const int size = vsapi->mapGetDataSize(props, "MVTools_vectors", 0, nullptr);
        int32_t claimed = 0;
        if (size >= 4)
            memcpy(&claimed, v, 4);
        if (size < 8 || claimed > size)
            e = 1;
    }
    if (e)
        *err = e;
    return e ? nullptr : v;
}

bool formatOk(const VSVideoInfo *vi) {
    return vsh::isConstantVideoFormat(vi) && vi->format.bitsPerSample <= 16 && vi->format.sampleType == stInteger && vi->format.subSamplingW <= 1 &&
           vi->format.subSamplingH <= 1 && (vi->format.colorFamily == cfYUV || vi->format.colorFamily == cfGray);
}

void planesRO(const VSAPI *vsapi, const VSFrame *f, int nplanes, const uint8_t *p[3], ptrdiff_t s[3]) {
    for (int i = 0; i < 3; i++) {
        p[i] = i < nplanes && f ? vsapi->getReadPtr(f, i) : nullptr;
        s[i] = i < nplanes && f ? vsapi->getStride(f, i) : 0;
    }
}

void planesRW(const VSAPI *vsapi, VSFrame *f, int nplanes, uint8_t *p[3], ptrdiff_t s[3]) {
    for (int i = 0; i < 3; i++) {
        p[i] = i < nplanes ? vsapi->getWritePtr(f, i) : nullptr;
        s[i] = i < nplanes ? vsapi->getStride(f, i) : 0;
    }
}

// Reads the Super_* properties of frame 0 of a super clip (MVAnalyse.c:523-538) into a super-params struct.
bool superFromClip(VSNode *super, const VSAPI *vsapi, const char *filter, mvb200_super_params *sp, std::string &err) {
    char msg[1024];
    const VSFrame *evil = vsapi->getFrame(0, super, msg, sizeof(msg));
    if (!evil) {
        err = std::string(filter) + ": failed to retrieve first frame from super clip. Error message: " + msg;
        return false;
    }
    const VSMap *props = vsapi->getFramePropertiesRO(evil);
    int e[6];
    int height = vsapi->mapGetIntSaturated(props, "Super_height", 0, &e[0]);
    int hpad = vsapi->mapGetIntSaturated(props, "Super_hpad", 0, &e[1]);
    int vpad = vsapi->mapGetIntSaturated(props, "Super_vpad", 0, &e[2]);
    int pel = vsapi->mapGetIntSaturated(props, "Super_pel", 0, &e[3]);
    int modeyuv = vsapi->mapGetIntSaturated(props, "Super_modeyuv", 0, &e[4]);
    int levels = vsapi->mapGetIntSaturated(props, "Super_levels", 0, &e[5]);
    vsapi->freeFrame(evil);
    for (int i = 0; i < 6; i++)
        if (e[i]) {
            err = std::string(filter) + ": required properties not found in first frame of super clip. Maybe clip didn't come from mv.Super? Was the first frame trimmed away?";
            return false;
        }
    const VSVideoInfo *vi = vsapi->getVideoInfo(super);
    if (height <= 0 || hpad < 0 || hpad >= vi->width / 2 || vpad < 0 || pel < 1 || pel > 4 || modeyuv < 0 || modeyuv > MVB200_YUVPLANES || levels < 1) {
        err = std::string(filter) + ": parameters from super clip appear to be wrong.";
        return false;
    }
    memset(sp, 0, sizeof(*sp));
    sp->width = vi->width - 2 * hpad;
    sp->height = height;
    sp->bits = vi->format.bitsPerSample;
    sp->gray = vi->format.colorFamily == cfGray;
    sp->ssw = vi->format.subSamplingW;
    sp->ssh = vi->format.subSamplingH;
    sp->hpad = hpad;
    sp->vpad = vpad;
    sp->pel = pel;
    sp->levels = levels;
    sp->chroma = (modeyuv & (MVB200_UPLANE | MVB200_VPLANE)) ? 1 : 0;
    sp->sharp = MVB200_UNSET;
    sp->rfilter = MVB200_UNSET;
    if (mvb200_super_init(sp) || sp->levels != levels) {
        err = std::string(filter) + ": parameters from super clip appear to be wrong.";
        return false;
    }
    return true;
}

// ------------------------------------------------------------------ Super
struct SuperData {
    VSNode *node, *pelclip;
    VSVideoInfo vi;
    mvb200_super_params sp;
    int usePelClip, pelPadded;
    VSNode *self; // this filter's output node (registry key), not a reference
};

// the properties mv.Super attaches to frame 0 of a super clip (MVSuper.c:111-120)
void setSuperProps(const VSAPI *vsapi, VSFrame *dst, const mvb200_super_params &sp) {
    VSMap *props = vsapi->getFramePropertiesRW(dst);
    vsapi->mapSetInt(props, "Super_height", sp.height, maReplace);
    vsapi->mapSetInt(props, "Super_hpad", sp.hpad, maReplace);
    vsapi->mapSetInt(props, "Super_vpad", sp.vpad, maReplace);
    vsapi->mapSetInt(props, "Super_pel", sp.pel, maReplace);
    vsapi->mapSetInt(props, "Super_modeyuv", sp.modeyuv, maReplace);
    vsapi->mapSetInt(props, "Super_levels", sp.levels, maReplace);
}

const VSFrame *VS_CC superGetFrame(int n, int activationReason, void *instanceData, void **, VSFrameContext *frameCtx, VSCore *core, const VSAPI *vsapi) {
    SuperData *d = (SuperData *)instanceData;
    if (activationReason == arInitial) {
        vsapi->requestFrameFilter(n, d->node, frameCtx);
        if (d->usePelClip)
            vsapi->requestFrameFilter(n, d->pelclip, frameCtx);
    } else if (activationReason == arAllFramesReady) {
        const VSFrame *src = vsapi->getFrameFilter(n, d->node, frameCtx);
        const VSFrame *srcPel = d->usePelClip ? vsapi->getFrameFilter(n, d->pelclip, frameCtx) : nullptr;
        VSFrame *dst = vsapi->newVideoFrame(&d->vi.format, d->vi.width, d->vi.height, src, core);
        const uint8_t *sp_[3], *pp_[3];
        ptrdiff_t ss[3], ps[3];
        uint8_t *dp[3];
        ptrdiff_t ds[3];
        planesRO(vsapi, src, d->vi.format.numPlanes, sp_, ss);
        planesRO(vsapi, srcPel, d->vi.format.numPlanes, pp_, ps);
        planesRW(vsapi, dst, d->vi.format.numPlanes, dp, ds);
        std::string err;
        int rc;
        {
            std::lock_guard<std::mutex> lock(g_engine_mutex);
            mvb200_ctx *ctx = engine(err);
            if (!ctx)
                rc = 1;
            else if (srcPel)
                rc = mvb200_super_frame_pel(ctx, &d->sp, sp_, ss, pp_, ps, d->pelPadded, dp, ds);
            else
                rc = mvb200_super_frame(ctx, &d->sp, sp_, ss, dp, ds);
            if (ctx && rc)
                err = mvb200_last_error();
        }
        vsapi->freeFrame(src);
        vsapi->freeFrame(srcPel);
        if (rc) {
            vsapi->setFilterError(("Super: " + err).c_str(), frameCtx);
            vsapi->freeFrame(dst);
            return nullptr;
        }
        if (n == 0) // MVSuper.c:111-120
            setSuperProps(vsapi, dst, d->sp);
        return dst;
    }
    return nullptr;
}

void VS_CC superFree(void *instanceData, VSCore *, const VSAPI *vsapi) {
    SuperData *d = (SuperData *)instanceData;
    unregisterSelf(d->self);
    vsapi->freeNode(d->node);
    vsapi->freeNode(d->pelclip);
    delete d;
}

void VS_CC superCreate(const VSMap *in, VSMap *out, void *, VSCore *core, const VSAPI *vsapi) {
    SuperData d;
    memset(&d.sp, 0, sizeof(d.sp));
    d.pelclip = nullptr;
    d.usePelClip = d.pelPadded = 0;
    d.sp.hpad = argInt(vsapi, in, "hpad");
    d.sp.vpad = argInt(vsapi, in, "vpad");
    d.sp.pel = argInt(vsapi, in, "pel");
    d.sp.levels = argInt(vsapi, in, "levels");
    d.sp.chroma = argInt(vsapi, in, "chroma");
    d.sp.sharp = argInt(vsapi, in, "sharp");
    d.sp.rfilter = argInt(vsapi, in, "rfilter");
    d.node = vsapi->mapGetNode(in, "clip", 0, nullptr);
    d.vi = *vsapi->getVideoInfo(d.node);
    if (!formatOk(&d.vi)) {
        vsapi->mapSetError(out, "Super: input clip must be GRAY, 420, 422, 440, or 444, up to 16 bits, with constant dimensions.");
        vsapi->freeNode(d.node);
        return;
    }
    d.sp.width = d.vi.width;
    d.sp.height = d.vi.height;
    d.sp.bits = d.vi.format.bitsPerSample;
    d.sp.gray = d.vi.format.colorFamily == cfGray;
    d.sp.ssw = d.vi.format.subSamplingW;
    d.sp.ssh = d.vi.format.subSamplingH;
    if (mvb200_super_init(&d.sp)) {
        vsapi->mapSetError(out, mvb200_last_error());
        vsapi->freeNode(d.node);
        return;
    }
    int err = 0;
    d.pelclip = vsapi->mapGetNode(in, "pelclip", 0, &err); // MVSuper.c:229-253
    if (d.pelclip) {
        const VSVideoInfo *pelvi = vsapi->getVideoInfo(d.pelclip);
        const bool same = vsh::isConstantVideoFormat(pelvi) && pelvi->format.colorFamily == d.vi.format.colorFamily && pelvi->format.sampleType == d.vi.format.sampleType &&
                          pelvi->format.bitsPerSample == d.vi.format.bitsPerSample && pelvi->format.subSamplingW == d.vi.format.subSamplingW &&
                          pelvi->format.subSamplingH == d.vi.format.subSamplingH;
        if (mvb200_super_pelclip_check(&d.sp, same, pelvi->width, pelvi->height, &d.usePelClip, &d.pelPadded)) {
            vsapi->mapSetError(out, mvb200_last_error());
            vsapi->freeNode(d.node);
            vsapi->freeNode(d.pelclip);
            return;
        }
    }
    d.vi.width = d.sp.super_width;
    d.vi.height = d.sp.super_height;
    SuperData *data = new SuperData(d);
    VSFilterDependency deps[1] = {{data->node, rpStrictSpatial}}; // (the reference declares the source clip only, MVSuper.c:270-272)
    vsapi->createVideoFilter(out, "Super", &data->vi, superGetFrame, superFree, fmParallel, deps, 1, data, core);
    data->self = registerSelf(vsapi, out, kNodeSuper, data);
}

// ---------------------------------------------------------------- Analyse
// Look-ahead (MVB200_LOOKAHEAD=<frames>, default 0 = off): a search is a strict serial chain, so ONE frame uses one warp of
// the GPU.  With a window of B frames the filter requests the super frames of [n, n + B) (and their reference frames) on
// a cache miss, searches all B chains in ONE engine call (mvb200_analyse_frames) and serves the following requests
// from the cache: the frame server's usual in-order requests then cost one launch per B frames instead of one per frame.
struct AnalyseCache {
    std::mutex m;                                   // serialises window computations of this instance
    std::map<int, std::vector<uint8_t>> blobs;      // frame -> MVTools_vectors
    std::deque<int> order;                          // insertion order, for eviction
    void put(int n, std::vector<uint8_t> &&v, size_t keep) {
        if (blobs.emplace(n, std::move(v)).second)
            order.push_back(n);
        while (order.size() > keep) {
            blobs.erase(order.front());
            order.pop_front();
        }
    }
};

struct AnalyseData {
    VSNode *node;
    const VSVideoInfo *vi;
    mvb200_super_params sp;
    mvb200_analyse_params ap;
    int lookahead;       // window length, 0 = one frame per call
    AnalyseCache *cache; // owned; null without look-ahead
    VSNode *self;        // this filter's output node (registry key), not a reference
    // fused path (see analyseFusedGetFrame): the super clip is this plugin's own mv.Super of `clip`
    FusedCache *fused;   // owned; null = the super clip is read frame by frame
    VSNode *clip;        // the Super's source clip (a reference of our own)
    mvb200_super_params ssp; // the Super node's own parameters (sharp / rfilter are not in the frame properties)
    mvb200_frame_layout sl;
    int64_t vstride;
};

int refFrameOf(const mvb200_analyse_params &ap, int n) { // MVAnalyse.c:84-109
    if (ap.ad.nDeltaFrame > 0)
        return n + (ap.ad.isBackward ? ap.ad.nDeltaFrame : -ap.ad.nDeltaFrame);
    return -ap.ad.nDeltaFrame;
}

// MVAnalyse.c:136-179: the vertical shift between fields of opposite parity; false = the _Field property is needed and missing
bool analyseFieldShift(const AnalyseData *d, const VSAPI *vsapi, const VSFrame *src, const VSFrame *ref, int n, int nref, int *fieldShift) {
    int err = 0;
    *fieldShift = 0;
    int srcTop = !!vsapi->mapGetInt(vsapi->getFramePropertiesRO(src), "_Field", 0, &err);
    bool missing = err && d->ap.fields && !d->ap.tff_exists;
    int refTop = 0;
    if (ref) {
        refTop = !!vsapi->mapGetInt(vsapi->getFramePropertiesRO(ref), "_Field", 0, &err);
        missing = missing || (err && d->ap.fields && !d->ap.tff_exists);
    }
    if (missing)
        return false;
    if (d->ap.tff_exists) {
        srcTop = d->ap.tff ^ (n % 2);
        refTop = d->ap.tff ^ (nref % 2);
    }
    if (ref && d->ap.fields && d->ap.ad.nPel > 1 && (d->ap.ad.nDeltaFrame % 2))
        *fieldShift = (srcTop && !refTop) ? d->ap.ad.nPel / 2 : ((refTop && !srcTop) ? -(d->ap.ad.nPel / 2) : 0);
    return true;
}

// searches the window [n, end) in one engine call and fills the cache; returns false with msg set on failure
bool analyseWindow(AnalyseData *d, const VSAPI *vsapi, VSFrameContext *frameCtx, int n, int end, std::string &msg) {
    std::map<int, const VSFrame *> held; // distinct super frames of the window, by frame number
    auto hold = [&](int k) {
        auto it = held.find(k);
        if (it == held.end())
            it = held.emplace(k, vsapi->getFrameFilter(k, d->node, frameCtx)).first;
        return it->second;
    };
    std::vector<int> jobFrame;
    std::vector<int32_t> srcIdx, refIdx, shifts;
    bool ok = true;
    for (int k = n; k < end && ok; k++) {
        if (d->cache->blobs.count(k))
            continue;
        const int nref = refFrameOf(d->ap, k);
        const bool refOk = nref >= 0 && nref < d->vi->numFrames;
        const VSFrame *src = hold(k), *ref = refOk ? hold(nref) : nullptr;
        int fs = 0;
        if (!analyseFieldShift(d, vsapi, src, ref, k, nref, &fs)) {
            msg = "_Field property not found in input frame. Therefore, you must pass tff argument.";
            ok = false;
            break;
        }
        jobFrame.push_back(k);
        srcIdx.push_back(k);
        refIdx.push_back(refOk ? nref : -1);
        shifts.push_back(fs);
    }
    if (ok && !jobFrame.empty()) {
        std::map<int, int> slot; // frame number -> index in the upload list
        std::vector<const uint8_t *> planes;
        std::vector<ptrdiff_t> strides;
        for (auto &h : held) {
            slot[h.first] = (int)slot.size();
            const uint8_t *p[3];
            ptrdiff_t st[3];
            planesRO(vsapi, h.second, d->vi->format.numPlanes, p, st);
            planes.insert(planes.end(), p, p + 3);
            strides.insert(strides.end(), st, st + 3);
        }
        for (size_t j = 0; j < jobFrame.size(); j++) {
            srcIdx[j] = slot[srcIdx[j]];
            refIdx[j] = refIdx[j] >= 0 ? slot[refIdx[j]] : -1;
        }
        std::vector<std::vector<uint8_t>> out(jobFrame.size(), std::vector<uint8_t>((size_t)d->ap.vectors_size));
        std::vector<uint8_t *> outp;
        for (auto &v : out)
            outp.push_back(v.data());
        int rc;
        {
            std::lock_guard<std::mutex> lock(g_engine_mutex);
            mvb200_ctx *ctx = engine(msg);
            rc = ctx ? mvb200_analyse_frames(ctx, &d->ap, &d->sp, (int)slot.size(), planes.data(), strides.data(), (int)jobFrame.size(), srcIdx.data(), refIdx.data(), shifts.data(), outp.data()) : 1;
            if (ctx && rc)
                msg = mvb200_last_error();
        }
        ok = rc == 0;
        for (size_t j = 0; ok && j < jobFrame.size(); j++)
            d->cache->put(jobFrame[j], std::move(out[j]), (size_t)d->lookahead * 4);
    }
    for (auto &h : held)
        vsapi->freeFrame(h.second);
    return ok;
}

const VSFrame *VS_CC analyseGetFrame(int n, int activationReason, void *instanceData, void **frameData, VSFrameContext *frameCtx, VSCore *core, const VSAPI *vsapi) {
    AnalyseData *d = (AnalyseData *)instanceData;
    const int nref = refFrameOf(d->ap, n);
    const bool refOk = nref >= 0 && nref < d->vi->numFrames;
    if (activationReason == arInitial) {
        // always what the reference requests (MVAnalyse.c:84-109): the single-frame path stays possible
        if (refOk) {
            vsapi->requestFrameFilter(n < nref ? n : nref, d->node, frameCtx);
            vsapi->requestFrameFilter(n < nref ? nref : n, d->node, frameCtx);
        } else {
            vsapi->requestFrameFilter(n, d->node, frameCtx);
        }
        if (d->lookahead > 1) {
            bool hit;
            {
                std::lock_guard<std::mutex> lock(d->cache->m);
                hit = d->cache->blobs.count(n) != 0;
            }
            if (!hit) { // the rest of the window and its reference frames
                const int end = n + d->lookahead < d->vi->numFrames ? n + d->lookahead : d->vi->numFrames;
                for (int k = n + 1; k < end; k++) {
                    vsapi->requestFrameFilter(k, d->node, frameCtx);
                    const int r = refFrameOf(d->ap, k);
                    if (r >= 0 && r < d->vi->numFrames)
                        vsapi->requestFrameFilter(r, d->node, frameCtx);
                }
                *frameData = (void *)(intptr_t)end; // remembered for arAllFramesReady: the window was requested
            }
        }
    } else if (activationReason == arAllFramesReady) {
        const VSFrame *src = vsapi->getFrameFilter(n, d->node, frameCtx);
        std::vector<uint8_t> vectors;
        std::string msg;
        bool have = false;
        if (d->lookahead > 1) {
            std::lock_guard<std::mutex> lock(d->cache->m);
            auto it = d->cache->blobs.find(n);
            const int end = (int)(intptr_t)*frameData;
            if (it == d->cache->blobs.end() && end > n) { // miss, and this request holds the window's frames
                if (!analyseWindow(d, vsapi, frameCtx, n, end, msg)) {
                    vsapi->setFilterError(("Analyse: " + msg).c_str(), frameCtx);
                    vsapi->freeFrame(src);
                    return nullptr;
                }
                it = d->cache->blobs.find(n);
            }
            if (it != d->cache->blobs.end()) {
                vectors = it->second;
                have = true;
            }
        }
        if (!have) { // one frame, one chain (also the fall-back when a cached blob was evicted in between)
            const VSFrame *ref = refOk ? vsapi->getFrameFilter(nref, d->node, frameCtx) : nullptr;
            int fieldShift = 0;
            if (!analyseFieldShift(d, vsapi, src, ref, n, nref, &fieldShift)) {
                vsapi->setFilterError("Analyse: _Field property not found in input frame. Therefore, you must pass tff argument.", frameCtx);
                vsapi->freeFrame(src);
                vsapi->freeFrame(ref);
                return nullptr;
            }
            vectors.resize((size_t)d->ap.vectors_size);
            const uint8_t *sp_[3], *rp[3];
            ptrdiff_t ss[3], rs[3];
            planesRO(vsapi, src, d->vi->format.numPlanes, sp_, ss);
            planesRO(vsapi, ref, d->vi->format.numPlanes, rp, rs);
            int rc;
            {
                std::lock_guard<std::mutex> lock(g_engine_mutex);
                mvb200_ctx *ctx = engine(msg);
                rc = ctx ? mvb200_analyse_frame(ctx, &d->ap, &d->sp, sp_, ss, ref ? rp : nullptr, ref ? rs : nullptr, fieldShift, vectors.data()) : 1;
                if (ctx && rc)
                    msg = mvb200_last_error();
            }
            vsapi->freeFrame(ref);
            if (rc) {
                vsapi->setFilterError(("Analyse: " + msg).c_str(), frameCtx);
                vsapi->freeFrame(src);
                return nullptr;
            }
        }
        // the output frame is a copy of the super frame carrying the two properties (MVAnalyse.c:224-239)
        VSFrame *dst = vsapi->copyFrame(src, core);
        VSMap *props = vsapi->getFramePropertiesRW(dst);
        const mvb200_analysis_data *ad = d->ap.divide ? &d->ap.ad_divided : &d->ap.ad;
        vsapi->mapSetData(props, "MVTools_MVAnalysisData", (const char *)ad, sizeof(*ad), dtBinary, maReplace);
        vsapi->mapSetData(props, "MVTools_vectors", (const char *)vectors.data(), (int)vectors.size(), dtBinary, maReplace);
        vsapi->freeFrame(src);
        return dst;
    }
    return nullptr;
}

// Fused path of a vector clip that something else than this plugin's fused consumers reads (another plugin, mv.Mask,
// mv.SCDetection, a script that outputs the vectors): when the super clip is this plugin's own mv.Super, a window of vector frames
// is ONE mvb200_analyse_clip_host call over SOURCE frames -- Super and the searches on the device, the super frames (the pixels an
// mv.Analyse frame carries, MVAnalyse.c:224) and the blobs back -- instead of the Super node's frames one by one, down and up again.
const VSFrame *VS_CC analyseGetFrame(int n, int activationReason, void *instanceData, void **frameData, VSFrameContext *frameCtx, VSCore *core, const VSAPI *vsapi);

const VSFrame *VS_CC analyseFusedGetFrame(int n, int activationReason, void *instanceData, void **frameData, VSFrameContext *frameCtx, VSCore *core, const VSAPI *vsapi) {
    AnalyseData *d = (AnalyseData *)instanceData;
    FusedCache *f = d->fused;
    // Frame 0 is what every consumer reads when the script is BUILT (adataFromVectorClip, MVAnalysisData.c:42-63): while no window
    // holds it, it is searched on its own from the super clip (one chain) instead of starting a whole window nobody may ever read
    // (the fused consumers never come back for frames).  frameData[1] remembers the choice between the two activations.
    if (activationReason == arInitial) {
        bool single = false;
        if (n == 0) {
            std::lock_guard<std::mutex> lock(f->m);
            const auto it = f->windows.find(0);
            single = it == f->windows.end() || it->second->state.load() == FusedWindow::Idle;
        }
        frameData[1] = single ? (void *)1 : nullptr;
    }
    if (frameData[1])
        return analyseGetFrame(n, activationReason, instanceData, frameData, frameCtx, core, vsapi);
    const int N = d->vi->numFrames, delta = d->ap.ad.nDeltaFrame;
    auto segment = [&](int w0, int w1, int *s0, int *s1) { // the window and the frames its vectors refer to
        *s0 = d->ap.ad.isBackward ? w0 : (w0 - delta > 0 ? w0 - delta : 0);
        *s1 = d->ap.ad.isBackward ? (w1 + delta < N ? w1 + delta : N) : w1;
    };
    auto run = [&](mvb200_ctx *ctx, const void *hostSrc, int s0, int s1, int w0, int w1, void *hostDst, int32_t *) {
        uint8_t *blobs = (uint8_t *)hostDst + (size_t)d->sl.frame_bytes * (size_t)(w1 - w0);
        return mvb200_analyse_clip_host_async(ctx, &d->ssp, &d->ap, hostSrc, s0, s1 - s0, w0, w1 - w0, hostDst, blobs, d->vstride);
    };
    const unsigned mask = (d->ssp.modeyuv & (MVB200_UPLANE | MVB200_VPLANE)) ? 7u : 1u;
    auto post = [&](VSFrame *dst, const uint8_t *hostDst, int w0, int w1) {
        if (n == 0) // the frame is a copy of super frame n (MVAnalyse.c:224), which carries Super_* on frame 0 (MVSuper.c:111-120)
            setSuperProps(vsapi, dst, d->ssp);
        if (!(mask & 2u)) // a luma-only super clip: the reference's super frames keep their zeroed chroma planes (MVSuper.c:75)
            for (int p = 1; p < d->vi->format.numPlanes; p++)
                memset(vsapi->getWritePtr(dst, p), 0, (size_t)vsapi->getStride(dst, p) * (size_t)vsapi->getFrameHeight(dst, p));
        VSMap *props = vsapi->getFramePropertiesRW(dst);
        const mvb200_analysis_data *ad = d->ap.divide ? &d->ap.ad_divided : &d->ap.ad;
        const uint8_t *blob = hostDst + (size_t)d->sl.frame_bytes * (size_t)(w1 - w0) + (size_t)d->vstride * (size_t)(n - w0);
        vsapi->mapSetData(props, "MVTools_MVAnalysisData", (const char *)ad, sizeof(*ad), dtBinary, maReplace);
        vsapi->mapSetData(props, "MVTools_vectors", (const char *)blob, d->ap.vectors_size, dtBinary, maReplace);
    };
    return fusedGetFrame("Analyse", f, d->clip, d->vi, n, n, 7u, activationReason, frameData, frameCtx, core, vsapi, segment, run, -1, nullptr, post);
}

void VS_CC analyseFree(void *instanceData, VSCore *, const VSAPI *vsapi) {
    AnalyseData *d = (AnalyseData *)instanceData;
    unregisterSelf(d->self);
    vsapi->freeNode(d->node);
    if (d->clip)
        vsapi->freeNode(d->clip);
    delete d->fused;
    delete d->cache;
    delete d;
}

void VS_CC analyseCreate(const VSMap *in, VSMap *out, void *, VSCore *core, const VSAPI *vsapi) {
    mvb200_analyse_args a;
    static const char *names[] = {"blksize", "blksizev", "levels", "search", "searchparam", "pelsearch", "isb", "lambda", "chroma", "delta", "truemotion", "lsad",
                                  "plevel", "global", "pnew", "pzero", "pglobal", "overlap", "overlapv", "divide", "badsad", "badrange", "opt", "meander", "trymany",
                                  "fields", "tff", "search_coarse", "dct"};
    int32_t *fields = (int32_t *)&a;
    for (size_t i = 0; i < sizeof(names) / sizeof(names[0]); i++)
        fields[i] = argInt(vsapi, in, names[i]);
    AnalyseData d;
    d.node = vsapi->mapGetNode(in, "super", 0, nullptr);
    d.vi = vsapi->getVideoInfo(d.node);
    if (!formatOk(d.vi)) {
        vsapi->mapSetError(out, "Analyse: Input clip must be GRAY, 420, 422, 440, or 444, up to 16 bits, with constant format and dimensions.");
        vsapi->freeNode(d.node);
        return;
    }
    std::string err;
    // argument validation that does not need the super clip comes first in the reference; mvb200_analyse_init
    // keeps that order internally, so only the clip lookup happens here
    if (!superFromClip(d.node, vsapi, "Analyse", &d.sp, err) || mvb200_analyse_init(&a, &d.sp, d.vi->numFrames, &d.ap)) {
        vsapi->mapSetError(out, err.empty() ? mvb200_last_error() : err.c_str());
        vsapi->freeNode(d.node);
        return;
    }
    // frames searched per engine call (see AnalyseCache).  Automatic: a request holds the window's SUPER frames (pel^2 + 1/3
    // times the size of a source frame each), so the window stays well below what fills the GPU -- the fused paths do that
    d.lookahead = lookaheadFromEnv();
    if (d.lookahead < 0) {
        mvb200_search_plan plan;
        d.lookahead = 32;
        if (mvb200_search_plan_query(&d.ap, &d.sp, 0, &plan) == 0 && plan.chains_per_wave < d.lookahead)
            d.lookahead = plan.chains_per_wave < 2 ? 0 : plan.chains_per_wave;
        if (d.lookahead > d.vi->numFrames)
            d.lookahead = d.vi->numFrames < 2 ? 0 : d.vi->numFrames;
    }
    d.cache = d.lookahead ? new AnalyseCache : nullptr;
    d.self = nullptr;
    d.fused = nullptr;
    d.clip = nullptr;
    if (d.lookahead && !d.ap.fields && d.ap.ad.nDeltaFrame > 0) { // the fused path: the super clip is this plugin's own mv.Super (no pelclip)
        std::lock_guard<std::mutex> lock(g_registry_mutex);
        const auto su = g_registry.find(d.node);
        if (su != g_registry.end() && su->second.kind == kNodeSuper) {
            const SuperData *S = (const SuperData *)su->second.data;
            const VSVideoInfo *cvi = vsapi->getVideoInfo(S->node);
            if (!S->usePelClip && cvi->numFrames == d.vi->numFrames) {
                d.clip = vsapi->addNodeRef(S->node);
                d.ssp = S->sp;
                mvb200_super_layout(&d.ssp, &d.sl);
                d.vstride = ((int64_t)d.ap.vectors_size + 255) / 256 * 256;
                d.fused = new FusedCache;
                mvb200_clip_layout(&d.ssp, &d.fused->cl);
                d.fused->out = d.sl;
                d.fused->outFrameBytes = (size_t)d.sl.frame_bytes;
                d.fused->outExtra = (size_t)d.vstride;
                d.fused->window = fusedWindowFrames(d.fused, lookaheadFromEnv(), &d.ap, &d.ssp, 1);
                d.lookahead = 0; // (the windows replace the look-ahead over super frames; frame 0 alone is searched as one chain)
                delete d.cache;
                d.cache = nullptr;
            }
        }
    }
    AnalyseData *data = new AnalyseData(d);
    if (data->fused) { // the Super's source clip is what the windows read; the super clip only for frame 0 at script-building time
        VSFilterDependency fdeps[2] = {{data->clip, rpGeneral}, {data->node, rpGeneral}};
        vsapi->createVideoFilter(out, "Analyse", data->vi, analyseFusedGetFrame, analyseFree, fmParallel, fdeps, 2, data, core);
    } else {
        VSFilterDependency deps[1] = {{data->node, rpGeneral}};
        vsapi->createVideoFilter(out, "Analyse", data->vi, analyseGetFrame, analyseFree, fmParallel, deps, 1, data, core);
    }
    data->self = registerSelf(vsapi, out, kNodeAnalyse, data);
}

struct RecalcData {
    VSNode *node, *vectors;
    const VSVideoInfo *vi;
    mvb200_super_params sp;
    mvb200_recalculate_params rp;
    VSNode *self; // this filter's output node (registry key), not a reference
};

// mv.Recalculate passes between a vector clip's mv.Analyse and its consumer (the refinemotion shape: Analyse -> Recalculate
// (-> Recalculate) -> DegrainN / Compensate), innermost first
constexpr int kMaxRecalc = 2;
struct FusedRecalc {
    int n = 0;
    mvb200_recalculate_params rp[kMaxRecalc];
};

// ------------------------------------------------- the canonical chain (recognised by the fused paths)
// True when `super` is this plugin's mv.Super of `clip` (no pelclip) and every vector clip is this plugin's mv.Analyse of
// that super clip with a relative delta and no field handling: the chain one *_clip_host engine call runs from SOURCE
// frames.  perPixel (the Flow* consumers): also chroma planes in the super clip unless the clip is gray, and no divide.
// *spOut / ap[i]: the nodes' own parameters (the Super's own sharp / rfilter: the frame properties do not carry them).
// rc (block consumers only; null = not accepted): a vector clip may also be up to kMaxRecalc of this plugin's mv.Recalculate
// nodes (same super clip, no field handling) on top of such an mv.Analyse; rc[i] receives their parameters, innermost first.
bool canonicalChain(VSNode *clip, VSNode *super, VSNode *const *vectors, int nvec, const mvb200_super_params &sp, int numFrames, bool perPixel, mvb200_super_params *spOut, mvb200_analyse_params *ap,
                    FusedRecalc *rc = nullptr) {
    std::lock_guard<std::mutex> lock(g_registry_mutex);
    const auto su = g_registry.find(super);
    if (su == g_registry.end() || su->second.kind != kNodeSuper)
        return false;
    const SuperData *S = (const SuperData *)su->second.data;
    if (S->node != clip || S->usePelClip || S->sp.width != sp.width || S->sp.height != sp.height || S->sp.bits != sp.bits || S->sp.pel != sp.pel || S->sp.hpad != sp.hpad || S->sp.vpad != sp.vpad || S->sp.levels != sp.levels ||
        S->sp.modeyuv != sp.modeyuv || (perPixel && !((S->sp.modeyuv & (MVB200_UPLANE | MVB200_VPLANE)) || sp.gray)))
        return false;
    for (int i = 0; i < nvec; i++) {
        VSNode *v = vectors[i];
        const RecalcData *passes[kMaxRecalc];
        int npass = 0;
        auto an = g_registry.find(v);
        while (an != g_registry.end() && an->second.kind == kNodeRecalc) { // outermost pass first
            const RecalcData *R = (const RecalcData *)an->second.data;
            if (!rc || npass == kMaxRecalc || R->node != super || R->rp.fields || R->rp.num_frames != numFrames)
                return false;
            passes[npass++] = R;
            an = g_registry.find(R->vectors);
        }
        if (an == g_registry.end() || an->second.kind != kNodeAnalyse)
            return false;
        if (rc) {
            rc[i].n = npass;
            for (int k = 0; k < npass; k++)
                rc[i].rp[k] = passes[npass - 1 - k]->rp;
        }
        const AnalyseData *A = (const AnalyseData *)an->second.data;
        if (A->node != super || A->ap.fields || (perPixel && A->ap.divide) || A->ap.ad.nDeltaFrame <= 0 || A->ap.num_frames != numFrames)
            return false;
        ap[i] = A->ap;
    }
    *spOut = S->sp;
    return true;
}

// ---------------------------------------------------------------- Degrain
// Fused path (see fusedGetFrame): ONE mvb200_degrain_clip_host_async call per window -- H2D of 3 MB source frames, Super,
// 2r searches of B chains each, DegrainN, D2H -- instead of 1 + 2r + 1 per-frame calls that move 17.9 MB super frames
// each way.  Results are identical (the engine's clip pipeline is what the sharded bench runs; frames outside the
// window only serve as references).
struct DegrainFused : FusedCache {
    mvb200_analyse_params ap[2 * MVB200_MAX_RADIUS]; // copies: no dependence on the Analyse instances' lifetime
    FusedRecalc rc[2 * MVB200_MAX_RADIUS];           // mv.Recalculate passes of each vector clip (usually none)
    int halo;
};

struct DegrainData {
    VSNode *node, *super, *vectors[2 * MVB200_MAX_RADIUS];
    const VSVideoInfo *vi;
    mvb200_super_params sp;
    mvb200_degrain_params dp;
    int radius;
    DegrainFused *fused; // owned; null = node-by-node evaluation
};

const VSFrame *VS_CC degrainFusedGetFrame(int n, int activationReason, void *instanceData, void **frameData, VSFrameContext *frameCtx, VSCore *core, const VSAPI *vsapi) {
    DegrainData *d = (DegrainData *)instanceData;
    DegrainFused *f = d->fused;
    const int N = d->vi->numFrames;
    unsigned mask = 0; // planes the filter does not process are copies of the source (MVDegrains.cpp:211-214)
    for (int p = 0; p < 3; p++)
        mask |= d->dp.process[p] ? 1u << p : 0u;
    auto segment = [&](int w0, int w1, int *s0, int *s1) { // the window and the halo the searches and the blend reach into
        *s0 = w0 - f->halo > 0 ? w0 - f->halo : 0;
        *s1 = w1 + f->halo < N ? w1 + f->halo : N;
    };
    auto run = [&](mvb200_ctx *ctx, const void *hostSrc, int s0, int s1, int w0, int w1, void *hostDst, int32_t *flags) {
        const int R2 = d->radius * 2, seg = s1 - s0;
        mvb200_analyse_params aps[2 * MVB200_MAX_RADIUS];
        const mvb200_analyse_params *app[2 * MVB200_MAX_RADIUS];
        for (int r = 0; r < R2; r++) { // the segment is the clip as far as the engine is concerned
            aps[r] = f->ap[r];
            aps[r].num_frames = seg;
            app[r] = &aps[r];
        }
        const mvb200_recalculate_params *rpp[2 * MVB200_MAX_RADIUS * kMaxRecalc];
        for (int r = 0; r < R2; r++)
            for (int k = 0; k < kMaxRecalc; k++)
                rpp[r * kMaxRecalc + k] = k < f->rc[r].n ? &f->rc[r].rp[k] : nullptr;
        // (asynchronous: fusedGetFrame records the completion event)
        return mvb200_degrain_recalc_clip_host_async(ctx, &d->sp, app, R2, rpp, kMaxRecalc, &d->dp, hostSrc, hostDst, seg, w0 - s0, w1 - w0);
    };
    static const char *const names[] = {"Degrain", "Degrain1", "Degrain2", "Degrain3", "Degrain4", "Degrain5", "Degrain6"}; // (the filter's own name in messages, as at create time)
    return fusedGetFrame(names[d->radius >= 1 && d->radius <= 6 ? d->radius : 0], f, d->node, d->vi, n, n, mask, activationReason, frameData, frameCtx, core, vsapi, segment, run);
}

const VSFrame *VS_CC degrainGetFrame(int n, int activationReason, void *instanceData, void **, VSFrameContext *frameCtx, VSCore *core, const VSAPI *vsapi) {
    DegrainData *d = (DegrainData *)instanceData;
    const int R2 = d->radius * 2;
    if (activationReason == arInitial) { // MVDegrains.cpp:92-109
        for (int r = 0; r < R2; r += 2) {
            vsapi->requestFrameFilter(n, d->vectors[r], frameCtx);
            vsapi->requestFrameFilter(n, d->vectors[r + 1], frameCtx);
            int offB = d->dp.vd[r].nDeltaFrame;
            if (n + offB < d->vi->numFrames)
                vsapi->requestFrameFilter(n + offB, d->super, frameCtx);
            int offF = -d->dp.vd[r + 1].nDeltaFrame;
            if (n + offF >= 0)
                vsapi->requestFrameFilter(n + offF, d->super, frameCtx);
        }
        vsapi->requestFrameFilter(n, d->node, frameCtx);
    } else if (activationReason == arAllFramesReady) {
        const VSFrame *src = vsapi->getFrameFilter(n, d->node, frameCtx);
        VSFrame *dst = vsapi->newVideoFrame(&d->vi->format, d->vi->width, d->vi->height, src, core);
        const VSFrame *vecFrames[2 * MVB200_MAX_RADIUS] = {nullptr};
        const VSFrame *refFrames[2 * MVB200_MAX_RADIUS] = {nullptr};
        const uint8_t *vecs[2 * MVB200_MAX_RADIUS];
        const uint8_t *refs[2 * MVB200_MAX_RADIUS * 3];
        ptrdiff_t refStrides[2 * MVB200_MAX_RADIUS * 3];
        bool bad = false;
        for (int r = 0; r < R2; r++) {
            vecFrames[r] = vsapi->getFrameFilter(n, d->vectors[r], frameCtx);
            int err = 0;
            vecs[r] = vectorBlob(vsapi, vecFrames[r], &err);
            bad = bad || err;
            int nref = n + d->dp.vd[r].nDeltaFrame * (d->dp.vd[r].isBackward ? 1 : -1);
            if (nref >= 0 && nref < d->vi->numFrames)
                refFrames[r] = vsapi->getFrameFilter(nref, d->super, frameCtx);
            planesRO(vsapi, refFrames[r], d->vi->format.numPlanes, refs + r * 3, refStrides + r * 3);
        }
        const uint8_t *sp_[3];
        ptrdiff_t ss[3];
        uint8_t *dp_[3];
        ptrdiff_t ds[3];
        planesRO(vsapi, src, d->vi->format.numPlanes, sp_, ss);
        planesRW(vsapi, dst, d->vi->format.numPlanes, dp_, ds);
        std::string msg = "MVTools_vectors property missing";
        int rc = 1;
        if (!bad) {
            std::lock_guard<std::mutex> lock(g_engine_mutex);
            mvb200_ctx *ctx = engine(msg);
            rc = ctx ? mvb200_degrain_frame(ctx, &d->dp, &d->sp, sp_, ss, refs, refStrides, vecs, dp_, ds) : 1;
            if (ctx && rc)
                msg = mvb200_last_error();
        }
        for (int r = 0; r < R2; r++) {
            vsapi->freeFrame(vecFrames[r]);
            vsapi->freeFrame(refFrames[r]);
        }
        vsapi->freeFrame(src);
        if (rc) {
            vsapi->setFilterError(("Degrain: " + msg).c_str(), frameCtx);
            vsapi->freeFrame(dst);
            return nullptr;
        }
        return dst;
    }
    return nullptr;
}

void VS_CC degrainFree(void *instanceData, VSCore *, const VSAPI *vsapi) {
    DegrainData *d = (DegrainData *)instanceData;
    vsapi->freeNode(d->node);
    vsapi->freeNode(d->super);
    for (int r = 0; r < d->radius * 2; r++)
        vsapi->freeNode(d->vectors[r]);
    delete d->fused;
    delete d;
}

void VS_CC degrainCreate(const VSMap *in, VSMap *out, void *userData, VSCore *core, const VSAPI *vsapi) {
    const int radius = (int)(intptr_t)userData;
    const std::string filter = "Degrain" + std::to_string(radius);
    static const char *vnames[12] = {"mvbw", "mvfw", "mvbw2", "mvfw2", "mvbw3", "mvfw3", "mvbw4", "mvfw4", "mvbw5", "mvfw5", "mvbw6", "mvfw6"};
    mvb200_degrain_args a;
    a.radius = radius;
    a.thsad = argInt64(vsapi, in, "thsad");
    a.thsadc = argInt64(vsapi, in, "thsadc");
    a.plane = argInt(vsapi, in, "plane");
    a.limit = argInt(vsapi, in, "limit");
    a.limitc = argInt(vsapi, in, "limitc");
    a.thscd1 = argInt64(vsapi, in, "thscd1");
    a.thscd2 = argInt(vsapi, in, "thscd2");
    DegrainData d;
    memset(&d, 0, sizeof(d));
    d.radius = radius;
    auto freeAll = [&]() {
        vsapi->freeNode(d.node);
        vsapi->freeNode(d.super);
        for (int r = 0; r < 2 * radius; r++)
            vsapi->freeNode(d.vectors[r]);
    };
    d.super = vsapi->mapGetNode(in, "super", 0, nullptr);
    std::string err;
    if (!superFromClip(d.super, vsapi, filter.c_str(), &d.sp, err)) {
        vsapi->mapSetError(out, err.c_str());
        freeAll();
        return;
    }
    mvb200_analysis_data vd[2 * MVB200_MAX_RADIUS];
    for (int r = 0; r < 2 * radius; r++) { // adataFromVectorClip, MVAnalysisData.c:34-65
        d.vectors[r] = vsapi->mapGetNode(in, vnames[r], 0, nullptr);
        char msg[1024];
        const VSFrame *evil = vsapi->getFrame(0, d.vectors[r], msg, sizeof(msg));
        if (!evil) {
            err = filter + ": Failed to retrieve first frame from " + vnames[r] + ". Error message: " + msg;
            break;
        }
        int e = 0;
        const VSMap *props = vsapi->getFramePropertiesRO(evil);
        const char *data = vsapi->mapGetData(props, "MVTools_MVAnalysisData", 0, &e);
        int size = e ? 0 : vsapi->mapGetDataSize(props, "MVTools_MVAnalysisData", 0, nullptr);
        if (e)
            err = filter + ": Property 'MVTools_MVAnalysisData' not found in first frame of " + vnames[r] + ".";
        else if (size != (int)sizeof(mvb200_analysis_data))
            err = filter + ": Property 'MVTools_MVAnalysisData' in first frame of " + vnames[r] + " has wrong size (" + std::to_string(size) + " instead of " +
                  std::to_string(sizeof(mvb200_analysis_data)) + ").";
        else
            memcpy(&vd[r], data, sizeof(vd[r]));
        vsapi->freeFrame(evil);
        if (!err.empty())
            break;
    }
    if (!err.empty()) {
        vsapi->mapSetError(out, err.c_str());
        freeAll();
        return;
    }
    d.node = vsapi->mapGetNode(in, "clip", 0, nullptr);
    d.vi = vsapi->getVideoInfo(d.node);
    if (mvb200_degrain_init(&a, &d.sp, vd, d.vi->width, d.vi->height, &d.dp)) {
        vsapi->mapSetError(out, mvb200_last_error());
        freeAll();
        return;
    }
    if (!formatOk(d.vi)) {
        vsapi->mapSetError(out, (filter + ": input clip must be GRAY, 420, 422, 440, or 444, up to 16 bits, with constant dimensions.").c_str());
        freeAll();
        return;
    }
    // the fused path: every input is one of this plugin's own nodes, wired as in the canonical script
    const int window = lookaheadFromEnv();
    if (window) {
        DegrainFused *f = new DegrainFused;
        mvb200_super_params ssp;
        if (canonicalChain(d.node, d.super, d.vectors, 2 * radius, d.sp, d.vi->numFrames, false, &ssp, f->ap, f->rc)) {
            int halo = 0;
            for (int r = 0; r < 2 * radius; r++)
                halo = f->ap[r].ad.nDeltaFrame > halo ? f->ap[r].ad.nDeltaFrame : halo;
            d.sp = ssp;
            mvb200_clip_layout(&d.sp, &f->cl);
            f->window = fusedWindowFrames(f, window, f->ap, &d.sp, 2 * radius);
            f->halo = halo;
            d.fused = f;
        } else {
            delete f;
        }
    }
    DegrainData *data = new DegrainData(d);
    if (data->fused) { // only the source clip is read on this path
        VSFilterDependency fdeps[1] = {{data->node, rpGeneral}};
        vsapi->createVideoFilter(out, filter.c_str(), data->vi, degrainFusedGetFrame, degrainFree, fmParallel, fdeps, 1, data, core);
        return;
    }
    VSFilterDependency deps[14] = {{data->node, rpStrictSpatial}, {data->super, rpGeneral}};
    for (int r = 0; r < 12; r++)
        deps[2 + r] = {data->vectors[r], rpStrictSpatial};
    // the reference passes numDeps = 2 + radius (sic, MVDegrains.cpp:790-808); all vector clips are listed here
    vsapi->createVideoFilter(out, filter.c_str(), data->vi, degrainGetFrame, degrainFree, fmParallel, deps, 2 + 2 * radius, data, core);
}

// adataFromVectorClip, MVAnalysisData.c:34-65
bool adataFromClip(VSNode *clip, const VSAPI *vsapi, const std::string &filter, const char *name, mvb200_analysis_data *ad, std::string &err) {
    char msg[1024];
    const VSFrame *evil = vsapi->getFrame(0, clip, msg, sizeof(msg));
    if (!evil) {
        err = filter + ": Failed to retrieve first frame from " + name + ". Error message: " + msg;
        return false;
    }
    int e = 0;
    const VSMap *props = vsapi->getFramePropertiesRO(evil);
    const char *data = vsapi->mapGetData(props, "MVTools_MVAnalysisData", 0, &e);
    int size = e ? 0 : vsapi->mapGetDataSize(props, "MVTools_MVAnalysisData", 0, nullptr);
    if (e)
        err = filter + ": Property 'MVTools_MVAnalysisData' not found in first frame of " + name + ".";
    else if (size != (int)sizeof(mvb200_analysis_data))
        err = filter + ": Property 'MVTools_MVAnalysisData' in first frame of " + name + " has wrong size (" + std::to_string(size) + " instead of " +
              std::to_string(sizeof(mvb200_analysis_data)) + ").";
    else
        memcpy(ad, data, sizeof(*ad));
    vsapi->freeFrame(evil);
    return err.empty();
}

// field shift from _Field props / tff (MVAnalyse.c:146-179, MVCompensate.c:191-229, MVRecalculate.c:112-166)
bool fieldShiftOf(const VSAPI *vsapi, const VSFrame *src, const VSFrame *ref, int n, int nref, bool fields, bool tffExists, int tff, int nPel, bool oddDelta, int *shift) {
    *shift = 0;
    int err = 0;
    int srcTop = !!vsapi->mapGetInt(vsapi->getFramePropertiesRO(src), "_Field", 0, &err);
    bool missing = err && fields && !tffExists;
    int refTop = 0;
    if (ref) {
        refTop = !!vsapi->mapGetInt(vsapi->getFramePropertiesRO(ref), "_Field", 0, &err);
        missing = missing || (err && fields && !tffExists);
    }
    if (missing)
        return false;
    if (tffExists) {
        srcTop = tff ^ (n % 2);
        refTop = tff ^ (nref % 2);
    }
    if (ref && fields && nPel > 1 && oddDelta)
        *shift = (srcTop && !refTop) ? nPel / 2 : ((refTop && !srcTop) ? -(nPel / 2) : 0);
    return true;
}

// ------------------------------------------------------------ Recalculate

const VSFrame *VS_CC recalcGetFrame(int n, int activationReason, void *instanceData, void **, VSFrameContext *frameCtx, VSCore *core, const VSAPI *vsapi) {
    RecalcData *d = (RecalcData *)instanceData;
    const int off = d->rp.ad.nDeltaFrame;
    const int nref = off > 0 ? (d->rp.ad.isBackward ? n + off : n - off) : -off;
    const bool refOk = nref >= 0 && nref < d->vi->numFrames;
    if (activationReason == arInitial) { // MVRecalculate.c:73-94
        vsapi->requestFrameFilter(n, d->vectors, frameCtx);
        if (refOk) {
            vsapi->requestFrameFilter(n < nref ? n : nref, d->node, frameCtx);
            vsapi->requestFrameFilter(n < nref ? nref : n, d->node, frameCtx);
        } else {
            vsapi->requestFrameFilter(n, d->node, frameCtx);
        }
    } else if (activationReason == arAllFramesReady) {
        const VSFrame *src = vsapi->getFrameFilter(n, d->node, frameCtx);
        const VSFrame *ref = refOk ? vsapi->getFrameFilter(nref, d->node, frameCtx) : nullptr;
        const VSFrame *mvn = vsapi->getFrameFilter(n, d->vectors, frameCtx);
        int e = 0;
        const uint8_t *old = vectorBlob(vsapi, mvn, &e);
        int fieldShift = 0;
        std::string msg;
        int rc = 1;
        if (e) {
            msg = "MVTools_vectors property missing";
        } else if (!fieldShiftOf(vsapi, src, ref, n, nref, d->rp.fields, d->rp.tff_exists, d->rp.tff, d->rp.ad.nPel, d->rp.ad.nDeltaFrame % 2, &fieldShift)) {
            msg = "_Field property not found in input frame. Therefore, you must pass tff argument.";
        } else {
            std::vector<uint8_t> vectors((size_t)d->rp.vectors_size);
            const uint8_t *sp_[3], *rp_[3];
            ptrdiff_t ss[3], rs[3];
            planesRO(vsapi, src, d->vi->format.numPlanes, sp_, ss);
            planesRO(vsapi, ref, d->vi->format.numPlanes, rp_, rs);
            {
                std::lock_guard<std::mutex> lock(g_engine_mutex);
                mvb200_ctx *ctx = engine(msg);
                rc = ctx ? mvb200_recalculate_frame(ctx, &d->rp, &d->sp, sp_, ss, ref ? rp_ : nullptr, ref ? rs : nullptr, old, fieldShift, vectors.data()) : 1;
                if (ctx && rc)
                    msg = mvb200_last_error();
            }
            if (!rc) {
                VSFrame *dst = vsapi->copyFrame(src, core);
                VSMap *props = vsapi->getFramePropertiesRW(dst);
                const mvb200_analysis_data *ad = d->rp.divide ? &d->rp.ad_divided : &d->rp.ad;
                vsapi->mapSetData(props, "MVTools_MVAnalysisData", (const char *)ad, sizeof(*ad), dtBinary, maReplace);
                vsapi->mapSetData(props, "MVTools_vectors", (const char *)vectors.data(), (int)vectors.size(), dtBinary, maReplace);
                vsapi->freeFrame(mvn);
                vsapi->freeFrame(ref);
                vsapi->freeFrame(src);
                return dst;
            }
        }
        vsapi->freeFrame(mvn);
        vsapi->freeFrame(ref);
        vsapi->freeFrame(src);
        vsapi->setFilterError(("Recalculate: " + msg).c_str(), frameCtx);
        return nullptr;
    }
    return nullptr;
}

void VS_CC recalcFree(void *instanceData, VSCore *, const VSAPI *vsapi) {
    RecalcData *d = (RecalcData *)instanceData;
    unregisterSelf(d->self);
    vsapi->freeNode(d->node);
    vsapi->freeNode(d->vectors);
    delete d;
}

void VS_CC recalcCreate(const VSMap *in, VSMap *out, void *, VSCore *core, const VSAPI *vsapi) {
    mvb200_recalculate_args a;
    a.thsad = argInt64(vsapi, in, "thsad");
    static const char *names[] = {"smooth", "blksize", "blksizev", "search", "searchparam", "lambda", "chroma", "truemotion", "pnew", "overlap", "overlapv", "divide",
                                  "opt", "meander", "fields", "tff", "dct"};
    int32_t *fields = &a.smooth;
    for (size_t i = 0; i < sizeof(names) / sizeof(names[0]); i++)
        fields[i] = argInt(vsapi, in, names[i]);
    RecalcData d;
    memset(&d, 0, sizeof(d));
    d.node = vsapi->mapGetNode(in, "super", 0, nullptr);
    d.vectors = vsapi->mapGetNode(in, "vectors", 0, nullptr);
    d.vi = vsapi->getVideoInfo(d.node);
    std::string err;
    mvb200_analysis_data old;
    if (!superFromClip(d.node, vsapi, "Recalculate", &d.sp, err) || !adataFromClip(d.vectors, vsapi, "Recalculate", "vectors", &old, err) ||
        mvb200_recalculate_init(&a, &d.sp, &old, d.vi->numFrames, &d.rp)) {
        vsapi->mapSetError(out, err.empty() ? mvb200_last_error() : err.c_str());
        vsapi->freeNode(d.node);
        vsapi->freeNode(d.vectors);
        return;
    }
    RecalcData *data = new RecalcData(d);
    VSFilterDependency deps[2] = {{data->node, rpGeneral}, {data->vectors, rpStrictSpatial}};
    vsapi->createVideoFilter(out, "Recalculate", data->vi, recalcGetFrame, recalcFree, fmParallel, deps, 2, data, core);
    data->self = registerSelf(vsapi, out, kNodeRecalc, data);
}

// ------------------------------------------------------------- Compensate
// Fused path (see fusedGetFrame): a window of B frames is one mvb200_compensate_clip_host call over source frames.
struct CompFused : FusedCache {
    mvb200_analyse_params ap;
    FusedRecalc rc; // mv.Recalculate passes between the search and the compensation (usually none)
    int halo;
};

struct CompData {
    VSNode *node, *super, *vectors;
    const VSVideoInfo *vi, *supervi;
    mvb200_super_params sp;
    mvb200_compensate_params cp;
    CompFused *fused; // owned; null = node-by-node evaluation
};

const VSFrame *VS_CC compFusedGetFrame(int n, int activationReason, void *instanceData, void **frameData, VSFrameContext *frameCtx, VSCore *core, const VSAPI *vsapi) {
    CompData *d = (CompData *)instanceData;
    CompFused *f = d->fused;
    const int N = d->vi->numFrames;
    const unsigned mask = (d->cp.super_modeyuv & (MVB200_UPLANE | MVB200_VPLANE)) ? 7u : 1u; // only the super clip's planes are produced
    auto segment = [&](int w0, int w1, int *s0, int *s1) {
        *s0 = w0 - f->halo > 0 ? w0 - f->halo : 0;
        *s1 = w1 + f->halo < N ? w1 + f->halo : N;
    };
    auto run = [&](mvb200_ctx *ctx, const void *hostSrc, int s0, int s1, int w0, int w1, void *hostDst, int32_t *flags) {
        const mvb200_recalculate_params *rpp[kMaxRecalc];
        for (int k = 0; k < kMaxRecalc; k++)
            rpp[k] = k < f->rc.n ? &f->rc.rp[k] : nullptr;
        return mvb200_compensate_recalc_clip_host_async(ctx, &d->sp, &f->ap, rpp, kMaxRecalc, &d->cp, hostSrc, s0, s1 - s0, w0, w1 - w0, hostDst);
    };
    const VSFrame *out = fusedGetFrame("Compensate", f, d->node, d->vi, n, n, mask, activationReason, frameData, frameCtx, core, vsapi, segment, run);
    if (out && n == 0) // the reference builds the output on super frame n (MVCompensate.c:113): frame 0 carries the Super_* properties
        setSuperProps(vsapi, const_cast<VSFrame *>(out), d->sp);
    return out;
}

const VSFrame *VS_CC compGetFrame(int n, int activationReason, void *instanceData, void **, VSFrameContext *frameCtx, VSCore *core, const VSAPI *vsapi) {
    CompData *d = (CompData *)instanceData;
    const int delta = d->cp.vd.nDeltaFrame;
    const int nref = delta > 0 ? n + (d->cp.vd.isBackward ? delta : -delta) : -delta;
    if (activationReason == arInitial) { // MVCompensate.c:78-97
        vsapi->requestFrameFilter(n, d->vectors, frameCtx);
        if (nref < n && nref >= 0)
            vsapi->requestFrameFilter(nref, d->super, frameCtx);
        vsapi->requestFrameFilter(n, d->super, frameCtx);
        if (nref >= n && nref < d->vi->numFrames)
            vsapi->requestFrameFilter(nref, d->super, frameCtx);
    } else if (activationReason == arAllFramesReady) {
        const VSFrame *src = vsapi->getFrameFilter(n, d->super, frameCtx);
        const VSFrame *ref = (nref >= 0 && nref < d->vi->numFrames) ? vsapi->getFrameFilter(nref, d->super, frameCtx) : nullptr;
        const VSFrame *mvn = vsapi->getFrameFilter(n, d->vectors, frameCtx);
        VSFrame *dst = vsapi->newVideoFrame(&d->vi->format, d->vi->width, d->vi->height, src, core);
        int e = 0;
        const uint8_t *vec = vectorBlob(vsapi, mvn, &e);
        int fieldShift = 0;
        std::string msg;
        int rc = 1;
        if (e) {
            msg = "MVTools_vectors property missing";
        } else if (!fieldShiftOf(vsapi, src, ref, n, nref, d->cp.fields && d->cp.vd.nPel > 1 && (nref - n) % 2 != 0 && ref, d->cp.tff_exists, d->cp.tff, d->cp.vd.nPel, true, &fieldShift)) {
            msg = "_Field property not found in input frame. Therefore, you must pass tff argument.";
        } else {
            const uint8_t *sp_[3], *rp_[3];
            ptrdiff_t ss[3], rs[3];
            uint8_t *dp_[3];
            ptrdiff_t ds[3];
            planesRO(vsapi, src, d->supervi->format.numPlanes, sp_, ss);
            planesRO(vsapi, ref, d->supervi->format.numPlanes, rp_, rs);
            planesRW(vsapi, dst, d->vi->format.numPlanes, dp_, ds);
            std::lock_guard<std::mutex> lock(g_engine_mutex);
            mvb200_ctx *ctx = engine(msg);
            rc = ctx ? mvb200_compensate_frame(ctx, &d->cp, &d->sp, sp_, ss, ref ? rp_ : nullptr, ref ? rs : nullptr, vec, fieldShift, dp_, ds) : 1;
            if (ctx && rc)
                msg = mvb200_last_error();
        }
        vsapi->freeFrame(mvn);
        vsapi->freeFrame(ref);
        vsapi->freeFrame(src);
        if (rc) {
            vsapi->setFilterError(("Compensate: " + msg).c_str(), frameCtx);
            vsapi->freeFrame(dst);
            return nullptr;
        }
        return dst;
    }
    return nullptr;
}

void VS_CC compFree(void *instanceData, VSCore *, const VSAPI *vsapi) {
    CompData *d = (CompData *)instanceData;
    vsapi->freeNode(d->node);
    vsapi->freeNode(d->super);
    vsapi->freeNode(d->vectors);
    delete d->fused;
    delete d;
}

void VS_CC compCreate(const VSMap *in, VSMap *out, void *, VSCore *core, const VSAPI *vsapi) {
    mvb200_compensate_args a;
    a.scbehavior = argInt(vsapi, in, "scbehavior");
    a.thsad = argInt64(vsapi, in, "thsad");
    a.fields = argInt(vsapi, in, "fields");
    int err = 0;
    a.time = vsapi->mapGetFloat(in, "time", 0, &err);
    if (err)
        a.time = NAN;
    a.thscd1 = argInt64(vsapi, in, "thscd1");
    a.thscd2 = argInt(vsapi, in, "thscd2");
    a.tff = argInt(vsapi, in, "tff");
    if (a.time == a.time && (a.time < 0.0 || a.time > 100.0)) { // checked before any clip is touched (MVCompensate.c:457-460)
        vsapi->mapSetError(out, "Compensate: time must be between 0.0 and 100.0 (inclusive).");
        return;
    }
    CompData d;
    memset(&d, 0, sizeof(d));
    auto freeAll = [&]() {
        vsapi->freeNode(d.node);
        vsapi->freeNode(d.super);
        vsapi->freeNode(d.vectors);
    };
    d.super = vsapi->mapGetNode(in, "super", 0, nullptr);
    d.supervi = vsapi->getVideoInfo(d.super);
    std::string msg;
    mvb200_analysis_data vd;
    if (!superFromClip(d.super, vsapi, "Compensate", &d.sp, msg)) {
        vsapi->mapSetError(out, msg.c_str());
        freeAll();
        return;
    }
    d.vectors = vsapi->mapGetNode(in, "vectors", 0, nullptr);
    if (!adataFromClip(d.vectors, vsapi, "Compensate", "vectors", &vd, msg)) {
        vsapi->mapSetError(out, msg.c_str());
        freeAll();
        return;
    }
    d.node = vsapi->mapGetNode(in, "clip", 0, nullptr);
    d.vi = vsapi->getVideoInfo(d.node);
    if (mvb200_compensate_init(&a, &d.sp, &vd, d.vi->width, d.vi->height, d.vi->numFrames, &d.cp)) {
        vsapi->mapSetError(out, mvb200_last_error());
        freeAll();
        return;
    }
    if (!formatOk(d.vi)) {
        vsapi->mapSetError(out, "Compensate: input clip must be GRAY, 420, 422, 440, or 444, up to 16 bits, with constant dimensions.");
        freeAll();
        return;
    }
    const int window = lookaheadFromEnv();
    mvb200_super_params ssp;
    mvb200_analyse_params sap;
    FusedRecalc src;
    if (window && !d.cp.fields && canonicalChain(d.node, d.super, &d.vectors, 1, d.sp, d.vi->numFrames, false, &ssp, &sap, &src)) { // the fused path
        d.fused = new CompFused;
        d.fused->ap = sap;
        d.fused->rc = src;
        d.sp = ssp;
        mvb200_clip_layout(&d.sp, &d.fused->cl);
        d.fused->window = fusedWindowFrames(d.fused, window, &sap, &d.sp, 1);
        d.fused->halo = sap.ad.nDeltaFrame;
    }
    CompData *data = new CompData(d);
    if (data->fused) { // only the source clip is read on this path
        VSFilterDependency fdeps[1] = {{data->node, rpGeneral}};
        vsapi->createVideoFilter(out, "Compensate", data->vi, compFusedGetFrame, compFree, fmParallel, fdeps, 1, data, core);
        return;
    }
    VSFilterDependency deps[2] = {{data->super, rpGeneral}, {data->vectors, rpGeneral}};
    vsapi->createVideoFilter(out, "Compensate", data->vi, compGetFrame, compFree, fmParallel, deps, 2, data, core);
}

// ------------------------------------------------------------ SCDetection
// mvscdetectionCreate / GetFrame, MVSCDetection.c:43-131
struct ScdData {
    VSNode *node, *vectors;
    const VSVideoInfo *vi;
    mvb200_scdetection_params p;
};

const VSFrame *VS_CC scdGetFrame(int n, int activationReason, void *instanceData, void **, VSFrameContext *frameCtx, VSCore *core, const VSAPI *vsapi) {
    ScdData *d = (ScdData *)instanceData;
    if (activationReason == arInitial) {
        vsapi->requestFrameFilter(n, d->vectors, frameCtx);
        vsapi->requestFrameFilter(n, d->node, frameCtx);
    } else if (activationReason == arAllFramesReady) {
        const VSFrame *src = vsapi->getFrameFilter(n, d->node, frameCtx);
        VSFrame *dst = vsapi->copyFrame(src, core);
        vsapi->freeFrame(src);
        const VSFrame *mvn = vsapi->getFrameFilter(n, d->vectors, frameCtx);
        int e = 0;
        const uint8_t *vec = vectorBlob(vsapi, mvn, &e);
        std::string msg;
        int rc = 1;
        int32_t sc = 0;
        if (e) {
            msg = "MVTools_vectors property missing";
        } else {
            std::lock_guard<std::mutex> lock(g_engine_mutex);
            mvb200_ctx *ctx = engine(msg);
            rc = ctx ? mvb200_scdetection_frame(ctx, &d->p, vec, &sc) : 1;
            if (ctx && rc)
                msg = mvb200_last_error();
        }
        vsapi->freeFrame(mvn);
        if (rc) {
            vsapi->setFilterError(("SCDetection: " + msg).c_str(), frameCtx);
            vsapi->freeFrame(dst);
            return nullptr;
        }
        vsapi->mapSetInt(vsapi->getFramePropertiesRW(dst), d->p.vd.isBackward ? "_SceneChangeNext" : "_SceneChangePrev", sc, maReplace);
        return dst;
    }
    return nullptr;
}

void VS_CC scdFree(void *instanceData, VSCore *, const VSAPI *vsapi) {
    ScdData *d = (ScdData *)instanceData;
    vsapi->freeNode(d->node);
    vsapi->freeNode(d->vectors);
    delete d;
}

void VS_CC scdCreate(const VSMap *in, VSMap *out, void *, VSCore *core, const VSAPI *vsapi) {
    ScdData d;
    memset(&d, 0, sizeof(d));
    d.vectors = vsapi->mapGetNode(in, "vectors", 0, nullptr);
    std::string msg;
    mvb200_analysis_data vd;
    if (!adataFromClip(d.vectors, vsapi, "SCDetection", "vectors", &vd, msg) || mvb200_scdetection_init(argInt64(vsapi, in, "thscd1"), argInt(vsapi, in, "thscd2"), &vd, &d.p)) {
        vsapi->mapSetError(out, msg.empty() ? mvb200_last_error() : msg.c_str());
        vsapi->freeNode(d.vectors);
        return;
    }
    d.node = vsapi->mapGetNode(in, "clip", 0, nullptr);
    d.vi = vsapi->getVideoInfo(d.node);
    ScdData *data = new ScdData(d);
    VSFilterDependency deps[2] = {{data->node, rpStrictSpatial}, {data->vectors, rpStrictSpatial}};
    vsapi->createVideoFilter(out, "SCDetection", data->vi, scdGetFrame, scdFree, fmParallel, deps, 2, data, core);
}

// ------------------------------------------------------------------- Mask
// mvmaskCreate / GetFrame, MVMask.c:74-344
struct MaskData {
    VSNode *node, *vectors;
    VSVideoInfo vi;
    mvb200_mask_params p;
};

const VSFrame *VS_CC maskGetFrame(int n, int activationReason, void *instanceData, void **, VSFrameContext *frameCtx, VSCore *core, const VSAPI *vsapi) {
    MaskData *d = (MaskData *)instanceData;
    if (activationReason == arInitial) {
        vsapi->requestFrameFilter(n, d->vectors, frameCtx);
        vsapi->requestFrameFilter(n, d->node, frameCtx);
    } else if (activationReason == arAllFramesReady) {
        const VSFrame *src = vsapi->getFrameFilter(n, d->node, frameCtx);
        VSFrame *dst = vsapi->newVideoFrame(&d->vi.format, d->vi.width, d->vi.height, src, core);
        const VSFrame *mvn = vsapi->getFrameFilter(n, d->vectors, frameCtx);
        int e = 0;
        const uint8_t *vec = vectorBlob(vsapi, mvn, &e);
        std::string msg;
        int rc = 1;
        if (e) {
            msg = "MVTools_vectors property missing";
        } else if (d->vi.width != d->p.nWidth || d->vi.height != d->p.nHeight || (d->vi.width >> d->vi.format.subSamplingW) != d->p.nWidthUV || (d->vi.height >> d->vi.format.subSamplingH) != d->p.nHeightUV) {
            msg = "the clip's dimensions or subsampling differ from the vectors' (the reference would write out of bounds)";
        } else {
            uint8_t *dp_[3];
            ptrdiff_t ds[3];
            planesRW(vsapi, dst, 3, dp_, ds);
            std::lock_guard<std::mutex> lock(g_engine_mutex);
            mvb200_ctx *ctx = engine(msg);
            rc = ctx ? mvb200_mask_frame(ctx, &d->p, vsapi->getReadPtr(src, 0), vsapi->getStride(src, 0), vec, dp_, ds) : 1;
            if (ctx && rc)
                msg = mvb200_last_error();
        }
        vsapi->freeFrame(mvn);
        vsapi->freeFrame(src);
        if (rc) {
            vsapi->setFilterError(("Mask: " + msg).c_str(), frameCtx);
            vsapi->freeFrame(dst);
            return nullptr;
        }
        return dst;
    }
    return nullptr;
}

void VS_CC maskFree(void *instanceData, VSCore *, const VSAPI *vsapi) {
    MaskData *d = (MaskData *)instanceData;
    vsapi->freeNode(d->node);
    vsapi->freeNode(d->vectors);
    delete d;
}

double argFloat(const VSAPI *vsapi, const VSMap *in, const char *name) {
    int err = 0;
    double v = vsapi->mapGetFloat(in, name, 0, &err);
    return err ? NAN : v;
}

void VS_CC maskCreate(const VSMap *in, VSMap *out, void *, VSCore *core, const VSAPI *vsapi) {
    mvb200_mask_args a;
    a.ml = argFloat(vsapi, in, "ml");
    a.gamma = argFloat(vsapi, in, "gamma");
    a.kind = argInt(vsapi, in, "kind");
    a.time = argFloat(vsapi, in, "time");
    a.ysc = argInt(vsapi, in, "ysc");
    a.thscd1 = argInt64(vsapi, in, "thscd1");
    a.thscd2 = argInt(vsapi, in, "thscd2");
    MaskData d;
    memset(&d, 0, sizeof(d));
    // the argument checks come before any clip is touched (MVMask.c:261-279): run them on a neutral vector description
    {
        mvb200_analysis_data probe;
        memset(&probe, 0, sizeof(probe));
        probe.nBlkSizeX = probe.nBlkSizeY = 8;
        probe.nBlkX = probe.nBlkY = 2;
        probe.xRatioUV = probe.yRatioUV = 1;
        probe.bitsPerSample = 8;
        mvb200_mask_args b = a;
        b.thscd1 = MVB200_UNSET64;
        if (mvb200_mask_init(&b, &probe, 8, 0, 0, 0, &d.p)) {
            vsapi->mapSetError(out, mvb200_last_error());
            return;
        }
    }
    d.vectors = vsapi->mapGetNode(in, "vectors", 0, nullptr);
    std::string msg;
    mvb200_analysis_data vd;
    if (!adataFromClip(d.vectors, vsapi, "Mask", "vectors", &vd, msg)) {
        vsapi->mapSetError(out, msg.c_str());
        vsapi->freeNode(d.vectors);
        return;
    }
    {   // scaleThSCD's error precedes the format check (MVMask.c:288-297)
        mvb200_scdetection_params t;
        if (mvb200_scdetection_init(a.thscd1, a.thscd2, &vd, &t)) {
            std::string m = mvb200_last_error();
            m.replace(0, strlen("SCDetection"), "Mask");
            vsapi->mapSetError(out, m.c_str());
            vsapi->freeNode(d.vectors);
            return;
        }
    }
    d.node = vsapi->mapGetNode(in, "clip", 0, nullptr);
    d.vi = *vsapi->getVideoInfo(d.node);
    const bool fmtOk = vsh::isConstantVideoFormat(&d.vi) && d.vi.format.bitsPerSample <= 8 && d.vi.format.subSamplingW <= 1 && d.vi.format.subSamplingH <= 1 &&
                       (d.vi.format.colorFamily == cfYUV || d.vi.format.colorFamily == cfGray);
    if (!fmtOk || mvb200_mask_init(&a, &vd, d.vi.format.bitsPerSample, d.vi.format.subSamplingW, d.vi.format.subSamplingH, d.vi.format.colorFamily == cfGray, &d.p)) {
        vsapi->mapSetError(out, fmtOk ? mvb200_last_error() : "Mask: input clip must be GRAY8, YUV420P8, YUV422P8, YUV440P8, or YUV444P8, with constant dimensions.");
        vsapi->freeNode(d.node);
        vsapi->freeNode(d.vectors);
        return;
    }
    if (d.vi.format.colorFamily == cfGray)
        vsapi->getVideoFormatByID(&d.vi.format, pfYUV444P8, core);
    MaskData *data = new MaskData(d);
    VSFilterDependency deps[2] = {{data->node, rpStrictSpatial}, {data->vectors, rpStrictSpatial}};
    vsapi->createVideoFilter(out, "Mask", &data->vi, maskGetFrame, maskFree, fmParallel, deps, 2, data, core);
}

// ----------------------------------------------------------------- Finest
// mvfinestCreate / GetFrame, MVFinest.c:47-207
struct FinestData {
    VSNode *super;
    VSVideoInfo vi;
    mvb200_super_params sp;
};

const VSFrame *VS_CC finestGetFrame(int n, int activationReason, void *instanceData, void **, VSFrameContext *frameCtx, VSCore *core, const VSAPI *vsapi) {
    FinestData *d = (FinestData *)instanceData;
    if (activationReason == arInitial) {
        vsapi->requestFrameFilter(n, d->super, frameCtx);
    } else if (activationReason == arAllFramesReady) {
        const VSFrame *ref = vsapi->getFrameFilter(n, d->super, frameCtx);
        VSFrame *dst = vsapi->newVideoFrame(&d->vi.format, d->vi.width, d->vi.height, ref, core);
        const uint8_t *sp_[3];
        ptrdiff_t ss[3];
        uint8_t *dp_[3];
        ptrdiff_t ds[3];
        planesRO(vsapi, ref, d->vi.format.numPlanes, sp_, ss);
        planesRW(vsapi, dst, d->vi.format.numPlanes, dp_, ds);
        std::string msg;
        int rc = 0;
        if (d->sp.pel == 1) { // :75-77, a plain copy of level 0 (each plane's own size: the reference copies the luma size for every plane)
            for (int i = 0; i < d->vi.format.numPlanes; i++)
                vsh::bitblt(dp_[i], ds[i], sp_[i], ss[i], (size_t)(i ? d->vi.width >> d->vi.format.subSamplingW : d->vi.width) * d->vi.format.bytesPerSample,
                            (size_t)(i ? d->vi.height >> d->vi.format.subSamplingH : d->vi.height));
        } else {
            std::lock_guard<std::mutex> lock(g_engine_mutex);
            mvb200_ctx *ctx = engine(msg);
            rc = ctx ? mvb200_finest_frame(ctx, &d->sp, sp_, ss, dp_, ds) : 1;
            if (ctx && rc)
                msg = mvb200_last_error();
        }
        vsapi->freeFrame(ref);
        if (rc) {
            vsapi->setFilterError(("Finest: " + msg).c_str(), frameCtx);
            vsapi->freeFrame(dst);
            return nullptr;
        }
        return dst;
    }
    return nullptr;
}

void VS_CC finestFree(void *instanceData, VSCore *, const VSAPI *vsapi) {
    FinestData *d = (FinestData *)instanceData;
    vsapi->freeNode(d->super);
    delete d;
}

void VS_CC finestCreate(const VSMap *in, VSMap *out, void *, VSCore *core, const VSAPI *vsapi) {
    FinestData d;
    memset(&d, 0, sizeof(d));
    d.super = vsapi->mapGetNode(in, "super", 0, nullptr);
    d.vi = *vsapi->getVideoInfo(d.super);
    if (!formatOk(&d.vi)) {
        vsapi->mapSetError(out, "Finest: input clip must be GRAY, 420, 422, 440, or 444, up to 16 bits, with constant dimensions.");
        vsapi->freeNode(d.super);
        return;
    }
    std::string msg;
    if (!superFromClip(d.super, vsapi, "Finest", &d.sp, msg)) {
        vsapi->mapSetError(out, msg.c_str());
        vsapi->freeNode(d.super);
        return;
    }
    d.vi.width = (d.sp.width + 2 * d.sp.hpad) * d.sp.pel; // :190-194
    d.vi.height = (d.sp.height + 2 * d.sp.vpad) * d.sp.pel;
    FinestData *data = new FinestData(d);
    VSFilterDependency deps[1] = {{data->super, rpStrictSpatial}};
    vsapi->createVideoFilter(out, "Finest", &data->vi, finestGetFrame, finestFree, fmParallel, deps, 1, data, core);
}

// ------------------------------------------------- fused paths of the per-pixel consumers
// A fused frame that the reference would have made from a SUPER frame carries that frame's properties: those of source
// frame `from` (read here; requested by fusedGetFrame as alsoFrame / propFrame) plus Super_* when it is frame 0.
void superFrameProps(const VSAPI *vsapi, const VSFrame *out, VSNode *clip, int from, const mvb200_super_params &sp, VSFrameContext *frameCtx) {
    VSFrame *w = const_cast<VSFrame *>(out); // (just created by fusedGetFrame, not shared yet)
    const VSFrame *ref = vsapi->getFrameFilter(from, clip, frameCtx);
    VSMap *props = vsapi->getFramePropertiesRW(w);
    vsapi->clearMap(props);
    vsapi->copyMap(vsapi->getFramePropertiesRO(ref), props);
    vsapi->freeFrame(ref);
    if (from == 0)
        setSuperProps(vsapi, w, sp);
}

// ------------------------------------------------------------------- Flow
// mvflowCreate / GetFrame, MVFlow.cpp:157-593.  The reference reads an mv.Finest clip made from `super`; the engine reads
// the pel sub-planes of the super frame itself (same samples), so this filter requests super frames.
// Fused path (see fusedGetFrame): a window of B frames is one mvb200_flow_clip_host call over source frames.
struct FlowFused : FusedCache {
    mvb200_analyse_params ap;
};

struct FlowData {
    VSNode *clip, *super, *vectors;
    const VSVideoInfo *vi, *supervi;
    mvb200_super_params sp;
    mvb200_flow_params fp;
    FlowFused *fused; // owned; null = node-by-node evaluation
};

const VSFrame *VS_CC flowFusedGetFrame(int n, int activationReason, void *instanceData, void **frameData, VSFrameContext *frameCtx, VSCore *core, const VSAPI *vsapi) {
    FlowData *d = (FlowData *)instanceData;
    FlowFused *f = d->fused;
    const int off = d->fp.vd.nDeltaFrame, N = d->vi->numFrames;
    const int nref = d->fp.vd.isBackward ? n + off : n - off;
    const bool haveRef = nref >= 0 && nref < N;
    auto segment = [&](int w0, int w1, int *s0, int *s1) {
        *s0 = w0 - off > 0 ? w0 - off : 0;
        *s1 = w1 + off < N ? w1 + off : N;
    };
    auto run = [&](mvb200_ctx *ctx, const void *hostSrc, int s0, int s1, int w0, int w1, void *hostDst, int32_t *flags) {
        return mvb200_flow_clip_host(ctx, &d->sp, &f->ap, &d->fp, hostSrc, s0, s1 - s0, w0, w1 - w0, hostDst, flags);
    };
    int32_t usable = 0;
    const VSFrame *out = fusedGetFrame("Flow", f, d->clip, d->vi, n, n, 7u, activationReason, frameData, frameCtx, core, vsapi, segment, run, haveRef ? nref : -1, &usable);
    if (out && usable && haveRef) // the reference builds a compensated frame on the reference frame's Finest frame (MVFlow.cpp:196-204)
        superFrameProps(vsapi, out, d->clip, nref, d->sp, frameCtx);
    return out;
}

const VSFrame *VS_CC flowGetFrame(int n, int activationReason, void *instanceData, void **, VSFrameContext *frameCtx, VSCore *core, const VSAPI *vsapi) {
    FlowData *d = (FlowData *)instanceData;
    const int offset = d->fp.vd.nDeltaFrame;
    const int nref = offset > 0 ? (d->fp.vd.isBackward ? n + offset : n - offset) : -offset; // :163-170
    const bool haveRef = nref >= 0 && nref < d->vi->numFrames;
    const bool needField = d->fp.fields && d->fp.vd.nPel > 1 && (nref - n) % 2 != 0;
    if (activationReason == arInitial) {
        vsapi->requestFrameFilter(n, d->vectors, frameCtx);
        if (haveRef) {
            vsapi->requestFrameFilter(nref, d->super, frameCtx);
            if (needField)
                vsapi->requestFrameFilter(n, d->super, frameCtx);
        }
        vsapi->requestFrameFilter(n, d->clip, frameCtx);
    } else if (activationReason == arAllFramesReady) {
        const VSFrame *mvn = vsapi->getFrameFilter(n, d->vectors, frameCtx);
        int e = 0;
        const uint8_t *vec = vectorBlob(vsapi, mvn, &e);
        const VSFrame *ref = haveRef ? vsapi->getFrameFilter(nref, d->super, frameCtx) : nullptr;
        VSFrame *dst = ref ? vsapi->newVideoFrame(&d->vi->format, d->vi->width, d->vi->height, ref, core) : nullptr;
        std::string msg;
        int rc = 1;
        int32_t usable = 0;
        int fieldShift = 0;
        bool fieldOk = true;
        if (ref && needField) {
            const VSFrame *srcSuper = vsapi->getFrameFilter(n, d->super, frameCtx);
            fieldOk = fieldShiftOf(vsapi, srcSuper, ref, n, nref, true, d->fp.tff_exists, d->fp.tff, d->fp.vd.nPel, true, &fieldShift);
            vsapi->freeFrame(srcSuper);
        }
        if (e) {
            msg = "MVTools_vectors property missing";
        } else if (!fieldOk) {
            msg = "_Field property not found in super frame. Therefore, you must pass tff argument.";
        } else if (!ref) {
            rc = 0; // out of range: the vectors are invalid, the source frame is returned
        } else {
            const uint8_t *rp_[3];
            ptrdiff_t rs[3];
            uint8_t *dp_[3];
            ptrdiff_t ds[3];
            planesRO(vsapi, ref, d->supervi->format.numPlanes, rp_, rs);
            planesRW(vsapi, dst, d->vi->format.numPlanes, dp_, ds);
            std::lock_guard<std::mutex> lock(g_engine_mutex);
            mvb200_ctx *ctx = engine(msg);
            rc = ctx ? mvb200_flow_frame(ctx, &d->fp, &d->sp, rp_, rs, vec, fieldShift, dp_, ds, &usable) : 1;
            if (ctx && rc)
                msg = mvb200_last_error();
        }
        vsapi->freeFrame(mvn);
        vsapi->freeFrame(ref);
        if (rc) {
            vsapi->setFilterError(("Flow: " + msg).c_str(), frameCtx);
            vsapi->freeFrame(dst);
            return nullptr;
        }
        if (!usable) { // :362-366
            vsapi->freeFrame(dst);
            return vsapi->getFrameFilter(n, d->clip, frameCtx);
        }
        return dst;
    }
    return nullptr;
}

void VS_CC flowFree(void *instanceData, VSCore *, const VSAPI *vsapi) {
    FlowData *d = (FlowData *)instanceData;
    vsapi->freeNode(d->clip);
    vsapi->freeNode(d->super);
    vsapi->freeNode(d->vectors);
    delete d->fused;
    delete d;
}

void VS_CC flowCreate(const VSMap *in, VSMap *out, void *, VSCore *core, const VSAPI *vsapi) {
    mvb200_flow_args a;
    a.time = argFloat(vsapi, in, "time");
    a.mode = argInt(vsapi, in, "mode");
    a.fields = argInt(vsapi, in, "fields");
    a.thscd1 = argInt64(vsapi, in, "thscd1");
    a.thscd2 = argInt(vsapi, in, "thscd2");
    a.tff = argInt(vsapi, in, "tff");
    // argument checks before any clip is touched (MVFlow.cpp:427-436)
    if (a.time == a.time && (a.time < 0.0 || a.time > 100.0)) {
        vsapi->mapSetError(out, "Flow: time must be between 0 and 100 % (inclusive).");
        return;
    }
    if (a.mode != MVB200_UNSET && (a.mode < 0 || a.mode > 1)) {
        vsapi->mapSetError(out, "Flow: mode must be 0 or 1.");
        return;
    }
    FlowData d;
    memset(&d, 0, sizeof(d));
    auto freeAll = [&]() {
        vsapi->freeNode(d.clip);
        vsapi->freeNode(d.super);
        vsapi->freeNode(d.vectors);
    };
    d.super = vsapi->mapGetNode(in, "super", 0, nullptr);
    d.supervi = vsapi->getVideoInfo(d.super);
    std::string msg;
    mvb200_analysis_data vd;
    if (!superFromClip(d.super, vsapi, "Flow", &d.sp, msg)) {
        vsapi->mapSetError(out, msg.c_str());
        freeAll();
        return;
    }
    d.vectors = vsapi->mapGetNode(in, "vectors", 0, nullptr);
    if (!adataFromClip(d.vectors, vsapi, "Flow", "vectors", &vd, msg)) {
        vsapi->mapSetError(out, msg.c_str());
        freeAll();
        return;
    }
    d.clip = vsapi->mapGetNode(in, "clip", 0, nullptr);
    d.vi = vsapi->getVideoInfo(d.clip);
    if (mvb200_flow_init(&a, &d.sp, &vd, d.vi->numFrames, &d.fp)) {
        vsapi->mapSetError(out, mvb200_last_error());
        freeAll();
        return;
    }
    if (!formatOk(d.vi)) {
        vsapi->mapSetError(out, "Flow: input clip must be GRAY, 420, 422, 440, or 444, up to 16 bits, with constant dimensions.");
        freeAll();
        return;
    }
    if (d.vi->width != vd.nWidth || d.vi->height != vd.nHeight || d.vi->format.bitsPerSample != d.sp.bits) {
        vsapi->mapSetError(out, "Flow: wrong source or super clip frame size.");
        freeAll();
        return;
    }
    const int window = lookaheadFromEnv();
    mvb200_super_params ssp;
    mvb200_analyse_params sap;
    if (window && !d.fp.fields && canonicalChain(d.clip, d.super, &d.vectors, 1, d.sp, d.vi->numFrames, true, &ssp, &sap)) {
        d.fused = new FlowFused;
        d.fused->ap = sap;
        d.sp = ssp;
        mvb200_clip_layout(&d.sp, &d.fused->cl);
        d.fused->window = fusedWindowFrames(d.fused, window, &sap, &d.sp, 1);
    }
    FlowData *data = new FlowData(d);
    if (data->fused) { // only the source clip is read on this path
        VSFilterDependency fdeps[1] = {{data->clip, rpGeneral}};
        vsapi->createVideoFilter(out, "Flow", data->vi, flowFusedGetFrame, flowFree, fmParallel, fdeps, 1, data, core);
        return;
    }
    VSFilterDependency deps[3] = {{data->clip, rpStrictSpatial}, {data->super, rpGeneral}, {data->vectors, rpStrictSpatial}};
    vsapi->createVideoFilter(out, "Flow", data->vi, flowGetFrame, flowFree, fmParallel, deps, 3, data, core);
}

// -------------------------------------------------------------- FlowInter
// mvflowinterCreate / GetFrame, MVFlowInter.c:83-713 (super frames are read where the reference reads its mv.Finest clip)
// Fused path (see fusedGetFrame): a window of B frames is one mvb200_flowinter_clip_host call over source frames (the
// window, delta frames before it and 2 x delta after it for the "extra" vectors).
struct FlowInterFused : FusedCache {
    mvb200_analyse_params apBw, apFw;
};

struct FlowInterData {
    VSNode *node, *super, *mvbw, *mvfw;
    const VSVideoInfo *vi, *supervi;
    mvb200_super_params sp, clipsp;
    mvb200_flowinter_params p;
    FlowInterFused *fused; // owned; null = node-by-node evaluation
};

const VSFrame *VS_CC flowInterFusedGetFrame(int n, int activationReason, void *instanceData, void **frameData, VSFrameContext *frameCtx, VSCore *core, const VSAPI *vsapi) {
    FlowInterData *d = (FlowInterData *)instanceData;
    FlowInterFused *f = d->fused;
    const int off = d->p.vd.nDeltaFrame, N = d->vi->numFrames;
    auto segment = [&](int w0, int w1, int *s0, int *s1) {
        *s0 = w0 - off > 0 ? w0 - off : 0;
        *s1 = w1 + 2 * off < N ? w1 + 2 * off : N;
    };
    auto run = [&](mvb200_ctx *ctx, const void *hostSrc, int s0, int s1, int w0, int w1, void *hostDst, int32_t *flags) {
        return mvb200_flowinter_clip_host(ctx, &d->sp, &f->apBw, &f->apFw, &d->p, hostSrc, s0, s1 - s0, w0, w1 - w0, hostDst, flags);
    };
    int32_t usable = 0;
    const VSFrame *out = fusedGetFrame("FlowInter", f, d->node, d->vi, n, n, 7u, activationReason, frameData, frameCtx, core, vsapi, segment, run, -1, &usable);
    if (out && usable && n == 0) // an interpolated frame is built on super frame n (MVFlowInter.c:131): Super_* on frame 0
        setSuperProps(vsapi, const_cast<VSFrame *>(out), d->sp);
    return out;
}

const uint8_t *vectorsOf(const VSAPI *vsapi, const VSFrame *f, bool *missing) {
    int e = 0;
    const uint8_t *v = vectorBlob(vsapi, f, &e);
    if (e)
        *missing = true;
    return e ? nullptr : v;
}

const VSFrame *VS_CC flowInterGetFrame(int n, int activationReason, void *instanceData, void **, VSFrameContext *frameCtx, VSCore *core, const VSAPI *vsapi) {
    FlowInterData *d = (FlowInterData *)instanceData;
    const int off = d->p.vd.nDeltaFrame, N = d->vi->numFrames;
    const bool inRange = n + off < N;
    if (activationReason == arInitial) { // :89-105
        if (inRange) {
            vsapi->requestFrameFilter(n, d->mvfw, frameCtx);
            vsapi->requestFrameFilter(n + off, d->mvfw, frameCtx);
            vsapi->requestFrameFilter(n, d->mvbw, frameCtx);
            vsapi->requestFrameFilter(n + off, d->mvbw, frameCtx);
            vsapi->requestFrameFilter(n, d->super, frameCtx);
            vsapi->requestFrameFilter(n + off, d->super, frameCtx);
        }
        vsapi->requestFrameFilter(n, d->node, frameCtx);
        vsapi->requestFrameFilter(VSMIN(n + off, N - 1), d->node, frameCtx);
    } else if (activationReason == arAllFramesReady) {
        std::string msg;
        int rc = 0;
        int32_t usable = 0;
        VSFrame *dst = nullptr;
        if (inRange) {
            const VSFrame *mvF = vsapi->getFrameFilter(n + off, d->mvfw, frameCtx), *mvB = vsapi->getFrameFilter(n, d->mvbw, frameCtx);
            const VSFrame *mvFF = vsapi->getFrameFilter(n, d->mvfw, frameCtx), *mvBB = vsapi->getFrameFilter(n + off, d->mvbw, frameCtx);
            const VSFrame *left = vsapi->getFrameFilter(n, d->super, frameCtx), *right = vsapi->getFrameFilter(n + off, d->super, frameCtx);
            bool missing = false;
            const uint8_t *vb = vectorsOf(vsapi, mvB, &missing), *vf = vectorsOf(vsapi, mvF, &missing), *vbb = vectorsOf(vsapi, mvBB, &missing), *vff = vectorsOf(vsapi, mvFF, &missing);
            dst = vsapi->newVideoFrame(&d->vi->format, d->vi->width, d->vi->height, left, core);
            if (missing) {
                rc = 1;
                msg = "MVTools_vectors property missing";
            } else {
                const uint8_t *lp[3], *rp[3];
                ptrdiff_t ls[3], rs[3];
                uint8_t *dp_[3];
                ptrdiff_t ds[3];
                planesRO(vsapi, left, d->supervi->format.numPlanes, lp, ls);
                planesRO(vsapi, right, d->supervi->format.numPlanes, rp, rs);
                planesRW(vsapi, dst, d->vi->format.numPlanes, dp_, ds);
                std::lock_guard<std::mutex> lock(g_engine_mutex);
                mvb200_ctx *ctx = engine(msg);
                rc = ctx ? mvb200_flowinter_frame(ctx, &d->p, &d->sp, d->p.time256, lp, ls, rp, rs, vb, vf, vbb, vff, dp_, ds, &usable) : 1;
                if (ctx && rc)
                    msg = mvb200_last_error();
            }
            vsapi->freeFrame(mvF);
            vsapi->freeFrame(mvB);
            vsapi->freeFrame(mvFF);
            vsapi->freeFrame(mvBB);
            vsapi->freeFrame(left);
            vsapi->freeFrame(right);
        }
        if (!rc && !usable) { // poor estimation, :412-447
            vsapi->freeFrame(dst);
            dst = nullptr;
            const VSFrame *src = vsapi->getFrameFilter(n, d->node, frameCtx);
            if (!d->p.blend)
                return src; // like ChangeFPS
            const VSFrame *ref = vsapi->getFrameFilter(VSMIN(n + off, N - 1), d->node, frameCtx);
            dst = vsapi->newVideoFrame(&d->vi->format, d->vi->width, d->vi->height, src, core);
            const uint8_t *sp_[3], *rp[3];
            ptrdiff_t ss[3], rs[3];
            uint8_t *dp_[3];
            ptrdiff_t ds[3];
            planesRO(vsapi, src, d->vi->format.numPlanes, sp_, ss);
            planesRO(vsapi, ref, d->vi->format.numPlanes, rp, rs);
            planesRW(vsapi, dst, d->vi->format.numPlanes, dp_, ds);
            {
                std::lock_guard<std::mutex> lock(g_engine_mutex);
                mvb200_ctx *ctx = engine(msg);
                rc = ctx ? mvb200_blend_frame(ctx, &d->clipsp, d->p.time256, sp_, ss, rp, rs, dp_, ds) : 1;
                if (ctx && rc)
                    msg = mvb200_last_error();
            }
            vsapi->freeFrame(src);
            vsapi->freeFrame(ref);
        }
        if (rc) {
            vsapi->setFilterError(("FlowInter: " + msg).c_str(), frameCtx);
            vsapi->freeFrame(dst);
            return nullptr;
        }
        return dst;
    }
    return nullptr;
}

void VS_CC flowInterFree(void *instanceData, VSCore *, const VSAPI *vsapi) {
    FlowInterData *d = (FlowInterData *)instanceData;
    vsapi->freeNode(d->node);
    vsapi->freeNode(d->super);
    vsapi->freeNode(d->mvbw);
    vsapi->freeNode(d->mvfw);
    delete d->fused;
    delete d;
}

void VS_CC flowInterCreate(const VSMap *in, VSMap *out, void *, VSCore *core, const VSAPI *vsapi) {
    mvb200_flowinter_args a;
    a.time = argFloat(vsapi, in, "time");
    a.ml = argFloat(vsapi, in, "ml");
    a.blend = argInt(vsapi, in, "blend");
    a.thscd1 = argInt64(vsapi, in, "thscd1");
    a.thscd2 = argInt(vsapi, in, "thscd2");
    // argument checks before any clip is touched (MVFlowInter.c:493-502; float comparisons as there)
    if (a.time == a.time && ((float)a.time < 0.0f || (float)a.time > 100.0f)) {
        vsapi->mapSetError(out, "FlowInter: time must be between 0 and 100 % (inclusive).");
        return;
    }
    if (a.ml == a.ml && (float)a.ml <= 0.0f) {
        vsapi->mapSetError(out, "FlowInter: ml must be greater than 0.");
        return;
    }
    FlowInterData d;
    memset(&d, 0, sizeof(d));
    auto freeAll = [&]() {
        vsapi->freeNode(d.node);
        vsapi->freeNode(d.super);
        vsapi->freeNode(d.mvbw);
        vsapi->freeNode(d.mvfw);
    };
    d.super = vsapi->mapGetNode(in, "super", 0, nullptr);
    d.supervi = vsapi->getVideoInfo(d.super);
    std::string msg;
    mvb200_analysis_data vdb, vdf;
    if (!superFromClip(d.super, vsapi, "FlowInter", &d.sp, msg)) {
        vsapi->mapSetError(out, msg.c_str());
        freeAll();
        return;
    }
    d.mvbw = vsapi->mapGetNode(in, "mvbw", 0, nullptr);
    d.mvfw = vsapi->mapGetNode(in, "mvfw", 0, nullptr);
    if (!adataFromClip(d.mvbw, vsapi, "FlowInter", "mvbw", &vdb, msg) || !adataFromClip(d.mvfw, vsapi, "FlowInter", "mvfw", &vdf, msg)) {
        vsapi->mapSetError(out, msg.c_str());
        freeAll();
        return;
    }
    d.node = vsapi->mapGetNode(in, "clip", 0, nullptr);
    d.vi = vsapi->getVideoInfo(d.node);
    if (mvb200_flowinter_init(&a, &d.sp, &vdb, &vdf, d.vi->numFrames, &d.p)) {
        vsapi->mapSetError(out, mvb200_last_error());
        freeAll();
        return;
    }
    if (!formatOk(d.vi)) {
        vsapi->mapSetError(out, "FlowInter: input clip must be GRAY, 420, 422, 440, or 444, up to 16 bits, with constant dimensions.");
        freeAll();
        return;
    }
    if (d.vi->width != vdb.nWidth || d.vi->height != vdb.nHeight || d.vi->format.bitsPerSample != d.sp.bits) {
        vsapi->mapSetError(out, "FlowInter: wrong source or super clip frame size.");
        freeAll();
        return;
    }
    // the source clip described as the engine describes clips (for the blend of unusable frames)
    memset(&d.clipsp, 0, sizeof(d.clipsp));
    d.clipsp = d.sp;
    const int window = lookaheadFromEnv();
    mvb200_super_params ssp;
    mvb200_analyse_params sap[2];
    VSNode *vecs[2] = {d.mvbw, d.mvfw};
    if (window && canonicalChain(d.node, d.super, vecs, 2, d.sp, d.vi->numFrames, true, &ssp, sap)) {
        d.fused = new FlowInterFused;
        d.fused->apBw = sap[0];
        d.fused->apFw = sap[1];
        d.sp = ssp;
        mvb200_clip_layout(&d.sp, &d.fused->cl);
        d.fused->window = fusedWindowFrames(d.fused, window, sap, &d.sp, 2);
    }
    FlowInterData *data = new FlowInterData(d);
    VSFilterDependency deps[4] = {{data->node, rpGeneral}, {data->super, rpGeneral}, {data->mvbw, rpGeneral}, {data->mvfw, rpGeneral}};
    if (data->fused) // only the source clip is read on this path
        vsapi->createVideoFilter(out, "FlowInter", data->vi, flowInterFusedGetFrame, flowInterFree, fmParallel, deps, 1, data, core);
    else
        vsapi->createVideoFilter(out, "FlowInter", data->vi, flowInterGetFrame, flowInterFree, fmParallel, deps, 4, data, core);
}

// ---------------------------------------------------------------- FlowFPS
// mvflowfpsCreate / GetFrame, MVFlowFPS.c:86-928.  The reference wraps the vector clips in a helper filter that
// attaches the upsized vector planes as frame properties; the engine upsizes inside the kernel, so the vector clips
// are used as they are.
// Fused path (see fusedGetFrame): a window of B OUTPUT frames is one mvb200_flowfps_clip_host call over the source
// frames the window needs (its left / right frames, delta before and 2 x delta after them for the "extra" vectors).
struct FlowFPSFused : FusedCache {
    mvb200_analyse_params apBw, apFw;
};

struct FlowFPSData {
    VSNode *node, *super, *mvbw, *mvfw;
    VSVideoInfo vi;
    const VSVideoInfo *oldvi, *supervi;
    mvb200_super_params sp;
    mvb200_flowfps_params p;
    FlowFPSFused *fused; // owned; null = node-by-node evaluation
};

const VSFrame *VS_CC flowFpsFusedGetFrame(int n, int activationReason, void *instanceData, void **frameData, VSFrameContext *frameCtx, VSCore *core, const VSAPI *vsapi) {
    FlowFPSData *d = (FlowFPSData *)instanceData;
    FlowFPSFused *f = d->fused;
    const int off = d->p.inter.vd.nDeltaFrame, N = d->oldvi->numFrames;
    int32_t nleft, time256;
    mvb200_flowfps_position(&d->p, n, &nleft, &time256);
    const int propFrame = VSMIN(time256 == 256 ? nleft + off : nleft, N - 1); // whose properties the output carries
    if (time256 == 0 || time256 == 256) { // the input frame itself (MVFlowFPS.c:93-101,137-141)
        if (activationReason == arInitial)
            vsapi->requestFrameFilter(propFrame, d->node, frameCtx);
        return activationReason == arAllFramesReady ? vsapi->getFrameFilter(propFrame, d->node, frameCtx) : nullptr;
    }
    auto segment = [&](int w0, int w1, int *s0, int *s1) {
        int32_t lmin, lmax, t;
        mvb200_flowfps_position(&d->p, w0, &lmin, &t);
        mvb200_flowfps_position(&d->p, w1 - 1, &lmax, &t);
        *s0 = lmin - off > 0 ? lmin - off : 0;
        *s1 = lmax + 2 * off + 1 < N ? lmax + 2 * off + 1 : N;
        if (*s0 > N - 1)
            *s0 = N - 1; // (output frames past the end are made of the last input frame)
    };
    auto run = [&](mvb200_ctx *ctx, const void *hostSrc, int s0, int s1, int w0, int w1, void *hostDst, int32_t *flags) {
        return mvb200_flowfps_clip_host(ctx, &d->sp, &f->apBw, &f->apFw, &d->p, hostSrc, s0, s1 - s0, w0, w1 - w0, hostDst, flags);
    };
    int32_t usable = 0;
    const VSFrame *out = fusedGetFrame("FlowFPS", f, d->node, &d->vi, n, propFrame, 7u, activationReason, frameData, frameCtx, core, vsapi, segment, run, -1, &usable);
    if (out && usable && nleft == 0) // an interpolated frame is built on the left super frame (MVFlowFPS.c:162): Super_* on frame 0
        setSuperProps(vsapi, const_cast<VSFrame *>(out), d->sp);
    return out;
}

const VSFrame *VS_CC flowFpsGetFrame(int n, int activationReason, void *instanceData, void **, VSFrameContext *frameCtx, VSCore *core, const VSAPI *vsapi) {
    FlowFPSData *d = (FlowFPSData *)instanceData;
    const int off = d->p.inter.vd.nDeltaFrame, N = d->oldvi->numFrames;
    int32_t nleft, time256;
    mvb200_flowfps_position(&d->p, n, &nleft, &time256);
    const int nright = nleft + off;
    const bool inRange = nleft < N && nright < N, extra = d->p.maskmode == 2;
    if (activationReason == arInitial) { // :93-123
        if (time256 == 0) {
            vsapi->requestFrameFilter(VSMIN(nleft, N - 1), d->node, frameCtx);
            return nullptr;
        } else if (time256 == 256) {
            vsapi->requestFrameFilter(VSMIN(nright, N - 1), d->node, frameCtx);
            return nullptr;
        }
        if (inRange) {
            if (extra)
                vsapi->requestFrameFilter(nleft, d->mvfw, frameCtx);
            vsapi->requestFrameFilter(nright, d->mvfw, frameCtx);
            vsapi->requestFrameFilter(nleft, d->mvbw, frameCtx);
            if (extra)
                vsapi->requestFrameFilter(nright, d->mvbw, frameCtx);
            vsapi->requestFrameFilter(nleft, d->super, frameCtx);
            vsapi->requestFrameFilter(nright, d->super, frameCtx);
        }
        vsapi->requestFrameFilter(VSMIN(nleft, N - 1), d->node, frameCtx);
        if (d->p.inter.blend)
            vsapi->requestFrameFilter(VSMIN(nright, N - 1), d->node, frameCtx);
    } else if (activationReason == arAllFramesReady) {
        if (time256 == 0)
            return vsapi->getFrameFilter(VSMIN(nleft, N - 1), d->node, frameCtx);
        if (time256 == 256)
            return vsapi->getFrameFilter(VSMIN(nright, N - 1), d->node, frameCtx);
        std::string msg;
        int rc = 0;
        int32_t usable = 0;
        VSFrame *dst = nullptr;
        if (inRange) {
            const VSFrame *mvF = vsapi->getFrameFilter(nright, d->mvfw, frameCtx), *mvB = vsapi->getFrameFilter(nleft, d->mvbw, frameCtx);
            const VSFrame *mvFF = extra ? vsapi->getFrameFilter(nleft, d->mvfw, frameCtx) : nullptr, *mvBB = extra ? vsapi->getFrameFilter(nright, d->mvbw, frameCtx) : nullptr;
            const VSFrame *left = vsapi->getFrameFilter(nleft, d->super, frameCtx), *right = vsapi->getFrameFilter(nright, d->super, frameCtx);
            bool missing = false;
            const uint8_t *vb = vectorsOf(vsapi, mvB, &missing), *vf = vectorsOf(vsapi, mvF, &missing);
            const uint8_t *vbb = extra ? vectorsOf(vsapi, mvBB, &missing) : nullptr, *vff = extra ? vectorsOf(vsapi, mvFF, &missing) : nullptr;
            dst = vsapi->newVideoFrame(&d->vi.format, d->vi.width, d->vi.height, left, core);
            if (missing) {
                rc = 1;
                msg = "MVTools_vectors property missing";
            } else {
                const uint8_t *lp[3], *rp[3];
                ptrdiff_t ls[3], rs[3];
                uint8_t *dp_[3];
                ptrdiff_t ds[3];
                planesRO(vsapi, left, d->supervi->format.numPlanes, lp, ls);
                planesRO(vsapi, right, d->supervi->format.numPlanes, rp, rs);
                planesRW(vsapi, dst, d->vi.format.numPlanes, dp_, ds);
                std::lock_guard<std::mutex> lock(g_engine_mutex);
                mvb200_ctx *ctx = engine(msg);
                rc = ctx ? mvb200_flowinter_frame(ctx, &d->p.inter, &d->sp, time256, lp, ls, rp, rs, vb, vf, vbb, vff, dp_, ds, &usable) : 1;
                if (ctx && rc)
                    msg = mvb200_last_error();
            }
            vsapi->freeFrame(mvF);
            vsapi->freeFrame(mvB);
            vsapi->freeFrame(mvFF);
            vsapi->freeFrame(mvBB);
            vsapi->freeFrame(left);
            vsapi->freeFrame(right);
        }
        if (!rc && !usable) { // poor estimation, :470-520
            vsapi->freeFrame(dst);
            dst = nullptr;
            const VSFrame *src = vsapi->getFrameFilter(VSMIN(nleft, N - 1), d->node, frameCtx);
            if (!d->p.inter.blend)
                return src;
            const VSFrame *ref = vsapi->getFrameFilter(VSMIN(nright, N - 1), d->node, frameCtx);
            dst = vsapi->newVideoFrame(&d->vi.format, d->vi.width, d->vi.height, src, core);
            const uint8_t *sp_[3], *rp[3];
            ptrdiff_t ss[3], rs[3];
            uint8_t *dp_[3];
            ptrdiff_t ds[3];
            planesRO(vsapi, src, d->vi.format.numPlanes, sp_, ss);
            planesRO(vsapi, ref, d->vi.format.numPlanes, rp, rs);
            planesRW(vsapi, dst, d->vi.format.numPlanes, dp_, ds);
            {
                std::lock_guard<std::mutex> lock(g_engine_mutex);
                mvb200_ctx *ctx = engine(msg);
                rc = ctx ? mvb200_blend_frame(ctx, &d->sp, time256, sp_, ss, rp, rs, dp_, ds) : 1;
                if (ctx && rc)
                    msg = mvb200_last_error();
            }
            vsapi->freeFrame(src);
            vsapi->freeFrame(ref);
        }
        if (rc) {
            vsapi->setFilterError(("FlowFPS: " + msg).c_str(), frameCtx);
            vsapi->freeFrame(dst);
            return nullptr;
        }
        return dst;
    }
    return nullptr;
}

void VS_CC flowFpsFree(void *instanceData, VSCore *, const VSAPI *vsapi) {
    FlowFPSData *d = (FlowFPSData *)instanceData;
    vsapi->freeNode(d->node);
    vsapi->freeNode(d->super);
    vsapi->freeNode(d->mvbw);
    vsapi->freeNode(d->mvfw);
    delete d->fused;
    delete d;
}

void VS_CC flowFpsCreate(const VSMap *in, VSMap *out, void *, VSCore *core, const VSAPI *vsapi) {
    mvb200_flowfps_args a;
    a.num = argInt64(vsapi, in, "num");
    a.den = argInt64(vsapi, in, "den");
    a.mask = argInt(vsapi, in, "mask");
    a.ml = argFloat(vsapi, in, "ml");
    a.blend = argInt(vsapi, in, "blend");
    a.thscd1 = argInt64(vsapi, in, "thscd1");
    a.thscd2 = argInt(vsapi, in, "thscd2");
    // argument checks before any clip is touched (MVFlowFPS.c:604-612)
    if (a.mask != MVB200_UNSET && (a.mask < 0 || a.mask > 2)) {
        vsapi->mapSetError(out, "FlowFPS: mask must be 0, 1, or 2.");
        return;
    }
    if (a.ml == a.ml && a.ml <= 0.0) {
        vsapi->mapSetError(out, "FlowFPS: ml must be greater than 0.");
        return;
    }
    FlowFPSData d;
    memset(&d, 0, sizeof(d));
    auto freeAll = [&]() {
        vsapi->freeNode(d.node);
        vsapi->freeNode(d.super);
        vsapi->freeNode(d.mvbw);
        vsapi->freeNode(d.mvfw);
    };
    d.super = vsapi->mapGetNode(in, "super", 0, nullptr);
    d.supervi = vsapi->getVideoInfo(d.super);
    std::string msg;
    mvb200_analysis_data vdb, vdf;
    if (!superFromClip(d.super, vsapi, "FlowFPS", &d.sp, msg)) {
        vsapi->mapSetError(out, msg.c_str());
        freeAll();
        return;
    }
    d.mvbw = vsapi->mapGetNode(in, "mvbw", 0, nullptr);
    d.mvfw = vsapi->mapGetNode(in, "mvfw", 0, nullptr);
    if (!adataFromClip(d.mvbw, vsapi, "FlowFPS", "mvbw", &vdb, msg) || !adataFromClip(d.mvfw, vsapi, "FlowFPS", "mvfw", &vdf, msg)) {
        vsapi->mapSetError(out, msg.c_str());
        freeAll();
        return;
    }
    d.node = vsapi->mapGetNode(in, "clip", 0, nullptr);
    d.oldvi = vsapi->getVideoInfo(d.node);
    d.vi = *d.oldvi;
    if (mvb200_flowfps_init(&a, &d.sp, &vdb, &vdf, d.vi.fpsNum, d.vi.fpsDen, d.vi.width, d.vi.height, d.vi.numFrames, &d.p)) {
        vsapi->mapSetError(out, mvb200_last_error());
        freeAll();
        return;
    }
    if (!formatOk(&d.vi) || d.vi.format.bitsPerSample != d.sp.bits) {
        vsapi->mapSetError(out, "FlowFPS: input clip must be GRAY, 420, 422, 440, or 444, up to 16 bits, with constant dimensions.");
        freeAll();
        return;
    }
    d.vi.fpsNum = d.p.fps_num;
    d.vi.fpsDen = d.p.fps_den;
    d.vi.numFrames = d.p.num_frames_out;
    const int window = lookaheadFromEnv();
    mvb200_super_params ssp;
    mvb200_analyse_params sap[2];
    VSNode *vecs[2] = {d.mvbw, d.mvfw};
    if (window && canonicalChain(d.node, d.super, vecs, 2, d.sp, d.oldvi->numFrames, true, &ssp, sap)) { // the fused path
        d.fused = new FlowFPSFused;
        d.fused->apBw = sap[0];
        d.fused->apFw = sap[1];
        d.sp = ssp;
        mvb200_clip_layout(&d.sp, &d.fused->cl);
        d.fused->window = fusedWindowFrames(d.fused, window, sap, &d.sp, 2);
    }
    FlowFPSData *data = new FlowFPSData(d);
    VSFilterDependency deps[4] = {{data->node, rpGeneral}, {data->super, rpGeneral}, {data->mvbw, rpGeneral}, {data->mvfw, rpGeneral}};
    if (data->fused)
        vsapi->createVideoFilter(out, "FlowFPS", &data->vi, flowFpsFusedGetFrame, flowFpsFree, fmParallel, deps, 1, data, core);
    else
        vsapi->createVideoFilter(out, "FlowFPS", &data->vi, flowFpsGetFrame, flowFpsFree, fmParallel, deps, 4, data, core);
    // AssumeFPS sets the _DurationNum and _DurationDen properties (MVFlowFPS.c:883-910)
    VSPlugin *stdPlugin = vsapi->getPluginByID("com.vapoursynth.std", core);
    VSMap *args = vsapi->createMap();
    VSNode *node = vsapi->mapGetNode(out, "clip", 0, nullptr);
    vsapi->mapSetNode(args, "clip", node, maReplace);
    vsapi->freeNode(node);
    vsapi->mapSetInt(args, "fpsnum", data->vi.fpsNum, maReplace);
    vsapi->mapSetInt(args, "fpsden", data->vi.fpsDen, maReplace);
    VSMap *ret = vsapi->invoke(stdPlugin, "AssumeFPS", args);
    if (vsapi->mapGetError(ret)) {
        vsapi->mapSetError(out, (std::string("FlowFPS: Failed to invoke AssumeFPS. Error message: ") + vsapi->mapGetError(ret)).c_str());
        vsapi->freeMap(args);
        vsapi->freeMap(ret);
        return;
    }
    node = vsapi->mapGetNode(ret, "clip", 0, nullptr);
    vsapi->freeMap(ret);
    vsapi->freeMap(args);
    vsapi->mapSetNode(out, "clip", node, maReplace);
    vsapi->freeNode(node);
}

// --------------------------------------------------------------- FlowBlur
// mvflowblurCreate / GetFrame, MVFlowBlur.c:144-567
// Fused path (see fusedGetFrame): a window of B frames is one mvb200_flowblur_clip_host call over source frames.
// (Frame 0 is never blurred -- it has no frame before it -- so no output ever carries super frame 0's Super_* properties.)
struct FlowBlurFused : FusedCache {
    mvb200_analyse_params apBw, apFw;
};

struct FlowBlurData {
    VSNode *node, *super, *mvbw, *mvfw;
    const VSVideoInfo *vi, *supervi;
    mvb200_super_params sp;
    mvb200_flowblur_params p;
    FlowBlurFused *fused; // owned; null = node-by-node evaluation
};

const VSFrame *VS_CC flowBlurFusedGetFrame(int n, int activationReason, void *instanceData, void **frameData, VSFrameContext *frameCtx, VSCore *core, const VSAPI *vsapi) {
    FlowBlurData *d = (FlowBlurData *)instanceData;
    FlowBlurFused *f = d->fused;
    const int off = d->p.vd.nDeltaFrame, N = d->vi->numFrames;
    auto segment = [&](int w0, int w1, int *s0, int *s1) {
        *s0 = w0 - off > 0 ? w0 - off : 0;
        *s1 = w1 + off < N ? w1 + off : N;
    };
    auto run = [&](mvb200_ctx *ctx, const void *hostSrc, int s0, int s1, int w0, int w1, void *hostDst, int32_t *flags) {
        return mvb200_flowblur_clip_host(ctx, &d->sp, &f->apBw, &f->apFw, &d->p, hostSrc, s0, s1 - s0, w0, w1 - w0, hostDst);
    };
    return fusedGetFrame("FlowBlur", f, d->node, d->vi, n, n, 7u, activationReason, frameData, frameCtx, core, vsapi, segment, run);
}

const VSFrame *VS_CC flowBlurGetFrame(int n, int activationReason, void *instanceData, void **, VSFrameContext *frameCtx, VSCore *core, const VSAPI *vsapi) {
    FlowBlurData *d = (FlowBlurData *)instanceData;
    const int off = d->p.vd.nDeltaFrame;
    const bool inRange = n - off >= 0 && n + off < d->vi->numFrames;
    if (activationReason == arInitial) { // :150-160
        if (inRange) {
            vsapi->requestFrameFilter(n - off, d->mvbw, frameCtx);
            vsapi->requestFrameFilter(n + off, d->mvfw, frameCtx);
        }
        vsapi->requestFrameFilter(n, d->super, frameCtx);
        vsapi->requestFrameFilter(n, d->node, frameCtx);
    } else if (activationReason == arAllFramesReady) {
        std::string msg;
        int rc = 0;
        int32_t usable = 0;
        VSFrame *dst = nullptr;
        if (inRange) {
            const VSFrame *mvF = vsapi->getFrameFilter(n + off, d->mvfw, frameCtx), *mvB = vsapi->getFrameFilter(n - off, d->mvbw, frameCtx);
            const VSFrame *ref = vsapi->getFrameFilter(n, d->super, frameCtx);
            bool missing = false;
            const uint8_t *vb = vectorsOf(vsapi, mvB, &missing), *vf = vectorsOf(vsapi, mvF, &missing);
            dst = vsapi->newVideoFrame(&d->vi->format, d->vi->width, d->vi->height, ref, core);
            if (missing) {
                rc = 1;
                msg = "MVTools_vectors property missing";
            } else {
                const uint8_t *rp[3];
                ptrdiff_t rs[3];
                uint8_t *dp_[3];
                ptrdiff_t ds[3];
                planesRO(vsapi, ref, d->supervi->format.numPlanes, rp, rs);
                planesRW(vsapi, dst, d->vi->format.numPlanes, dp_, ds);
                std::lock_guard<std::mutex> lock(g_engine_mutex);
                mvb200_ctx *ctx = engine(msg);
                rc = ctx ? mvb200_flowblur_frame(ctx, &d->p, &d->sp, rp, rs, vb, vf, dp_, ds, &usable) : 1;
                if (ctx && rc)
                    msg = mvb200_last_error();
            }
            vsapi->freeFrame(mvF);
            vsapi->freeFrame(mvB);
            vsapi->freeFrame(ref);
        }
        if (rc) {
            vsapi->setFilterError(("FlowBlur: " + msg).c_str(), frameCtx);
            vsapi->freeFrame(dst);
            return nullptr;
        }
        if (!usable) { // :331-335
            vsapi->freeFrame(dst);
            return vsapi->getFrameFilter(n, d->node, frameCtx);
        }
        return dst;
    }
    return nullptr;
}

void VS_CC flowBlurFree(void *instanceData, VSCore *, const VSAPI *vsapi) {
    FlowBlurData *d = (FlowBlurData *)instanceData;
    vsapi->freeNode(d->node);
    vsapi->freeNode(d->super);
    vsapi->freeNode(d->mvbw);
    vsapi->freeNode(d->mvfw);
    delete d->fused;
    delete d;
}

void VS_CC flowBlurCreate(const VSMap *in, VSMap *out, void *, VSCore *core, const VSAPI *vsapi) {
    mvb200_flowblur_args a;
    a.blur = argFloat(vsapi, in, "blur");
    a.prec = argInt(vsapi, in, "prec");
    a.thscd1 = argInt64(vsapi, in, "thscd1");
    a.thscd2 = argInt(vsapi, in, "thscd2");
    // argument checks before any clip is touched (MVFlowBlur.c:374-382)
    if (a.blur == a.blur && ((float)a.blur < 0.0f || (float)a.blur > 200.0f)) {
        vsapi->mapSetError(out, "FlowBlur: blur must be between 0 and 200 % (inclusive).");
        return;
    }
    if (a.prec != MVB200_UNSET && a.prec < 1) {
        vsapi->mapSetError(out, "FlowBlur: prec must be at least 1.");
        return;
    }
    FlowBlurData d;
    memset(&d, 0, sizeof(d));
    auto freeAll = [&]() {
        vsapi->freeNode(d.node);
        vsapi->freeNode(d.super);
        vsapi->freeNode(d.mvbw);
        vsapi->freeNode(d.mvfw);
    };
    d.super = vsapi->mapGetNode(in, "super", 0, nullptr);
    d.supervi = vsapi->getVideoInfo(d.super);
    std::string msg;
    mvb200_analysis_data vdb, vdf;
    if (!superFromClip(d.super, vsapi, "FlowBlur", &d.sp, msg)) {
        vsapi->mapSetError(out, msg.c_str());
        freeAll();
        return;
    }
    d.mvbw = vsapi->mapGetNode(in, "mvbw", 0, nullptr);
    d.mvfw = vsapi->mapGetNode(in, "mvfw", 0, nullptr);
    if (!adataFromClip(d.mvbw, vsapi, "FlowBlur", "mvbw", &vdb, msg) || !adataFromClip(d.mvfw, vsapi, "FlowBlur", "mvfw", &vdf, msg)) {
        vsapi->mapSetError(out, msg.c_str());
        freeAll();
        return;
    }
    d.node = vsapi->mapGetNode(in, "clip", 0, nullptr);
    d.vi = vsapi->getVideoInfo(d.node);
    if (mvb200_flowblur_init(&a, &d.sp, &vdb, &vdf, d.vi->numFrames, &d.p)) {
        vsapi->mapSetError(out, mvb200_last_error());
        freeAll();
        return;
    }
    if (!formatOk(d.vi)) {
        vsapi->mapSetError(out, "FlowBlur: input clip must be GRAY, 420, 422, 440, or 444, up to 16 bits, with constant dimensions.");
        freeAll();
        return;
    }
    if (d.vi->width != vdb.nWidth || d.vi->height != vdb.nHeight || d.vi->format.bitsPerSample != d.sp.bits) {
        vsapi->mapSetError(out, "FlowBlur: wrong source or super clip frame size.");
        freeAll();
        return;
    }
    const int window = lookaheadFromEnv();
    mvb200_super_params ssp;
    mvb200_analyse_params sap[2];
    VSNode *vecs[2] = {d.mvbw, d.mvfw};
    if (window && canonicalChain(d.node, d.super, vecs, 2, d.sp, d.vi->numFrames, true, &ssp, sap)) {
        d.fused = new FlowBlurFused;
        d.fused->apBw = sap[0];
        d.fused->apFw = sap[1];
        d.sp = ssp;
        mvb200_clip_layout(&d.sp, &d.fused->cl);
        d.fused->window = fusedWindowFrames(d.fused, window, sap, &d.sp, 2);
    }
    FlowBlurData *data = new FlowBlurData(d);
    if (data->fused) { // only the source clip is read on this path
        VSFilterDependency fdeps[1] = {{data->node, rpGeneral}};
        vsapi->createVideoFilter(out, "FlowBlur", data->vi, flowBlurFusedGetFrame, flowBlurFree, fmParallel, fdeps, 1, data, core);
        return;
    }
    VSFilterDependency deps[4] = {{data->node, rpStrictSpatial}, {data->super, rpStrictSpatial}, {data->mvbw, rpGeneral}, {data->mvfw, rpGeneral}};
    vsapi->createVideoFilter(out, "FlowBlur", data->vi, flowBlurGetFrame, flowBlurFree, fmParallel, deps, 4, data, core);
}

// --------------------------------------------------------------- BlockFPS
// mvblockfpsCreate / GetFrame, MVBlockFPS.c:240-1034
// Fused path (see fusedGetFrame): a window of B OUTPUT frames is one mvb200_blockfps_clip_host call over the source
// frames the window needs (every output frame's left and right frames; the two searches look at each other's frame).
struct BlockFPSFused : FusedCache {
    mvb200_analyse_params apBw, apFw;
};

struct BlockFPSData {
    VSNode *node, *super, *mvbw, *mvfw;
    VSVideoInfo vi;
    const VSVideoInfo *oldvi, *supervi;
    mvb200_super_params sp;
    mvb200_blockfps_params p;
    BlockFPSFused *fused; // owned; null = node-by-node evaluation
};

const VSFrame *VS_CC blockFpsFusedGetFrame(int n, int activationReason, void *instanceData, void **frameData, VSFrameContext *frameCtx, VSCore *core, const VSAPI *vsapi) {
    BlockFPSData *d = (BlockFPSData *)instanceData;
    BlockFPSFused *f = d->fused;
    const int off = d->p.vd.nDeltaFrame, N = d->oldvi->numFrames;
    int32_t nleft, time256;
    mvb200_blockfps_position(&d->p, n, &nleft, &time256);
    const int propFrame = VSMIN(time256 == 256 ? nleft + off : nleft, N - 1); // whose properties the output carries
    if (time256 == 0 || time256 == 256) { // the input frame itself (MVBlockFPS.c:246-253)
        if (activationReason == arInitial)
            vsapi->requestFrameFilter(propFrame, d->node, frameCtx);
        return activationReason == arAllFramesReady ? vsapi->getFrameFilter(propFrame, d->node, frameCtx) : nullptr;
    }
    auto segment = [&](int w0, int w1, int *s0, int *s1) {
        int32_t lmin, lmax, t;
        mvb200_blockfps_position(&d->p, w0, &lmin, &t);
        mvb200_blockfps_position(&d->p, w1 - 1, &lmax, &t);
        *s0 = lmin < N - 1 ? lmin : N - 1; // (output frames past the end are made of the last input frame)
        *s1 = lmax + off + 1 < N ? lmax + off + 1 : N;
    };
    auto run = [&](mvb200_ctx *ctx, const void *hostSrc, int s0, int s1, int w0, int w1, void *hostDst, int32_t *flags) {
        return mvb200_blockfps_clip_host(ctx, &d->sp, &f->apBw, &f->apFw, &d->p, hostSrc, s0, s1 - s0, w0, w1 - w0, hostDst, flags);
    };
    int32_t usable = 0;
    const VSFrame *out = fusedGetFrame("BlockFPS", f, d->node, &d->vi, n, propFrame, 7u, activationReason, frameData, frameCtx, core, vsapi, segment, run, -1, &usable);
    if (out && usable && nleft == 0) // an interpolated frame is built on the left super frame (MVBlockFPS.c:287): Super_* on frame 0
        setSuperProps(vsapi, const_cast<VSFrame *>(out), d->sp);
    return out;
}

const VSFrame *VS_CC blockFpsGetFrame(int n, int activationReason, void *instanceData, void **, VSFrameContext *frameCtx, VSCore *core, const VSAPI *vsapi) {
    BlockFPSData *d = (BlockFPSData *)instanceData;
    const int off = d->p.vd.nDeltaFrame, N = d->oldvi->numFrames;
    int32_t nleft, time256;
    mvb200_blockfps_position(&d->p, n, &nleft, &time256);
    const int nright = nleft + off;
    const bool inRange = nleft < N && nright < N;
    if (activationReason == arInitial) { // :246-272
        if (time256 == 0) {
            vsapi->requestFrameFilter(VSMIN(nleft, N - 1), d->node, frameCtx);
            return nullptr;
        } else if (time256 == 256) {
            vsapi->requestFrameFilter(VSMIN(nright, N - 1), d->node, frameCtx);
            return nullptr;
        }
        if (inRange) {
            vsapi->requestFrameFilter(nright, d->mvfw, frameCtx);
            vsapi->requestFrameFilter(nleft, d->mvbw, frameCtx);
            vsapi->requestFrameFilter(nleft, d->super, frameCtx);
            vsapi->requestFrameFilter(nright, d->super, frameCtx);
        }
        vsapi->requestFrameFilter(VSMIN(nleft, N - 1), d->node, frameCtx);
        if (d->p.blend)
            vsapi->requestFrameFilter(VSMIN(nright, N - 1), d->node, frameCtx);
    } else if (activationReason == arAllFramesReady) {
        if (time256 == 0)
            return vsapi->getFrameFilter(VSMIN(nleft, N - 1), d->node, frameCtx);
        if (time256 == 256)
            return vsapi->getFrameFilter(VSMIN(nright, N - 1), d->node, frameCtx);
        const int planes = (d->sp.modeyuv & (MVB200_UPLANE | MVB200_VPLANE)) ? 3 : 1;
        std::string msg;
        int rc = 0;
        int32_t usable = 0;
        VSFrame *dst = nullptr;
        if (inRange) {
            const VSFrame *mvF = vsapi->getFrameFilter(nright, d->mvfw, frameCtx), *mvB = vsapi->getFrameFilter(nleft, d->mvbw, frameCtx);
            const VSFrame *left = vsapi->getFrameFilter(nleft, d->super, frameCtx), *right = vsapi->getFrameFilter(nright, d->super, frameCtx);
            bool missing = false;
            const uint8_t *vb = vectorsOf(vsapi, mvB, &missing), *vf = vectorsOf(vsapi, mvF, &missing);
            dst = vsapi->newVideoFrame(&d->vi.format, d->vi.width, d->vi.height, left, core);
            if (missing) {
                rc = 1;
                msg = "MVTools_vectors property missing";
            } else {
                const uint8_t *lp[3], *rp[3];
                ptrdiff_t ls[3], rs[3];
                uint8_t *dp_[3];
                ptrdiff_t ds[3];
                planesRO(vsapi, left, d->supervi->format.numPlanes, lp, ls);
                planesRO(vsapi, right, d->supervi->format.numPlanes, rp, rs);
                planesRW(vsapi, dst, d->vi.format.numPlanes, dp_, ds);
                std::lock_guard<std::mutex> lock(g_engine_mutex);
                mvb200_ctx *ctx = engine(msg);
                rc = ctx ? mvb200_blockfps_frame(ctx, &d->p, &d->sp, time256, lp, ls, rp, rs, vb, vf, dp_, ds, &usable) : 1;
                if (ctx && rc)
                    msg = mvb200_last_error();
            }
            vsapi->freeFrame(mvF);
            vsapi->freeFrame(mvB);
            vsapi->freeFrame(left);
            vsapi->freeFrame(right);
        }
        if (!rc && !usable) { // poor estimation, :636-670
            vsapi->freeFrame(dst);
            dst = nullptr;
            const VSFrame *src = vsapi->getFrameFilter(VSMIN(nleft, N - 1), d->node, frameCtx);
            if (!d->p.blend)
                return src;
            const VSFrame *ref = vsapi->getFrameFilter(VSMIN(nright, N - 1), d->node, frameCtx);
            dst = vsapi->newVideoFrame(&d->vi.format, d->vi.width, d->vi.height, src, core);
            const uint8_t *sp_[3], *rp[3];
            ptrdiff_t ss[3], rs[3];
            uint8_t *dp_[3];
            ptrdiff_t ds[3];
            planesRO(vsapi, src, d->vi.format.numPlanes, sp_, ss);
            planesRO(vsapi, ref, d->vi.format.numPlanes, rp, rs);
            planesRW(vsapi, dst, d->vi.format.numPlanes, dp_, ds);
            mvb200_super_params used = d->sp; // only the planes the super clip carries are blended (:652)
            used.nplanes = planes < d->sp.nplanes ? planes : d->sp.nplanes;
            {
                std::lock_guard<std::mutex> lock(g_engine_mutex);
                mvb200_ctx *ctx = engine(msg);
                rc = ctx ? mvb200_blend_frame(ctx, &used, time256, sp_, ss, rp, rs, dp_, ds) : 1;
                if (ctx && rc)
                    msg = mvb200_last_error();
            }
            vsapi->freeFrame(src);
            vsapi->freeFrame(ref);
        }
        if (rc) {
            vsapi->setFilterError(("BlockFPS: " + msg).c_str(), frameCtx);
            vsapi->freeFrame(dst);
            return nullptr;
        }
        return dst;
    }
    return nullptr;
}

void VS_CC blockFpsFree(void *instanceData, VSCore *, const VSAPI *vsapi) {
    BlockFPSData *d = (BlockFPSData *)instanceData;
    vsapi->freeNode(d->node);
    vsapi->freeNode(d->super);
    vsapi->freeNode(d->mvbw);
    vsapi->freeNode(d->mvfw);
    delete d->fused;
    delete d;
}

void VS_CC blockFpsCreate(const VSMap *in, VSMap *out, void *, VSCore *core, const VSAPI *vsapi) {
    mvb200_blockfps_args a;
    a.num = argInt64(vsapi, in, "num");
    a.den = argInt64(vsapi, in, "den");
    a.mode = argInt(vsapi, in, "mode");
    a.ml = argFloat(vsapi, in, "ml");
    a.blend = argInt(vsapi, in, "blend");
    a.thscd1 = argInt64(vsapi, in, "thscd1");
    a.thscd2 = argInt(vsapi, in, "thscd2");
    if (a.mode != MVB200_UNSET && (a.mode < 0 || a.mode > 8)) { // before any clip is touched (MVBlockFPS.c:779-782)
        vsapi->mapSetError(out, "BlockFPS: mode must be between 0 and 8 (inclusive).");
        return;
    }
    BlockFPSData d;
    memset(&d, 0, sizeof(d));
    auto freeAll = [&]() {
        vsapi->freeNode(d.node);
        vsapi->freeNode(d.super);
        vsapi->freeNode(d.mvbw);
        vsapi->freeNode(d.mvfw);
    };
    d.super = vsapi->mapGetNode(in, "super", 0, nullptr);
    d.supervi = vsapi->getVideoInfo(d.super);
    std::string msg;
    mvb200_analysis_data vdb, vdf;
    if (!superFromClip(d.super, vsapi, "BlockFPS", &d.sp, msg)) {
        vsapi->mapSetError(out, msg.c_str());
        freeAll();
        return;
    }
    d.mvbw = vsapi->mapGetNode(in, "mvbw", 0, nullptr);
    d.mvfw = vsapi->mapGetNode(in, "mvfw", 0, nullptr);
    if (!adataFromClip(d.mvbw, vsapi, "BlockFPS", "mvbw", &vdb, msg) || !adataFromClip(d.mvfw, vsapi, "BlockFPS", "mvfw", &vdf, msg)) {
        vsapi->mapSetError(out, msg.c_str());
        freeAll();
        return;
    }
    d.node = vsapi->mapGetNode(in, "clip", 0, nullptr);
    d.oldvi = vsapi->getVideoInfo(d.node);
    d.vi = *d.oldvi;
    if (mvb200_blockfps_init(&a, &d.sp, &vdb, &vdf, d.vi.fpsNum, d.vi.fpsDen, d.vi.width, d.vi.height, d.vi.numFrames, &d.p)) {
        vsapi->mapSetError(out, mvb200_last_error());
        freeAll();
        return;
    }
    if (!formatOk(&d.vi) || d.vi.format.bitsPerSample != d.sp.bits) {
        vsapi->mapSetError(out, "BlockFPS: input clip must be GRAY, 420, 422, 440, or 444, up to 16 bits, with constant dimensions.");
        freeAll();
        return;
    }
    d.vi.fpsNum = d.p.fps_num;
    d.vi.fpsDen = d.p.fps_den;
    d.vi.numFrames = d.p.num_frames_out;
    d.fused = nullptr;
    const int window = lookaheadFromEnv();
    mvb200_super_params ssp;
    mvb200_analyse_params sap[2];
    VSNode *vecs[2] = {d.mvbw, d.mvfw};
    if (window && canonicalChain(d.node, d.super, vecs, 2, d.sp, d.oldvi->numFrames, true, &ssp, sap)) { // the fused path
        d.fused = new BlockFPSFused;
        d.fused->apBw = sap[0];
        d.fused->apFw = sap[1];
        d.sp = ssp;
        mvb200_clip_layout(&d.sp, &d.fused->cl);
        d.fused->window = fusedWindowFrames(d.fused, window, sap, &d.sp, 2);
    }
    BlockFPSData *data = new BlockFPSData(d);
    VSFilterDependency deps[4] = {{data->node, rpGeneral}, {data->super, rpGeneral}, {data->mvbw, rpGeneral}, {data->mvfw, rpGeneral}};
    if (data->fused) // only the source clip is read on this path
        vsapi->createVideoFilter(out, "BlockFPS", &data->vi, blockFpsFusedGetFrame, blockFpsFree, fmParallel, deps, 1, data, core);
    else
        vsapi->createVideoFilter(out, "BlockFPS", &data->vi, blockFpsGetFrame, blockFpsFree, fmParallel, deps, 4, data, core);
    // AssumeFPS sets the _DurationNum and _DurationDen properties (MVBlockFPS.c:990-1012)
    VSPlugin *stdPlugin = vsapi->getPluginByID("com.vapoursynth.std", core);
    VSMap *args = vsapi->createMap();
    VSNode *node = vsapi->mapGetNode(out, "clip", 0, nullptr);
    vsapi->mapSetNode(args, "clip", node, maReplace);
    vsapi->freeNode(node);
    vsapi->mapSetInt(args, "fpsnum", data->vi.fpsNum, maReplace);
    vsapi->mapSetInt(args, "fpsden", data->vi.fpsDen, maReplace);
    VSMap *ret = vsapi->invoke(stdPlugin, "AssumeFPS", args);
    if (vsapi->mapGetError(ret)) {
        vsapi->mapSetError(out, (std::string("BlockFPS: Failed to invoke AssumeFPS. Error message: ") + vsapi->mapGetError(ret)).c_str());
        vsapi->freeMap(args);
        vsapi->freeMap(ret);
        return;
    }
    node = vsapi->mapGetNode(ret, "clip", 0, nullptr);
    vsapi->freeMap(ret);
    vsapi->freeMap(args);
    vsapi->mapSetNode(out, "clip", node, maReplace);
    vsapi->freeNode(node);
}

} // namespace

VS_EXTERNAL_API(void) VapourSynthPluginInit2(VSPlugin *plugin, const VSPLUGINAPI *vspapi) {
    vspapi->configPlugin("com.nodame.mvtools", "mv", "MVTools v24 (B200 engine)", VS_MAKE_VERSION(24, 0), VS_MAKE_VERSION(VAPOURSYNTH_API_MAJOR, VAPOURSYNTH_API_MINOR), 0, plugin);
    // MVSuper.c:279-291
    vspapi->registerFunction("Super",
                             "clip:vnode;hpad:int:opt;vpad:int:opt;pel:int:opt;levels:int:opt;chroma:int:opt;sharp:int:opt;rfilter:int:opt;pelclip:vnode:opt;opt:int:opt;",
                             "clip:vnode;", superCreate, nullptr, plugin);
    // MVAnalyse.c:639-671
    vspapi->registerFunction("Analyse",
                             "super:vnode;blksize:int:opt;blksizev:int:opt;levels:int:opt;search:int:opt;searchparam:int:opt;pelsearch:int:opt;isb:int:opt;lambda:int:opt;"
                             "chroma:int:opt;delta:int:opt;truemotion:int:opt;lsad:int:opt;plevel:int:opt;global:int:opt;pnew:int:opt;pzero:int:opt;pglobal:int:opt;"
                             "overlap:int:opt;overlapv:int:opt;divide:int:opt;badsad:int:opt;badrange:int:opt;opt:int:opt;meander:int:opt;trymany:int:opt;fields:int:opt;"
                             "tff:int:opt;search_coarse:int:opt;dct:int:opt;",
                             "clip:vnode;", analyseCreate, nullptr, plugin);
    // MVDegrains.cpp:813-932
    static const char *vnames[12] = {"mvbw", "mvfw", "mvbw2", "mvfw2", "mvbw3", "mvfw3", "mvbw4", "mvfw4", "mvbw5", "mvfw5", "mvbw6", "mvfw6"};
    for (int radius = 1; radius <= 6; radius++) {
        std::string args = "clip:vnode;super:vnode;";
        for (int r = 0; r < 2 * radius; r++)
            args += std::string(vnames[r]) + ":vnode;";
        args += "thsad:int:opt;thsadc:int:opt;plane:int:opt;limit:int:opt;limitc:int:opt;thscd1:int:opt;thscd2:int:opt;opt:int:opt;";
        vspapi->registerFunction(("Degrain" + std::to_string(radius)).c_str(), args.c_str(), "clip:vnode;", degrainCreate, (void *)(intptr_t)radius, plugin);
    }
    // MVCompensate.c:579-592, MVRecalculate.c:557-579
    vspapi->registerFunction("Compensate",
                             "clip:vnode;super:vnode;vectors:vnode;scbehavior:int:opt;thsad:int:opt;fields:int:opt;time:float:opt;thscd1:int:opt;thscd2:int:opt;opt:int:opt;"
                             "tff:int:opt;",
                             "clip:vnode;", compCreate, nullptr, plugin);
    vspapi->registerFunction("Recalculate",
                             "super:vnode;vectors:vnode;thsad:int:opt;smooth:int:opt;blksize:int:opt;blksizev:int:opt;search:int:opt;searchparam:int:opt;lambda:int:opt;"
                             "chroma:int:opt;truemotion:int:opt;pnew:int:opt;overlap:int:opt;overlapv:int:opt;divide:int:opt;opt:int:opt;meander:int:opt;fields:int:opt;"
                             "tff:int:opt;dct:int:opt;",
                             "clip:vnode;", recalcCreate, nullptr, plugin);
    // MVMask.c:347-362, MVFinest.c:210-216, MVFlow.cpp:596-610, MVSCDetection.c:134-142
    vspapi->registerFunction("Mask", "clip:vnode;vectors:vnode;ml:float:opt;gamma:float:opt;kind:int:opt;time:float:opt;ysc:int:opt;thscd1:int:opt;thscd2:int:opt;opt:int:opt;", "clip:vnode;", maskCreate, nullptr, plugin);
    vspapi->registerFunction("Finest", "super:vnode;opt:int:opt;", "clip:vnode;", finestCreate, nullptr, plugin);
    vspapi->registerFunction("Flow", "clip:vnode;super:vnode;vectors:vnode;time:float:opt;mode:int:opt;fields:int:opt;thscd1:int:opt;thscd2:int:opt;opt:int:opt;tff:int:opt;", "clip:vnode;", flowCreate, nullptr, plugin);
    vspapi->registerFunction("SCDetection", "clip:vnode;vectors:vnode;thscd1:int:opt;thscd2:int:opt;", "clip:vnode;", scdCreate, nullptr, plugin);
    // MVBlockFPS.c:1017-1033
    vspapi->registerFunction("BlockFPS", "clip:vnode;super:vnode;mvbw:vnode;mvfw:vnode;num:int:opt;den:int:opt;mode:int:opt;ml:float:opt;blend:int:opt;thscd1:int:opt;thscd2:int:opt;opt:int:opt;", "clip:vnode;", blockFpsCreate, nullptr, plugin);
    // MVFlowBlur.c:550-566
    vspapi->registerFunction("FlowBlur", "clip:vnode;super:vnode;mvbw:vnode;mvfw:vnode;blur:float:opt;prec:int:opt;thscd1:int:opt;thscd2:int:opt;opt:int:opt;", "clip:vnode;", flowBlurCreate, nullptr, plugin);
    // MVFlowFPS.c:913-928
    vspapi->registerFunction("FlowFPS", "clip:vnode;super:vnode;mvbw:vnode;mvfw:vnode;num:int:opt;den:int:opt;mask:int:opt;ml:float:opt;blend:int:opt;thscd1:int:opt;thscd2:int:opt;opt:int:opt;", "clip:vnode;", flowFpsCreate, nullptr, plugin);
    // MVFlowInter.c:696-711
    vspapi->registerFunction("FlowInter", "clip:vnode;super:vnode;mvbw:vnode;mvfw:vnode;time:float:opt;ml:float:opt;blend:int:opt;thscd1:int:opt;thscd2:int:opt;opt:int:opt;", "clip:vnode;", flowInterCreate, nullptr, plugin);
}
