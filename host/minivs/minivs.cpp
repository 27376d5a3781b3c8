// A frame-server stand-in (test and measurement harness) -- not part of the product path.
//
// minivs: a minimal, self-contained VapourSynth API-4 *host* (the side that a
// real VapourSynth core plays).  No VapourSynth runtime exists in the build
// container, so the parity tests and the CPU-baseline leg of bench.py use this
// host to drive EITHER plugin through the plugin API the reference defines
// (/root/reference/src/EntryPoint.c:28-50):
//   * oracle/_ref/libmvtools_ref.so  (the unmodified reference sources), and
//   * vapoursynth-mvtools_b200/libmvtools.so (this repo's drop-in plugin),
// with the same "script" (Super -> Analyse xN -> DegrainN ...), and compare the
// returned frames and frame properties byte for byte.
//
// It implements the VSAPI members the five hot-path filters use (SURVEY.md
// section 8b lists them) plus what a driver needs; every other member is NULL.
// Frames are evaluated synchronously and recursively: arInitial collects
// requestFrameFilter() calls, the requests are evaluated depth-first, then
// arAllFramesReady runs.  A per-node LRU cache plus an in-flight table make it
// safe (and non-redundant) to pull frames from several host threads at once,
// which is how the fmParallel CPU baseline is timed (mvs_render).
//
// A plain C API (mvs_*) at the bottom is what tests/ bind with ctypes.

#include <VapourSynth4.h>

#include <dlfcn.h>
#include <malloc.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <utility>
#include <vector>

namespace {

constexpr int kAlign = 64;

struct Value {
    int type = ptUnset;
    int64_t i = 0;
    double f = 0;
    std::string data;
    int hint = dtUnknown;
    VSNode *node = nullptr;
    const VSFrame *frame = nullptr;
};

} // namespace

struct VSMap {
    std::map<std::string, std::vector<Value>> kv;
    bool hasError = false;
    std::string error;
    ~VSMap();
    void clear();
};

struct VSFrame {
    std::atomic<int> refs{1};
    VSVideoFormat fmt{};
    int width = 0, height = 0;
    uint8_t *plane[3] = {nullptr, nullptr, nullptr};
    ptrdiff_t stride[3] = {0, 0, 0};
    VSMap props;
    int pw(int p) const { return p ? width >> fmt.subSamplingW : width; }
    int ph(int p) const { return p ? height >> fmt.subSamplingH : height; }
    ~VSFrame() {
        for (auto &p : plane)
            free(p);
    }
};

struct VSCore {
    std::vector<VSPlugin *> plugins;
    std::mutex logMutex;
};

struct VSPluginFunction {
    std::string name, args, ret;
    VSPublicFunction fn = nullptr;
    void *userData = nullptr;
};

struct VSPlugin {
    VSCore *core = nullptr;
    void *handle = nullptr;
    std::string id, ns, name, path;
    int version = 0, apiVersion = 0;
    std::map<std::string, VSPluginFunction> funcs;
};

struct VSNode {
    std::atomic<int> refs{1};
    VSCore *core = nullptr;
    std::string name;
    VSVideoInfo vi{};
    // filter
    VSFilterGetFrame getFrame = nullptr;
    VSFilterFree freeFn = nullptr;
    void *instanceData = nullptr;
    std::vector<VSNode *> deps; // referenced for lifetime only
    // source clip
    std::vector<const VSFrame *> sourceFrames;
    std::map<int, int> sourceDelayMs; // test aid: fetching these frames of a source clip takes this long
    // cache
    std::mutex m;
    std::condition_variable cv;
    size_t cacheSize = 24;
    std::list<std::pair<int, const VSFrame *>> lru;
    std::map<int, int> inflight;
    std::atomic<long> computed{0};
};

struct VSFrameContext {
    std::vector<std::pair<VSNode *, int>> requests;
    std::map<std::pair<VSNode *, int>, const VSFrame *> ready;
    bool hasError = false;
    std::string error;
};

static const VSAPI *g_api();

VSMap::~VSMap() { clear(); }

void VSMap::clear() {
    for (auto &e : kv)
        for (auto &v : e.second) {
            if (v.node)
                g_api()->freeNode(v.node);
            if (v.frame)
                g_api()->freeFrame(v.frame);
        }
    kv.clear();
    hasError = false;
    error.clear();
}

namespace {

// ------------------------------------------------------------------ frames

VSFrame *allocFrame(const VSVideoFormat *f, int w, int h) {
    VSFrame *fr = new VSFrame;
    fr->fmt = *f;
    fr->width = w;
    fr->height = h;
    for (int p = 0; p < f->numPlanes; p++) {
        size_t row = (size_t)fr->pw(p) * f->bytesPerSample;
        fr->stride[p] = (ptrdiff_t)((row + kAlign - 1) / kAlign * kAlign);
        size_t bytes = (size_t)fr->stride[p] * fr->ph(p);
        void *mem = nullptr;
        if (posix_memalign(&mem, kAlign, bytes ? bytes : kAlign))
            abort();
        // Real VapourSynth hands out uninitialised memory; poison it so that a
        // filter relying on zeroed frames is caught by the parity tests.
        memset(mem, 0xA5, bytes);
        fr->plane[p] = (uint8_t *)mem;
    }
    return fr;
}

void copyProps(const VSMap &src, VSMap &dst) {
    dst.clear();
    for (auto &e : src.kv) {
        auto &out = dst.kv[e.first];
        for (auto v : e.second) {
            if (v.node)
                g_api()->addNodeRef(v.node);
            if (v.frame)
                g_api()->addFrameRef(v.frame);
            out.push_back(v);
        }
    }
}

// copyMap of the API: values of src replace same-named keys of dst, other keys of dst stay
void VS_CC copyMap(const VSMap *src, VSMap *dst) {
    for (auto &e : src->kv) {
        auto &out = dst->kv[e.first];
        for (auto v : out) {
            if (v.node)
                g_api()->freeNode(v.node);
            if (v.frame)
                g_api()->freeFrame(v.frame);
        }
        out.clear();
        for (auto v : e.second) {
            if (v.node)
                g_api()->addNodeRef(v.node);
            if (v.frame)
                g_api()->addFrameRef(v.frame);
            out.push_back(v);
        }
    }
}

VSFrame *VS_CC newVideoFrame(const VSVideoFormat *format, int width, int height, const VSFrame *propSrc, VSCore *) {
    VSFrame *f = allocFrame(format, width, height);
    if (propSrc)
        copyProps(propSrc->props, f->props);
    return f;
}

VSFrame *VS_CC copyFrame(const VSFrame *src, VSCore *) {
    VSFrame *f = allocFrame(&src->fmt, src->width, src->height);
    for (int p = 0; p < src->fmt.numPlanes; p++)
        memcpy(f->plane[p], src->plane[p], (size_t)src->stride[p] * src->ph(p));
    copyProps(src->props, f->props);
    return f;
}

void VS_CC freeFrame(const VSFrame *f) {
    if (!f)
        return;
    VSFrame *m = const_cast<VSFrame *>(f);
    if (m->refs.fetch_sub(1) == 1)
        delete m;
}

const VSFrame *VS_CC addFrameRef(const VSFrame *f) {
    const_cast<VSFrame *>(f)->refs.fetch_add(1);
    return f;
}

const VSMap *VS_CC getFramePropertiesRO(const VSFrame *f) { return &f->props; }
VSMap *VS_CC getFramePropertiesRW(VSFrame *f) { return &f->props; }
ptrdiff_t VS_CC getStride(const VSFrame *f, int plane) { return f->stride[plane]; }
const uint8_t *VS_CC getReadPtr(const VSFrame *f, int plane) { return f->plane[plane]; }
uint8_t *VS_CC getWritePtr(VSFrame *f, int plane) { return f->plane[plane]; }
const VSVideoFormat *VS_CC getVideoFrameFormat(const VSFrame *f) { return &f->fmt; }
int VS_CC getFrameType(const VSFrame *) { return mtVideo; }
int VS_CC getFrameWidth(const VSFrame *f, int plane) { return f->pw(plane); }
int VS_CC getFrameHeight(const VSFrame *f, int plane) { return f->ph(plane); }

int VS_CC queryVideoFormat(VSVideoFormat *format, int colorFamily, int sampleType, int bitsPerSample, int subSamplingW, int subSamplingH, VSCore *) {
    if (colorFamily != cfGray && colorFamily != cfYUV && colorFamily != cfRGB)
        return 0;
    if (bitsPerSample < 8 || bitsPerSample > 32)
        return 0;
    format->colorFamily = colorFamily;
    format->sampleType = sampleType;
    format->bitsPerSample = bitsPerSample;
    format->bytesPerSample = bitsPerSample <= 8 ? 1 : (bitsPerSample <= 16 ? 2 : 4);
    format->subSamplingW = subSamplingW;
    format->subSamplingH = subSamplingH;
    format->numPlanes = colorFamily == cfGray ? 1 : 3;
    return 1;
}

// inverse of VS_MAKE_VIDEO_ID (VapourSynth4.h)
int VS_CC getVideoFormatByID(VSVideoFormat *format, uint32_t id, VSCore *core) {
    return queryVideoFormat(format, (id >> 28) & 0xF, (id >> 24) & 0xF, (id >> 16) & 0xFF, (id >> 8) & 0xFF, id & 0xFF, core);
}

// ------------------------------------------------------------------- nodes

void VS_CC freeNode(VSNode *node) {
    if (!node)
        return;
    if (node->refs.fetch_sub(1) == 1) {
        if (node->freeFn)
            node->freeFn(node->instanceData, node->core, g_api());
        for (auto &e : node->lru)
            freeFrame(e.second);
        for (auto f : node->sourceFrames)
            freeFrame(f);
        delete node;
    }
}

VSNode *VS_CC addNodeRef(VSNode *node) {
    node->refs.fetch_add(1);
    return node;
}

int VS_CC getNodeType(VSNode *) { return mtVideo; }
const VSVideoInfo *VS_CC getVideoInfo(VSNode *node) { return &node->vi; }

int VS_CC mapSetNode(VSMap *map, const char *key, VSNode *node, int append);

VSNode *VS_CC createVideoFilter2(const char *name, const VSVideoInfo *vi, VSFilterGetFrame getFrame, VSFilterFree free, int, const VSFilterDependency *, int, void *instanceData, VSCore *core) {
    VSNode *n = new VSNode;
    n->core = core;
    n->name = name;
    n->vi = *vi;
    n->getFrame = getFrame;
    n->freeFn = free;
    n->instanceData = instanceData;
    return n;
}

void VS_CC createVideoFilter(VSMap *out, const char *name, const VSVideoInfo *vi, VSFilterGetFrame getFrame, VSFilterFree free, int filterMode, const VSFilterDependency *deps, int numDeps, void *instanceData, VSCore *core) {
    VSNode *n = createVideoFilter2(name, vi, getFrame, free, filterMode, deps, numDeps, instanceData, core);
    mapSetNode(out, "clip", n, maAppend);
    freeNode(n);
}

// frame evaluation ---------------------------------------------------------

const VSFrame *evalFrame(VSNode *node, int n, std::string &err);

const VSFrame *computeFrame(VSNode *node, int n, std::string &err) {
    if (!node->getFrame) { // source clip
        int idx = n < 0 ? 0 : (n >= (int)node->sourceFrames.size() ? (int)node->sourceFrames.size() - 1 : n);
        const auto slow = node->sourceDelayMs.find(idx);
        if (slow != node->sourceDelayMs.end())
            std::this_thread::sleep_for(std::chrono::milliseconds(slow->second));
        return addFrameRef(node->sourceFrames[idx]);
    }
    VSFrameContext ctx;
    void *frameData[4] = {nullptr, nullptr, nullptr, nullptr}; // (API 4: four pointers per request, zeroed)
    const VSFrame *r = node->getFrame(n, arInitial, node->instanceData, frameData, &ctx, node->core, g_api());
    if (ctx.hasError) {
        err = ctx.error;
        if (r)
            freeFrame(r);
        return nullptr;
    }
    if (r)
        return r;
    bool failed = false;
    for (auto &rq : ctx.requests) {
        if (ctx.ready.count(rq))
            continue;
        const VSFrame *f = evalFrame(rq.first, rq.second, err);
        if (!f) {
            failed = true;
            break;
        }
        ctx.ready[rq] = f;
    }
    if (!failed) {
        r = node->getFrame(n, arAllFramesReady, node->instanceData, frameData, &ctx, node->core, g_api());
        if (ctx.hasError) {
            err = ctx.error;
            if (r)
                freeFrame(r);
            r = nullptr;
        } else if (!r) {
            err = node->name + ": getFrame returned no frame and no error";
        }
    }
    if (failed) // a requested frame failed: the filter gets the chance to release what it keeps in frameData
        node->getFrame(n, arError, node->instanceData, frameData, &ctx, node->core, g_api());
    for (auto &e : ctx.ready)
        freeFrame(e.second);
    node->computed.fetch_add(1);
    return r;
}

const VSFrame *evalFrame(VSNode *node, int n, std::string &err) {
    if (n < 0 || n >= node->vi.numFrames) { // VapourSynth clamps out-of-range requests
        if (n < 0)
            n = 0;
        if (n >= node->vi.numFrames)
            n = node->vi.numFrames - 1;
    }
    {
        std::unique_lock<std::mutex> lk(node->m);
        for (;;) {
            for (auto it = node->lru.begin(); it != node->lru.end(); ++it)
                if (it->first == n) {
                    node->lru.splice(node->lru.begin(), node->lru, it);
                    return addFrameRef(it->second);
                }
            if (node->inflight.count(n))
                node->cv.wait(lk);
            else
                break;
        }
        node->inflight[n] = 1;
    }
    const VSFrame *f = computeFrame(node, n, err);
    {
        std::unique_lock<std::mutex> lk(node->m);
        node->inflight.erase(n);
        if (f && node->cacheSize > 0) {
            node->lru.emplace_front(n, addFrameRef(f));
            while (node->lru.size() > node->cacheSize) {
                freeFrame(node->lru.back().second);
                node->lru.pop_back();
            }
        }
    }
    node->cv.notify_all();
    return f;
}

const VSFrame *VS_CC getFrame(int n, VSNode *node, char *errorMsg, int bufSize) {
    std::string err;
    const VSFrame *f = evalFrame(node, n, err);
    if (!f && errorMsg && bufSize > 0)
        snprintf(errorMsg, (size_t)bufSize, "%s", err.c_str());
    return f;
}

const VSFrame *VS_CC getFrameFilter(int n, VSNode *node, VSFrameContext *ctx) {
    if (n < 0)
        n = 0;
    if (n >= node->vi.numFrames)
        n = node->vi.numFrames - 1;
    auto it = ctx->ready.find({node, n});
    if (it == ctx->ready.end())
        return nullptr;
    return addFrameRef(it->second);
}

void VS_CC requestFrameFilter(int n, VSNode *node, VSFrameContext *ctx) {
    if (n < 0)
        n = 0;
    if (n >= node->vi.numFrames)
        n = node->vi.numFrames - 1;
    ctx->requests.emplace_back(node, n);
}

void VS_CC setFilterError(const char *msg, VSFrameContext *ctx) {
    ctx->hasError = true;
    ctx->error = msg ? msg : "";
}

// -------------------------------------------------------------------- maps

VSMap *VS_CC createMap(void) { return new VSMap; }
void VS_CC freeMap(VSMap *m) { delete m; }
void VS_CC clearMap(VSMap *m) { m->clear(); }

void VS_CC mapSetError(VSMap *m, const char *msg) {
    m->clear();
    m->hasError = true;
    m->error = msg ? msg : "";
}

const char *VS_CC mapGetError(const VSMap *m) { return m->hasError ? m->error.c_str() : nullptr; }

int VS_CC mapNumKeys(const VSMap *m) { return (int)m->kv.size(); }

const char *VS_CC mapGetKey(const VSMap *m, int index) {
    auto it = m->kv.begin();
    std::advance(it, index);
    return it->first.c_str();
}

int VS_CC mapDeleteKey(VSMap *m, const char *key) {
    auto it = m->kv.find(key);
    if (it == m->kv.end())
        return 0;
    for (auto &v : it->second) {
        if (v.node)
            freeNode(v.node);
        if (v.frame)
            freeFrame(v.frame);
    }
    m->kv.erase(it);
    return 1;
}

int VS_CC mapNumElements(const VSMap *m, const char *key) {
    auto it = m->kv.find(key);
    return it == m->kv.end() ? -1 : (int)it->second.size();
}

int VS_CC mapGetType(const VSMap *m, const char *key) {
    auto it = m->kv.find(key);
    return (it == m->kv.end() || it->second.empty()) ? ptUnset : it->second[0].type;
}

const Value *lookup(const VSMap *m, const char *key, int index, int type, int *error) {
    int e = peSuccess;
    const Value *r = nullptr;
    if (m->hasError) {
        e = peError;
    } else {
        auto it = m->kv.find(key);
        if (it == m->kv.end())
            e = peUnset;
        else if (index < 0 || index >= (int)it->second.size())
            e = peIndex;
        else if (it->second[index].type != type)
            e = peType;
        else
            r = &it->second[index];
    }
    if (error)
        *error = e;
    else if (e != peSuccess) {
        fprintf(stderr, "minivs: fatal: property read of '%s' failed with %d and no error pointer\n", key, e);
        abort();
    }
    return r;
}

Value &store(VSMap *m, const char *key, int append) {
    auto &vec = m->kv[key];
    if (append == maReplace) {
        for (auto &v : vec) {
            if (v.node)
                freeNode(v.node);
            if (v.frame)
                freeFrame(v.frame);
        }
        vec.clear();
    }
    vec.emplace_back();
    return vec.back();
}

int64_t VS_CC mapGetInt(const VSMap *m, const char *key, int index, int *error) {
    const Value *v = lookup(m, key, index, ptInt, error);
    return v ? v->i : 0;
}

int VS_CC mapGetIntSaturated(const VSMap *m, const char *key, int index, int *error) {
    int64_t v = mapGetInt(m, key, index, error);
    if (v > INT32_MAX)
        return INT32_MAX;
    if (v < INT32_MIN)
        return INT32_MIN;
    return (int)v;
}

int VS_CC mapSetInt(VSMap *m, const char *key, int64_t i, int append) {
    Value &v = store(m, key, append);
    v.type = ptInt;
    v.i = i;
    return 0;
}

double VS_CC mapGetFloat(const VSMap *m, const char *key, int index, int *error) {
    const Value *v = lookup(m, key, index, ptFloat, error);
    return v ? v->f : 0.0;
}

float VS_CC mapGetFloatSaturated(const VSMap *m, const char *key, int index, int *error) {
    return (float)mapGetFloat(m, key, index, error);
}

int VS_CC mapSetFloat(VSMap *m, const char *key, double d, int append) {
    Value &v = store(m, key, append);
    v.type = ptFloat;
    v.f = d;
    return 0;
}

const char *VS_CC mapGetData(const VSMap *m, const char *key, int index, int *error) {
    const Value *v = lookup(m, key, index, ptData, error);
    return v ? v->data.data() : nullptr;
}

int VS_CC mapGetDataSize(const VSMap *m, const char *key, int index, int *error) {
    const Value *v = lookup(m, key, index, ptData, error);
    return v ? (int)v->data.size() : -1;
}

int VS_CC mapGetDataTypeHint(const VSMap *m, const char *key, int index, int *error) {
    const Value *v = lookup(m, key, index, ptData, error);
    return v ? v->hint : dtUnknown;
}

int VS_CC mapSetData(VSMap *m, const char *key, const char *data, int size, int type, int append) {
    Value &v = store(m, key, append);
    v.type = ptData;
    v.hint = type;
    if (size < 0)
        v.data = data;
    else
        v.data.assign(data, (size_t)size);
    return 0;
}

VSNode *VS_CC mapGetNode(const VSMap *m, const char *key, int index, int *error) {
    const Value *v = lookup(m, key, index, ptVideoNode, error);
    return v ? addNodeRef(v->node) : nullptr;
}

int VS_CC mapSetNode(VSMap *m, const char *key, VSNode *node, int append) {
    Value &v = store(m, key, append);
    v.type = ptVideoNode;
    v.node = addNodeRef(node);
    return 0;
}

int VS_CC mapConsumeNode(VSMap *m, const char *key, VSNode *node, int append) {
    int r = mapSetNode(m, key, node, append);
    freeNode(node);
    return r;
}

const VSFrame *VS_CC mapGetFrame(const VSMap *m, const char *key, int index, int *error) {
    const Value *v = lookup(m, key, index, ptVideoFrame, error);
    return v ? addFrameRef(v->frame) : nullptr;
}

int VS_CC mapSetFrame(VSMap *m, const char *key, const VSFrame *f, int append) {
    Value &v = store(m, key, append);
    v.type = ptVideoFrame;
    v.frame = addFrameRef(f);
    return 0;
}

// ----------------------------------------------------------------- plugins

int VS_CC pluginGetAPIVersion(void) { return VAPOURSYNTH_API_VERSION; }

int VS_CC configPlugin(const char *identifier, const char *pluginNamespace, const char *name, int pluginVersion, int apiVersion, int, VSPlugin *plugin) {
    plugin->id = identifier;
    plugin->ns = pluginNamespace;
    plugin->name = name;
    plugin->version = pluginVersion;
    plugin->apiVersion = apiVersion;
    return 1;
}

int VS_CC registerFunction(const char *name, const char *args, const char *returnType, VSPublicFunction argsFunc, void *functionData, VSPlugin *plugin) {
    VSPluginFunction f;
    f.name = name;
    f.args = args;
    f.ret = returnType;
    f.fn = argsFunc;
    f.userData = functionData;
    plugin->funcs[name] = f;
    return 1;
}

const VSPLUGINAPI g_pluginApi = {pluginGetAPIVersion, configPlugin, registerFunction};

VSMap *VS_CC invoke(VSPlugin *plugin, const char *name, const VSMap *args) {
    VSMap *out = new VSMap;
    auto it = plugin->funcs.find(name);
    if (it == plugin->funcs.end()) {
        mapSetError(out, (std::string("minivs: no function named ") + name).c_str());
        return out;
    }
    // Reject argument names the function did not register (what VapourSynth does).
    for (auto &e : args->kv) {
        std::string needle = e.first + ":";
        const std::string &a = it->second.args;
        size_t pos = a.find(needle);
        bool ok = false;
        while (pos != std::string::npos) {
            if (pos == 0 || a[pos - 1] == ';') {
                ok = true;
                break;
            }
            pos = a.find(needle, pos + 1);
        }
        if (!ok) {
            mapSetError(out, (std::string(name) + ": Function does not take argument(s) named " + e.first).c_str());
            return out;
        }
    }
    it->second.fn(args, out, it->second.userData, plugin->core, g_api());
    return out;
}

VSPlugin *VS_CC getPluginByID(const char *identifier, VSCore *core) {
    for (VSPlugin *p : core->plugins)
        if (p->id == identifier)
            return p;
    return nullptr;
}

int VS_CC apiGetAPIVersion(void) { return VAPOURSYNTH_API_VERSION; }

void VS_CC logMessage(int msgType, const char *msg, VSCore *) {
    fprintf(stderr, "minivs[%d]: %s\n", msgType, msg);
    if (msgType == mtFatal)
        abort();
}

VSAPI makeApi() {
    VSAPI a;
    memset(&a, 0, sizeof(a));
    a.createVideoFilter = createVideoFilter;
    a.createVideoFilter2 = createVideoFilter2;
    a.freeNode = freeNode;
    a.addNodeRef = addNodeRef;
    a.getNodeType = getNodeType;
    a.getVideoInfo = getVideoInfo;
    a.newVideoFrame = newVideoFrame;
    a.freeFrame = freeFrame;
    a.addFrameRef = addFrameRef;
    a.copyFrame = copyFrame;
    a.getFramePropertiesRO = getFramePropertiesRO;
    a.getFramePropertiesRW = getFramePropertiesRW;
    a.getStride = getStride;
    a.getReadPtr = getReadPtr;
    a.getWritePtr = getWritePtr;
    a.getVideoFrameFormat = getVideoFrameFormat;
    a.getFrameType = getFrameType;
    a.getFrameWidth = getFrameWidth;
    a.getFrameHeight = getFrameHeight;
    a.queryVideoFormat = queryVideoFormat;
    a.getVideoFormatByID = getVideoFormatByID;
    a.getPluginByID = getPluginByID;
    a.getFrame = getFrame;
    a.getFrameFilter = getFrameFilter;
    a.requestFrameFilter = requestFrameFilter;
    a.setFilterError = setFilterError;
    a.createMap = createMap;
    a.freeMap = freeMap;
    a.clearMap = clearMap;
    a.copyMap = copyMap;
    a.mapSetError = mapSetError;
    a.mapGetError = mapGetError;
    a.mapNumKeys = mapNumKeys;
    a.mapGetKey = mapGetKey;
    a.mapDeleteKey = mapDeleteKey;
    a.mapNumElements = mapNumElements;
    a.mapGetType = mapGetType;
    a.mapGetInt = mapGetInt;
    a.mapGetIntSaturated = mapGetIntSaturated;
    a.mapSetInt = mapSetInt;
    a.mapGetFloat = mapGetFloat;
    a.mapGetFloatSaturated = mapGetFloatSaturated;
    a.mapSetFloat = mapSetFloat;
    a.mapGetData = mapGetData;
    a.mapGetDataSize = mapGetDataSize;
    a.mapGetDataTypeHint = mapGetDataTypeHint;
    a.mapSetData = mapSetData;
    a.mapGetNode = mapGetNode;
    a.mapSetNode = mapSetNode;
    a.mapConsumeNode = mapConsumeNode;
    a.mapGetFrame = mapGetFrame;
    a.mapSetFrame = mapSetFrame;
    a.registerFunction = registerFunction;
    a.invoke = invoke;
    a.getAPIVersion = apiGetAPIVersion;
    a.logMessage = logMessage;
    return a;
}

} // namespace

static const VSAPI *g_api() {
    static const VSAPI api = makeApi();
    return &api;
}

VS_EXTERNAL_API(const VSAPI *) getVapourSynthAPI(int) { return g_api(); }

// ===================================================================== C API
// (bound with ctypes from tests/ and bench.py's cpu_baseline leg)

static thread_local std::string t_err;

#define MVS_API extern "C" __attribute__((visibility("default")))

MVS_API const char *mvs_last_error() { return t_err.c_str(); }


// ------------------------------------------------ built-in std.AssumeFPS
// mv.FlowFPS / mv.BlockFPS finish by invoking std.AssumeFPS (MVFlowFPS.c:883-910): a pass-through that changes the
// frame rate in the video info and stamps _DurationNum / _DurationDen on every frame, as VapourSynth's does.
namespace {
struct AssumeFPSData {
    VSNode *node;
    VSVideoInfo vi;
};
const VSFrame *VS_CC assumeFpsGetFrame(int n, int activationReason, void *instanceData, void **, VSFrameContext *frameCtx, VSCore *core, const VSAPI *vsapi) {
    AssumeFPSData *d = (AssumeFPSData *)instanceData;
    if (activationReason == arInitial) {
        vsapi->requestFrameFilter(n, d->node, frameCtx);
    } else if (activationReason == arAllFramesReady) {
        const VSFrame *src = vsapi->getFrameFilter(n, d->node, frameCtx);
        VSFrame *dst = vsapi->copyFrame(src, core);
        vsapi->freeFrame(src);
        VSMap *props = vsapi->getFramePropertiesRW(dst);
        vsapi->mapSetInt(props, "_DurationNum", d->vi.fpsDen, maReplace);
        vsapi->mapSetInt(props, "_DurationDen", d->vi.fpsNum, maReplace);
        return dst;
    }
    return nullptr;
}
void VS_CC assumeFpsFree(void *instanceData, VSCore *, const VSAPI *vsapi) {
    AssumeFPSData *d = (AssumeFPSData *)instanceData;
    vsapi->freeNode(d->node);
    delete d;
}
void VS_CC assumeFpsCreate(const VSMap *in, VSMap *out, void *, VSCore *core, const VSAPI *vsapi) {
    int err = 0;
    AssumeFPSData *d = new AssumeFPSData;
    d->node = vsapi->mapGetNode(in, "clip", 0, nullptr);
    d->vi = *vsapi->getVideoInfo(d->node);
    int64_t num = vsapi->mapGetInt(in, "fpsnum", 0, &err);
    int64_t den = vsapi->mapGetInt(in, "fpsden", 0, &err);
    if (err)
        den = 1;
    if (num <= 0 || den <= 0) {
        vsapi->mapSetError(out, "AssumeFPS: invalid framerate specified");
        vsapi->freeNode(d->node);
        delete d;
        return;
    }
    d->vi.fpsNum = num;
    d->vi.fpsDen = den;
    VSFilterDependency deps[1] = {{d->node, rpStrictSpatial}};
    vsapi->createVideoFilter(out, "AssumeFPS", &d->vi, assumeFpsGetFrame, assumeFpsFree, fmParallel, deps, 1, d, core);
}
VSPlugin *makeStdPlugin(VSCore *core) {
    VSPlugin *p = new VSPlugin;
    p->core = core;
    configPlugin("com.vapoursynth.std", "std", "minivs built-ins", 1, VAPOURSYNTH_API_VERSION, 0, p);
    registerFunction("AssumeFPS", "clip:vnode;src:vnode:opt;fpsnum:int:opt;fpsden:int:opt;", "clip:vnode;", assumeFpsCreate, nullptr, p);
    return p;
}
} // namespace

MVS_API void *mvs_core_create() {
    // VapourSynth recycles frame memory through a pool; the nearest plain-malloc equivalent: keep freed frame buffers in
    // the heap instead of returning every multi-megabyte frame to the kernel (mmap / munmap and page faults per frame)
    mallopt(M_MMAP_THRESHOLD, 1 << 30);
    mallopt(M_TRIM_THRESHOLD, 1 << 30);
    VSCore *core = new VSCore;
    core->plugins.push_back(makeStdPlugin(core));
    return core;
}

MVS_API void mvs_core_free(void *core) { delete (VSCore *)core; }

MVS_API void *mvs_load_plugin(void *core, const char *path) {
    void *h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) {
        t_err = dlerror();
        return nullptr;
    }
    auto init = (VSInitPlugin)dlsym(h, "VapourSynthPluginInit2");
    if (!init) {
        t_err = std::string(path) + ": VapourSynthPluginInit2 not exported";
        return nullptr;
    }
    VSPlugin *p = new VSPlugin;
    p->core = (VSCore *)core;
    p->handle = h;
    p->path = path;
    init(p, &g_pluginApi);
    ((VSCore *)core)->plugins.push_back(p);
    return p;
}

MVS_API const char *mvs_plugin_info(void *plugin, int what) {
    VSPlugin *p = (VSPlugin *)plugin;
    switch (what) {
    case 0: return p->id.c_str();
    case 1: return p->ns.c_str();
    case 2: return p->name.c_str();
    }
    return "";
}

MVS_API int mvs_plugin_num_functions(void *plugin) { return (int)((VSPlugin *)plugin)->funcs.size(); }

MVS_API const char *mvs_plugin_function_name(void *plugin, int i) {
    auto &f = ((VSPlugin *)plugin)->funcs;
    auto it = f.begin();
    std::advance(it, i);
    return it->first.c_str();
}

MVS_API const char *mvs_plugin_function_args(void *plugin, const char *name) {
    auto &f = ((VSPlugin *)plugin)->funcs;
    auto it = f.find(name);
    return it == f.end() ? nullptr : it->second.args.c_str();
}

MVS_API void *mvs_map_create() { return new VSMap; }
MVS_API void mvs_map_free(void *m) { delete (VSMap *)m; }
MVS_API void mvs_map_set_int(void *m, const char *k, int64_t v) { mapSetInt((VSMap *)m, k, v, maReplace); }
MVS_API void mvs_map_set_float(void *m, const char *k, double v) { mapSetFloat((VSMap *)m, k, v, maReplace); }
MVS_API void mvs_map_set_node(void *m, const char *k, void *node) { mapSetNode((VSMap *)m, k, (VSNode *)node, maReplace); }
MVS_API const char *mvs_map_error(void *m) { return mapGetError((VSMap *)m); }

MVS_API void *mvs_map_get_node(void *m, const char *k) {
    int err = 0;
    return mapGetNode((VSMap *)m, k, 0, &err);
}

MVS_API void *mvs_invoke(void *plugin, const char *name, void *args) { return invoke((VSPlugin *)plugin, name, (VSMap *)args); }

MVS_API void *mvs_source_create(void *core, int colorFamily, int bits, int ssw, int ssh, int width, int height, int numFrames) {
    VSNode *n = new VSNode;
    n->core = (VSCore *)core;
    n->name = "source";
    if (!queryVideoFormat(&n->vi.format, colorFamily, stInteger, bits, ssw, ssh, nullptr)) {
        delete n;
        t_err = "bad format";
        return nullptr;
    }
    n->vi.fpsNum = 25;
    n->vi.fpsDen = 1;
    n->vi.width = width;
    n->vi.height = height;
    n->vi.numFrames = numFrames;
    n->cacheSize = 0;
    for (int i = 0; i < numFrames; i++) {
        VSFrame *f = allocFrame(&n->vi.format, width, height);
        for (int p = 0; p < f->fmt.numPlanes; p++)
            memset(f->plane[p], 0, (size_t)f->stride[p] * f->ph(p));
        n->sourceFrames.push_back(f);
    }
    return n;
}

MVS_API void mvs_source_set_plane(void *node, int n, int plane, const uint8_t *data, int64_t stride) {
    VSFrame *f = const_cast<VSFrame *>(((VSNode *)node)->sourceFrames[n]);
    size_t row = (size_t)f->pw(plane) * f->fmt.bytesPerSample;
    for (int y = 0; y < f->ph(plane); y++)
        memcpy(f->plane[plane] + (size_t)y * f->stride[plane], data + (size_t)y * stride, row);
}

// frame n of a source clip becomes another reference to frame m (long cyclic clips without the memory)
MVS_API void mvs_source_alias(void *node, int n, int m) {
    VSNode *nd = (VSNode *)node;
    if (n == m)
        return;
    freeFrame(nd->sourceFrames[n]);
    nd->sourceFrames[n] = addFrameRef(nd->sourceFrames[m]);
}

MVS_API void mvs_source_set_prop_int(void *node, int n, const char *key, int64_t v) {
    VSFrame *f = const_cast<VSFrame *>(((VSNode *)node)->sourceFrames[n]);
    mapSetInt(&f->props, key, v, maReplace);
}

// (set before any frame is requested)
MVS_API void mvs_source_set_delay(void *node, int n, int ms) { ((VSNode *)node)->sourceDelayMs[n] = ms; }

MVS_API void mvs_node_free(void *node) { freeNode((VSNode *)node); }
MVS_API void mvs_node_set_cache(void *node, int size) { ((VSNode *)node)->cacheSize = (size_t)size; }
MVS_API long mvs_node_computed(void *node) { return ((VSNode *)node)->computed.load(); }

// out[0..8] = width, height, numFrames, colorFamily, bits, bytes, ssw, ssh, numPlanes
MVS_API void mvs_node_info(void *node, int *out) {
    const VSVideoInfo &vi = ((VSNode *)node)->vi;
    out[0] = vi.width;
    out[1] = vi.height;
    out[2] = vi.numFrames;
    out[3] = vi.format.colorFamily;
    out[4] = vi.format.bitsPerSample;
    out[5] = vi.format.bytesPerSample;
    out[6] = vi.format.subSamplingW;
    out[7] = vi.format.subSamplingH;
    out[8] = vi.format.numPlanes;
}

MVS_API const void *mvs_get_frame(void *node, int n) {
    std::string err;
    const VSFrame *f = evalFrame((VSNode *)node, n, err);
    if (!f)
        t_err = err;
    return f;
}

MVS_API void mvs_frame_free(const void *f) { freeFrame((const VSFrame *)f); }
MVS_API int mvs_frame_width(const void *f, int p) { return ((const VSFrame *)f)->pw(p); }
MVS_API int mvs_frame_height(const void *f, int p) { return ((const VSFrame *)f)->ph(p); }
MVS_API int64_t mvs_frame_stride(const void *f, int p) { return ((const VSFrame *)f)->stride[p]; }
MVS_API const uint8_t *mvs_frame_plane(const void *f, int p) { return ((const VSFrame *)f)->plane[p]; }

MVS_API int mvs_frame_prop_int(const void *f, const char *key, int64_t *out) {
    int err = 0;
    int64_t v = mapGetInt(&((const VSFrame *)f)->props, key, 0, &err);
    if (err)
        return 0;
    *out = v;
    return 1;
}

MVS_API int mvs_frame_prop_data_size(const void *f, const char *key) {
    int err = 0;
    int s = mapGetDataSize(&((const VSFrame *)f)->props, key, 0, &err);
    return err ? -1 : s;
}

MVS_API const char *mvs_frame_prop_data(const void *f, const char *key) {
    int err = 0;
    return mapGetData(&((const VSFrame *)f)->props, key, 0, &err);
}

// Pull frames [n0, n1) of `node` with `threads` host threads (round-robin, the way
// VapourSynth's fmParallel pool keeps `threads` requests in flight).  Returns the
// wall-clock seconds, or a negative number on error.  Frames are dropped.
MVS_API double mvs_render(void *node, int n0, int n1, int threads) {
    std::atomic<int> next{n0};
    std::atomic<bool> bad{false};
    std::string firstErr;
    std::mutex em;
    auto t0 = std::chrono::steady_clock::now();
    auto work = [&]() {
        for (;;) {
            int n = next.fetch_add(1);
            if (n >= n1 || bad.load())
                return;
            std::string err;
            const VSFrame *f = evalFrame((VSNode *)node, n, err);
            if (!f) {
                std::lock_guard<std::mutex> g(em);
                if (!bad.exchange(true))
                    firstErr = err;
                return;
            }
            freeFrame(f);
        }
    };
    std::vector<std::thread> pool;
    for (int i = 0; i < threads; i++)
        pool.emplace_back(work);
    for (auto &t : pool)
        t.join();
    auto t1 = std::chrono::steady_clock::now();
    if (bad.load()) {
        t_err = firstErr;
        return -1.0;
    }
    return std::chrono::duration<double>(t1 - t0).count();
}
