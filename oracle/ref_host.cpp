// ref_host.cpp -- TEST INFRASTRUCTURE (parity oracle), not part of the shipped product.
//
// A fake Arnold host for the UNMODIFIED reference plugin.  oracle/Makefile compiles
// /root/reference/src/zoic.cpp, untouched and from where it lies, into
// oracle/_ref/libzoic_ref.so against include/arnold_shim/ai.h; this file becomes
// oracle/_ref/libzoic_refhost.so and provides
//   * the 18 Ai* host functions the plugin calls,
//   * an interposed `uint32_t xor128()` (the plugin's calls go through the PLT, so a definition
//     that sits earlier in the global lookup scope wins -- load this library with RTLD_GLOBAL
//     BEFORE libzoic_ref.so): setup draws come from a fresh canonical-seed stream, and every
//     camera sample gets its own retry stream seeded from (seed, sample index),
//   * a small C API (zref_*) that drives NodeLoader -> Initialize -> Update -> CreateRay x N -> Finish.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
#include <ai.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <map>
#include <string>
#include <thread>
#include <vector>

// ------------------------------------------------------------------------------------------------
// Zero-filling replacement of the global allocation functions.  The reference never initialises
// Lensdata::apertureElement for lens tables without a stop row (src/zoic.cpp:532, set only at :922), so what
// it reads is whatever `new cameraData()` happens to return.  The ruling is "0"; this
// replaceable operator new makes the unmodified plugin behave that way deterministically.
// ------------------------------------------------------------------------------------------------
#include <new>
void* operator new(size_t n) {
    void* p = calloc(1, n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void* operator new[](size_t n) { return operator new(n); }
void operator delete(void* p) noexcept { free(p); }
void operator delete[](void* p) noexcept { free(p); }
void operator delete(void* p, size_t) noexcept { free(p); }
void operator delete[](void* p, size_t) noexcept { free(p); }

// ------------------------------------------------------------------------------------------------
// RNG interposer
// ------------------------------------------------------------------------------------------------
namespace {
struct XorState { uint32_t x, y, z, w; };
const XorState kCanonical = {123456789u, 362436069u, 521288629u, 88675123u};  // reference seed constants
thread_local XorState g_rng = kCanonical;
thread_local uint64_t g_draws = 0;

inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
const uint64_t kGolden = 0x9E3779B97F4A7C15ull;

// per-sample retry stream (repo convention, DESIGN.md "retry RNG"): two SplitMix64 outputs -> 128-bit state
inline XorState sample_stream(uint64_t seed, uint64_t index) {
    uint64_t h0 = mix64(seed + kGolden * (index + 1));
    uint64_t h1 = mix64(h0 + kGolden);
    XorState s;
    s.x = (uint32_t)h0;
    s.y = (uint32_t)(h0 >> 32);
    s.z = (uint32_t)h1;
    s.w = (uint32_t)(h1 >> 32) | 1u;
    return s;
}
}  // namespace

// Same recurrence as the plugin's own generator (Marsaglia xorshift128); state is ours.
uint32_t xor128(void) {
    XorState& s = g_rng;
    uint32_t t = s.x ^ (s.x << 11);
    s.x = s.y; s.y = s.z; s.z = s.w;
    ++g_draws;
    return s.w = (s.w ^ (s.w >> 19) ^ t ^ (t >> 8));
}

// ------------------------------------------------------------------------------------------------
// fake node + host functions
// ------------------------------------------------------------------------------------------------
struct AtNode {
    std::map<std::string, float> flt;      // floats, bools (0/1) and the lensModel enum live here
    std::map<std::string, std::string> str;
    void* local = nullptr;
    // in-memory bokeh texture served by AiTexture*
    std::vector<float> image;
    unsigned iw = 0, ih = 0, inch = 0;
};

namespace {
thread_local AtNode* g_current = nullptr;      // node whose texture AiTexture* serves
thread_local std::string* g_log = nullptr;     // captured AiMsg* output
thread_local int g_abort = 0;
int g_verbose = 0;

void vlog(const char* tag, const char* fmt, va_list ap) {
    char buf[1024];
    vsnprintf(buf, sizeof buf, fmt, ap);
    if (g_log) { *g_log += tag; *g_log += buf; *g_log += "\n"; }
    if (g_verbose) fprintf(stderr, "%s%s\n", tag, buf);
}
}  // namespace

float AiNodeGetFlt(const AtNode* n, const char* k) { auto it = n->flt.find(k); return it == n->flt.end() ? 0.0f : it->second; }
bool AiNodeGetBool(const AtNode* n, const char* k) { return AiNodeGetFlt(n, k) != 0.0f; }
int AiNodeGetInt(const AtNode* n, const char* k) { return (int)AiNodeGetFlt(n, k); }
AtString AiNodeGetStr(const AtNode* n, const char* k) {
    auto it = n->str.find(k);
    return AtString(it == n->str.end() ? "" : it->second.c_str());
}
void AiNodeSetLocalData(AtNode* n, void* d) { n->local = d; }
void* AiNodeGetLocalData(const AtNode* n) { return n->local; }
void AiCameraInitialize(AtNode*) {}
void AiCameraUpdate(AtNode*, bool) {}
void AiMsgInfo(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vlog("I ", fmt, ap); va_end(ap); }
void AiMsgWarning(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vlog("W ", fmt, ap); va_end(ap); }
void AiMsgError(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vlog("E ", fmt, ap); va_end(ap); }
void AiRenderAbort() { g_abort = 1; }
void* AiMalloc(size_t b) { return malloc(b); }
void AiFree(void* p) { free(p); }
void AiAddMemUsage(int64_t, const AtString) {}
bool AiTextureGetResolution(const AtString, unsigned* w, unsigned* h) {
    if (!g_current || g_current->image.empty()) return false;
    *w = g_current->iw; *h = g_current->ih; return true;
}
bool AiTextureGetNumChannels(const AtString, unsigned* c) {
    if (!g_current || g_current->image.empty()) return false;
    *c = g_current->inch; return true;
}
bool AiTextureLoad(const AtString, bool, unsigned, void* out) {
    if (!g_current || g_current->image.empty()) return false;
    memcpy(out, g_current->image.data(), g_current->image.size() * sizeof(float));
    return true;
}
void AiShimDeclareFlt(AtList*, const char*, float) {}
void AiShimDeclareBool(AtList*, const char*, bool) {}
void AiShimDeclareStr(AtList*, const char*, const char*) {}
void AiShimDeclareEnum(AtList*, const char*, int, const char**) {}

// ------------------------------------------------------------------------------------------------
// C driver API (ctypes-friendly)
// ------------------------------------------------------------------------------------------------
extern "C" {

// mirrors the 14 node parameters (reference src/zoic.cpp:1547-1562)
struct zref_params {
    float sensorWidth, sensorHeight, focalLength, fStop, focalDistance;
    int useImage;
    int lensModel;         // 0 THINLENS, 1 RAYTRACED
    int kolbSamplingLUT;
    int useDof;
    float opticalVignettingDistance, opticalVignettingRadius, exposureControl;
    const char* lensDataPath;
    const char* bokehPath;  // any non-empty label; pixels come from zref_create's image argument
};

struct zref_camera {
    AtNode node;
    std::string log;
    const AtCommonMethods* cm = nullptr;
    const AtCameraNodeMethods* dm = nullptr;
    int aborted = 0;
};

struct Plugin { void* handle; AtNodeLib lib; std::string path; };
static std::vector<Plugin> g_plugins;

void zref_set_verbose(int v) { g_verbose = v; }

// dlopen a plugin (the compiled reference, or any library exporting NodeLoader) and run its NodeLoader.
// Returns a plugin index >= 0, or a negative error.
int zref_open(const char* plugin_path) {
    for (size_t i = 0; i < g_plugins.size(); ++i)
        if (g_plugins[i].path == plugin_path) return (int)i;
    Plugin pl;
    pl.path = plugin_path;
    pl.handle = dlopen(plugin_path, RTLD_NOW | RTLD_LOCAL);
    if (!pl.handle) { fprintf(stderr, "zref_open: %s\n", dlerror()); return -1; }
    typedef bool (*loader_t)(int, AtNodeLib*);
    loader_t loader = (loader_t)dlsym(pl.handle, "NodeLoader");
    if (!loader) return -2;
    memset(&pl.lib, 0, sizeof pl.lib);
    if (!loader(0, &pl.lib)) return -3;
    AtNodeLib scratch;
    memset(&scratch, 0, sizeof scratch);
    if (loader(1, &scratch)) return -4;  // the plugin exports exactly one node
    g_plugins.push_back(pl);
    return (int)g_plugins.size() - 1;
}

const char* zref_node_name(int plugin) { return g_plugins[plugin].lib.name; }
const char* zref_node_version(int plugin) { return g_plugins[plugin].lib.version; }
int zref_node_type(int plugin) { return g_plugins[plugin].lib.node_type; }
int zref_output_type(int plugin) { return g_plugins[plugin].lib.output_type; }

zref_camera* zref_create(int plugin, const zref_params* p, const float* image, int w, int h, int nch) {
    if (plugin < 0 || plugin >= (int)g_plugins.size()) return nullptr;
    zref_camera* c = new zref_camera();
    const AtNodeMethods* m = (const AtNodeMethods*)g_plugins[plugin].lib.methods;
    c->cm = m->cmethods;
    c->dm = (const AtCameraNodeMethods*)m->dmethods;
    AtNode& n = c->node;
    n.flt["sensorWidth"] = p->sensorWidth;
    n.flt["sensorHeight"] = p->sensorHeight;
    n.flt["focalLength"] = p->focalLength;
    n.flt["fStop"] = p->fStop;
    n.flt["focalDistance"] = p->focalDistance;
    n.flt["useImage"] = p->useImage ? 1.0f : 0.0f;
    n.flt["lensModel"] = (float)p->lensModel;
    n.flt["kolbSamplingLUT"] = p->kolbSamplingLUT ? 1.0f : 0.0f;
    n.flt["useDof"] = p->useDof ? 1.0f : 0.0f;
    n.flt["opticalVignettingDistance"] = p->opticalVignettingDistance;
    n.flt["opticalVignettingRadius"] = p->opticalVignettingRadius;
    n.flt["exposureControl"] = p->exposureControl;
    n.str["lensDataPath"] = p->lensDataPath ? p->lensDataPath : "";
    n.str["bokehPath"] = p->bokehPath ? p->bokehPath : "";
    if (image && w > 0 && h > 0 && nch > 0) {
        n.image.assign(image, image + (size_t)w * h * nch);
        n.iw = w; n.ih = h; n.inch = nch;
    }
    g_current = &n;
    g_log = &c->log;
    g_abort = 0;
    g_rng = kCanonical;  // ruling: setup (exit-pupil LUT) consumes a fresh canonical-seed stream
    g_draws = 0;
    c->cm->Parameters(nullptr, nullptr);
    c->cm->Initialize(&n);
    c->cm->Update(&n);
    c->aborted = g_abort;
    g_log = nullptr;
    return c;
}

int zref_aborted(const zref_camera* c) { return c->aborted; }
const char* zref_log(const zref_camera* c) { return c->log.c_str(); }

// samples: n x (sx, sy, lensx, lensy).  Outputs: origin_w n x (ox, oy, oz, weight), dir_tries n x (dx, dy, dz, tries).
// `first_index` is the global index of samples[0] (the retry stream is seeded from seed and the global index).
// stats (may be NULL): [0] rays with weight != 0 ... exactly: [0] success, [1] vignetted (zero weight), [2] attempts.
static void generate_range(zref_camera* c, const float* samples, uint64_t begin, uint64_t end, uint64_t first_index,
                           uint64_t seed, float* origin_w, float* dir_tries, uint64_t* stats) {
    g_current = &c->node;
    g_log = nullptr;
    uint64_t ok = 0, vig = 0, attempts = 0;
    for (uint64_t i = begin; i < end; ++i) {
        AtCameraInput in;
        in.sx = samples[4 * i + 0];
        in.sy = samples[4 * i + 1];
        in.dsx = in.dsy = 0.0f;
        in.lensx = samples[4 * i + 2];
        in.lensy = samples[4 * i + 3];
        in.relative_time = 0.0f;
        AtCameraOutput out;
        memset((void*)&out, 0, sizeof out);
        out.weight = 1.0f;
        g_rng = sample_stream(seed, first_index + i);
        g_draws = 0;
        c->dm->CreateRay(&c->node, in, out, 0);
        uint64_t tries = g_draws / 2;
        origin_w[4 * i + 0] = out.origin.x;
        origin_w[4 * i + 1] = out.origin.y;
        origin_w[4 * i + 2] = out.origin.z;
        origin_w[4 * i + 3] = out.weight.r;
        dir_tries[4 * i + 0] = out.dir.x;
        dir_tries[4 * i + 1] = out.dir.y;
        dir_tries[4 * i + 2] = out.dir.z;
        dir_tries[4 * i + 3] = (float)tries;
        attempts += 1 + tries;
        if (out.weight.r == 0.0f) ++vig; else ++ok;
    }
    if (stats) { stats[0] = ok; stats[1] = vig; stats[2] = attempts; }
}

void zref_generate(zref_camera* c, const float* samples, uint64_t n, uint64_t first_index, uint64_t seed,
                   float* origin_w, float* dir_tries, uint64_t* stats) {
    generate_range(c, samples, 0, n, first_index, seed, origin_w, dir_tries, stats);
}

// The plugin "as shipped": `nthreads` render threads call CreateRay on ONE node at the same time, like Arnold's render
// threads do.  The reference's counters (succesRays, vignettedRays, ... src/zoic.cpp:533-534) are plain
// members of the node's shared data, incremented without synchronisation by every thread -- one contended cache line;
// its retry RNG is interposed per thread here, so the rays themselves still equal the single-thread run.
void zref_generate_mt(zref_camera* c, const float* samples, uint64_t n, uint64_t first_index, uint64_t seed,
                      float* origin_w, float* dir_tries, uint64_t* stats, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    std::vector<std::thread> pool;
    std::vector<uint64_t> part((size_t)nthreads * 3, 0);
    for (int t = 0; t < nthreads; ++t) {
        const uint64_t b = n * (uint64_t)t / (uint64_t)nthreads, e = n * (uint64_t)(t + 1) / (uint64_t)nthreads;
        pool.emplace_back(generate_range, c, samples, b, e, first_index, seed, origin_w, dir_tries, &part[(size_t)t * 3]);
    }
    for (auto& th : pool) th.join();
    if (stats) {
        stats[0] = stats[1] = stats[2] = 0;
        for (int t = 0; t < nthreads; ++t) for (int k = 0; k < 3; ++k) stats[k] += part[(size_t)t * 3 + k];
    }
}

// One CreateRay with an explicit retry stream made of the given draws (for pinning the argument
// evaluation order of the plugin's two-draw call sites).  draws beyond `ndraws` continue the stream.
void zref_generate_one_with_state(zref_camera* c, const float* sample, const uint32_t state[4],
                                  float* origin_w, float* dir_tries, float* derivs /*dOdy(3), dDdy(3)*/) {
    g_current = &c->node;
    AtCameraInput in;
    in.sx = sample[0]; in.sy = sample[1]; in.dsx = in.dsy = 0.0f;
    in.lensx = sample[2]; in.lensy = sample[3]; in.relative_time = 0.0f;
    AtCameraOutput out;
    memset((void*)&out, 0, sizeof out);
    out.weight = 1.0f;
    g_rng.x = state[0]; g_rng.y = state[1]; g_rng.z = state[2]; g_rng.w = state[3];
    g_draws = 0;
    c->dm->CreateRay(&c->node, in, out, 0);
    origin_w[0] = out.origin.x; origin_w[1] = out.origin.y; origin_w[2] = out.origin.z; origin_w[3] = out.weight.r;
    dir_tries[0] = out.dir.x; dir_tries[1] = out.dir.y; dir_tries[2] = out.dir.z; dir_tries[3] = (float)(g_draws / 2);
    if (derivs) {
        derivs[0] = out.dOdy.x; derivs[1] = out.dOdy.y; derivs[2] = out.dOdy.z;
        derivs[3] = out.dDdy.x; derivs[4] = out.dDdy.y; derivs[5] = out.dDdy.z;
    }
}

int zref_reverse_ray(zref_camera* c) {
    AtVector a(0, 0, 0), b(0, 0, -1);
    AtVector2 ps(0, 0);
    return c->dm->ReverseRay(&c->node, a, b, 0.0f, ps) ? 1 : 0;
}

void zref_destroy(zref_camera* c) {
    if (!c) return;
    g_current = &c->node;
    g_log = &c->log;
    c->cm->Finish(&c->node);  // deletes the plugin's local data
    g_log = nullptr;
    g_current = nullptr;
    delete c;
}

// retry-stream seeding exposed for the oracle's own tests
void zref_sample_stream(uint64_t seed, uint64_t index, uint32_t out[4]) {
    XorState s = sample_stream(seed, index);
    out[0] = s.x; out[1] = s.y; out[2] = s.z; out[3] = s.w;
}

}  // extern "C"
