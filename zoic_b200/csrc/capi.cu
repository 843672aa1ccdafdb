// capi.cu -- the C ABI declared in include/zoicb.h.
#include <cuda_runtime.h>

#include <atomic>
#include <map>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/zoicb.h"
#include "gnu_sort.h"
#include "host_setup.h"
#include "kernels.h"
#include "lens_math.cuh"

using namespace zoicb;

static_assert(sizeof(zoicb_ray) == 32 && sizeof(RayRecord) == 32, "a ray is one 32-byte record");

namespace {

thread_local std::string g_last_error;
std::atomic<uint64_t> g_launches{0};

zoicb_status fail(zoicb_status code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
zoicb_status cuda_fail(cudaError_t e, const char* what) {
    return fail(ZOICB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define ZCUDA(call, what)                                         \
    do {                                                          \
        cudaError_t e__ = (call);                                 \
        if (e__ != cudaSuccess) return cuda_fail(e__, what);      \
    } while (0)

void count_launches(int k) { if (k > 0) g_launches.fetch_add((uint64_t)k, std::memory_order_relaxed); }

}  // namespace

struct zoicb_ctx {
    int device = 0;
    int mode = ZOICB_MODE_GUARDED;
    HostCamera host;
    // device tables (bokeh)
    float* d_cdf_row = nullptr;
    int32_t* d_row_idx = nullptr;
    float* d_cdf_col = nullptr;
    uint16_t* d_rel_col = nullptr;
    uint16_t* d_row_guide = nullptr;
    uint16_t* d_col_guide = nullptr;
    float* d_dxy = nullptr;   // dx_of_col[w] then dy_of_row[h]
    DeviceStats* d_stats = nullptr;
    // guarded-mode scratch, one per stream the caller uses (stream order serialises reuse)
    std::vector<float> base_guards;
    std::mutex ws_mu;
    std::map<cudaStream_t, Workspace> workspaces;
    // host-buffer pipeline (zoicb_generate_host)
    static constexpr int kSlots = 3;
    uint64_t chunk = 0;
    cudaStream_t streams[kSlots] = {nullptr, nullptr, nullptr};
    float4* d_in[kSlots] = {nullptr, nullptr, nullptr};
    RayRecord* d_r[kSlots] = {nullptr, nullptr, nullptr};
    float4* h_in[kSlots] = {nullptr, nullptr, nullptr};   // pinned staging, only for pageable callers
    RayRecord* h_r[kSlots] = {nullptr, nullptr, nullptr};
    std::mutex host_mu;
    // optional: events recorded around the device-side bokeh table build (zoicb_build_bokeh_tables)
    cudaEvent_t bokeh_ev0 = nullptr, bokeh_ev1 = nullptr;
};

namespace {

// exit-pupil LUT candidates classified on the GPU; host replays the bbox update
bool lut_trace_gpu(void* user, const LensState& lens, const float* film_x, int n_film, const uint32_t* draws,
                   int per_film, uint8_t* accept) {
    (void)user;
    const size_t total = (size_t)n_film * per_film;
    float* d_film = nullptr;
    uint32_t* d_draws = nullptr;
    uint8_t* d_acc = nullptr;
    bool ok = false;
    int launches = 0;
    do {
        if (cudaMalloc(&d_film, n_film * sizeof(float)) != cudaSuccess) break;
        if (cudaMalloc(&d_draws, total * 2 * sizeof(uint32_t)) != cudaSuccess) break;
        if (cudaMalloc(&d_acc, total) != cudaSuccess) break;
        if (cudaMemcpy(d_film, film_x, n_film * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) break;
        if (cudaMemcpy(d_draws, draws, total * 2 * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess) break;
        if (launch_lut_trace(lens, d_film, n_film, per_film, d_draws, d_acc, nullptr, &launches) != cudaSuccess) break;
        if (cudaMemcpy(accept, d_acc, total, cudaMemcpyDeviceToHost) != cudaSuccess) break;
        ok = true;
    } while (0);
    count_launches(launches);
    cudaFree(d_film);
    cudaFree(d_draws);
    cudaFree(d_acc);
    if (!ok) cudaGetLastError();
    return ok;
}

// Image-based aperture tables built on the GPU (bokeh_build.cu).  The tables stay on the device in
// the context; a copy comes back for zoicb_get_bokeh_tables.
bool bokeh_build_gpu(void* user, const float* rgb, int w, int h, int nch, HostBokeh* out) {
    zoicb_ctx* c = static_cast<zoicb_ctx*>(user);
    const size_t np = (size_t)w * h;
    const size_t ng_row = (size_t)h + kBokehGuidePad, ng_col = (size_t)h * (w + kBokehGuidePad);
    float *d_rgb = nullptr, *d_work = nullptr, *d_total = nullptr, *d_row_mass = nullptr;
    int32_t* d_scratch = nullptr;
    bool ok = false;
    int launches = 0;
    do {
        if (cudaMalloc(&d_rgb, np * nch * sizeof(float)) != cudaSuccess) break;
        if (cudaMalloc(&d_work, np * sizeof(float)) != cudaSuccess) break;
        if (cudaMalloc(&d_scratch, np * sizeof(int32_t)) != cudaSuccess) break;
        if (cudaMalloc(&d_total, 2 * sizeof(float)) != cudaSuccess) break;
        if (cudaMalloc(&d_row_mass, h * sizeof(float)) != cudaSuccess) break;
        if (cudaMalloc(&c->d_cdf_row, h * sizeof(float)) != cudaSuccess) break;
        if (cudaMalloc(&c->d_row_idx, h * sizeof(int32_t)) != cudaSuccess) break;
        if (cudaMalloc(&c->d_cdf_col, np * sizeof(float)) != cudaSuccess) break;
        if (cudaMalloc(&c->d_rel_col, np * sizeof(uint16_t)) != cudaSuccess) break;
        if (cudaMalloc(&c->d_row_guide, ng_row * sizeof(uint16_t)) != cudaSuccess) break;
        if (cudaMalloc(&c->d_col_guide, ng_col * sizeof(uint16_t)) != cudaSuccess) break;
        if (cudaMemcpy(d_rgb, rgb, np * nch * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) break;
        if (c->bokeh_ev0) cudaEventRecord(c->bokeh_ev0, nullptr);
        if (launch_bokeh_build(d_rgb, w, h, nch, d_work, d_scratch, d_total, d_row_mass, c->d_cdf_row, c->d_row_idx,
                               c->d_cdf_col, c->d_rel_col, c->d_row_guide, c->d_col_guide, nullptr, &launches) != cudaSuccess) break;
        if (c->bokeh_ev1) { cudaEventRecord(c->bokeh_ev1, nullptr); cudaEventSynchronize(c->bokeh_ev1); }
        out->w = w; out->h = h;
        out->cdf_row.resize(h); out->row_indices.resize(h); out->cdf_column.resize(np); out->column_indices.resize(np);
        out->row_guide.resize(ng_row); out->col_guide.resize(ng_col);
        std::vector<uint16_t> rel(np);
        if (cudaMemcpy(out->cdf_row.data(), c->d_cdf_row, h * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) break;
        if (cudaMemcpy(out->row_indices.data(), c->d_row_idx, h * sizeof(int32_t), cudaMemcpyDeviceToHost) != cudaSuccess) break;
        if (cudaMemcpy(out->cdf_column.data(), c->d_cdf_col, np * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) break;
        if (cudaMemcpy(rel.data(), c->d_rel_col, np * sizeof(uint16_t), cudaMemcpyDeviceToHost) != cudaSuccess) break;
        if (cudaMemcpy(out->row_guide.data(), c->d_row_guide, ng_row * sizeof(uint16_t), cudaMemcpyDeviceToHost) != cudaSuccess) break;
        if (cudaMemcpy(out->col_guide.data(), c->d_col_guide, ng_col * sizeof(uint16_t), cudaMemcpyDeviceToHost) != cudaSuccess) break;
        {   // lens coordinates per column / per row, with the reference's operations (src/zoic.cpp:441,466,479-484)
            std::vector<float> dxy((size_t)w + h);
            for (int col = 0; col < w; ++col) dxy[col] = xmul(xdiv((float)(col - (h - 1) / 2), (float)w), 2.0f);
            for (int row = 0; row < h; ++row) dxy[(size_t)w + row] = xmul(xdiv(xmul((float)(row - (w - 1) / 2), -1.0f), (float)h), 2.0f);
            if (cudaMalloc(&c->d_dxy, dxy.size() * sizeof(float)) != cudaSuccess) break;
            if (cudaMemcpy(c->d_dxy, dxy.data(), dxy.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) break;
        }
        // the reference's columnIndices hold global pixel indices (row * width + column), src/zoic.cpp:365-391
        for (int r = 0; r < h; ++r)
            for (int k = 0; k < w; ++k) out->column_indices[(size_t)r * w + k] = r * w + (int32_t)rel[(size_t)r * w + k];
        ok = true;
    } while (0);
    count_launches(launches);
    cudaFree(d_rgb); cudaFree(d_work); cudaFree(d_scratch); cudaFree(d_total); cudaFree(d_row_mass);
    if (!ok) { cudaGetLastError(); out->w = out->h = 0; }
    return ok;
}

// Scratch for the guarded mode on `st`, large enough for n samples: room for n/24 undecided samples
// (measured rates are 1e-4 .. 1.7e-2); anything beyond the capacity is settled inline by the kernel.
cudaError_t get_workspace(zoicb_ctx* c, cudaStream_t st, uint64_t n, Workspace* out) {
    std::lock_guard<std::mutex> lock(c->ws_mu);
    Workspace& w = c->workspaces[st];
    unsigned long long want = n / 24 + 4096;
    if (want > (1ull << 27)) want = 1ull << 27;
    cudaError_t e;
    if (!w.counters) {
        if ((e = cudaMalloc(&w.counters, 4 * sizeof(unsigned long long))) != cudaSuccess) return e;
    }
    if (w.capacity < want) {
        if (w.queue) {
            if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
            cudaFree(w.queue);
            w.queue = nullptr;
            w.capacity = 0;
        }
        if ((e = cudaMalloc(&w.queue, want * sizeof(unsigned long long))) != cudaSuccess) return e;
        w.capacity = want;
    }
    *out = w;
    return cudaSuccess;
}

void free_ctx(zoicb_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (auto& kv : c->workspaces) { cudaFree(kv.second.counters); cudaFree(kv.second.queue); }
    cudaFree(c->d_cdf_row); cudaFree(c->d_row_idx); cudaFree(c->d_cdf_col); cudaFree(c->d_rel_col);
    cudaFree(c->d_row_guide); cudaFree(c->d_col_guide); cudaFree(c->d_dxy);
    cudaFree(c->d_stats);
    for (int s = 0; s < zoicb_ctx::kSlots; ++s) {
        if (c->streams[s]) cudaStreamDestroy(c->streams[s]);
        cudaFree(c->d_in[s]); cudaFree(c->d_r[s]);
        if (c->h_in[s]) cudaFreeHost(c->h_in[s]);
        if (c->h_r[s]) cudaFreeHost(c->h_r[s]);
    }
    delete c;
}

bool is_pinned_host(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

}  // namespace

extern "C" {

void zoicb_default_params(zoicb_params* p) {
    if (!p) return;
    p->sensorWidth = 3.6f; p->sensorHeight = 2.4f; p->focalLength = 2.0f; p->fStop = 4.0f;
    p->focalDistance = 100.0f; p->useImage = 0; p->lensModel = ZOICB_RAYTRACED; p->kolbSamplingLUT = 1;
    p->useDof = 1; p->opticalVignettingDistance = 0.0f; p->opticalVignettingRadius = 1.0f;
    p->exposureControl = 0.0f; p->lensDataPath = ""; p->bokehPath = "";
}

zoicb_status zoicb_create(const zoicb_params* params, const float* rgb, int width, int height, int nch, int device,
                          zoicb_ctx** out) {
    if (!params || !out) return fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_create: null argument");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(ZOICB_ERR_CUDA, "zoicb_create: no CUDA device (libzoicb has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_create: bad device index");
    ZCUDA(cudaSetDevice(device), "cudaSetDevice");
    cudaDeviceProp prop;
    ZCUDA(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties");
    if (prop.major != 10) return fail(ZOICB_ERR_CUDA, "zoicb_create: kernels are built for sm_100a only; device is " + std::string(prop.name));

    zoicb_ctx* c = new zoicb_ctx();
    c->device = device;
    std::string err;
    zoicb_status rc = build_camera(*params, rgb, width, height, nch, &c->host, &err, lut_trace_gpu, nullptr, bokeh_build_gpu, c);
    if (rc != ZOICB_OK) { free_ctx(c); return fail(rc, "zoicb_create: " + err); }

    auto bail = [&](cudaError_t ce, const char* what) { free_ctx(c); return cuda_fail(ce, what); };
    if ((e = cudaMalloc(&c->d_stats, sizeof(DeviceStats))) != cudaSuccess) return bail(e, "cudaMalloc(stats)");
    if ((e = cudaMemset(c->d_stats, 0, sizeof(DeviceStats))) != cudaSuccess) return bail(e, "cudaMemset(stats)");
    const HostBokeh& hb = c->host.bokeh;
    if (hb.valid()) {   // tables were built on the device by bokeh_build_gpu and stayed there
        BokehTables& bt = c->host.state.bokeh;
        bt.cdf_row = c->d_cdf_row; bt.row_indices = c->d_row_idx; bt.cdf_column = c->d_cdf_col; bt.rel_column = c->d_rel_col;
        bt.row_guide = c->d_row_guide; bt.col_guide = c->d_col_guide;
        bt.dx_of_col = c->d_dxy; bt.dy_of_row = c->d_dxy + hb.w;
        bt.w = hb.w; bt.h = hb.h; bt.valid = 1;
    }
    *out = c;
    return ZOICB_OK;
}

void zoicb_destroy(zoicb_ctx* ctx) { free_ctx(ctx); }

zoicb_status zoicb_set_mode(zoicb_ctx* ctx, int mode) {
    if (!ctx) return fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_set_mode: null context");
    if (mode != ZOICB_MODE_EXACT && mode != ZOICB_MODE_GUARDED) return fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_set_mode: unknown mode");
    ctx->mode = mode;
    return ZOICB_OK;
}
int zoicb_get_mode(const zoicb_ctx* ctx) { return ctx ? ctx->mode : -1; }

zoicb_status zoicb_set_guard_scale(zoicb_ctx* ctx, float scale) {
    if (!ctx || !(scale >= 0.0f)) return fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_set_guard_scale: bad argument");
    if (ctx->base_guards.empty()) {
        for (int i = 0; i < kMaxElements; ++i) {
            ctx->base_guards.push_back(ctx->host.state.lens.e[i].rim2_guard);
            ctx->base_guards.push_back(ctx->host.state.lens.e[i].dt_guard);
            ctx->base_guards.push_back(ctx->host.state.lens.e[i].miss_guard);
        }
        ctx->base_guards.push_back(ctx->host.state.thin.ov_guard);
    }
    for (int i = 0; i < kMaxElements; ++i) {
        ctx->host.state.lens.e[i].rim2_guard = ctx->base_guards[3 * i] * scale;
        ctx->host.state.lens.e[i].dt_guard = ctx->base_guards[3 * i + 1] * scale;
        ctx->host.state.lens.e[i].miss_guard = ctx->base_guards[3 * i + 2] * scale;
    }
    ctx->host.state.thin.ov_guard = ctx->base_guards[3 * kMaxElements] * scale;
    ctx->host.state.guard_scale = scale;
    return ZOICB_OK;
}

zoicb_status zoicb_generate(zoicb_ctx* ctx, const void* d_samples, uint64_t n, uint64_t first_index, uint64_t rng_seed,
                            zoicb_ray* d_rays, void* stream) {
    if (!ctx) return fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_generate: null context");
    if (n == 0) return ZOICB_OK;
    if (!d_samples || !d_rays) return fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_generate: null buffer");
    if (((uintptr_t)d_rays & 31u) || ((uintptr_t)d_samples & 15u))
        return fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_generate: samples must be 16-byte and rays 32-byte aligned");
    ZCUDA(cudaSetDevice(ctx->device), "cudaSetDevice");
    int launches = 0;
    Workspace ws = {nullptr, nullptr, 0};
    ZCUDA(get_workspace(ctx, (cudaStream_t)stream, n, &ws), "workspace");
    cudaError_t e = launch_generate(ctx->host.state, ctx->mode, (const float4*)d_samples, n, first_index, rng_seed,
                                    (RayRecord*)d_rays, ctx->d_stats, (cudaStream_t)stream, ws, &launches);
    count_launches(launches);
    if (e != cudaSuccess) return cuda_fail(e, "zoicb_generate launch");
    return ZOICB_OK;
}

zoicb_status zoicb_generate_host(zoicb_ctx* ctx, const float* h_samples, uint64_t n, uint64_t first_index,
                                 uint64_t rng_seed, zoicb_ray* h_rays) {
    if (!ctx) return fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_generate_host: null context");
    if (n == 0) return ZOICB_OK;
    if (!h_samples || !h_rays) return fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_generate_host: null buffer");
    std::lock_guard<std::mutex> lock(ctx->host_mu);
    ZCUDA(cudaSetDevice(ctx->device), "cudaSetDevice");
    constexpr int K = zoicb_ctx::kSlots;
    if (!ctx->chunk) {
        // samples per pipeline slot: 2 Mi by default (32 MiB up, 64 MiB down; small enough that the un-overlapped first
        // upload and last download do not show, profiles/r01_ab_pool2.txt); ZOICB_HOST_CHUNK_LOG2 overrides (A/B)
        const char* cl = getenv("ZOICB_HOST_CHUNK_LOG2");
        const int lg = cl ? atoi(cl) : 21;
        ctx->chunk = 1ull << (lg >= 16 && lg <= 26 ? lg : 21);
        for (int s = 0; s < K; ++s) {
            ZCUDA(cudaStreamCreateWithFlags(&ctx->streams[s], cudaStreamNonBlocking), "cudaStreamCreate");
            ZCUDA(cudaMalloc(&ctx->d_in[s], ctx->chunk * sizeof(float4)), "cudaMalloc(staging)");
            ZCUDA(cudaMalloc(&ctx->d_r[s], ctx->chunk * sizeof(RayRecord)), "cudaMalloc(staging)");
        }
    }
    const bool direct = is_pinned_host(h_samples) && is_pinned_host(h_rays);
    if (!direct && !ctx->h_in[0]) {
        for (int s = 0; s < K; ++s) {
            ZCUDA(cudaMallocHost(&ctx->h_in[s], ctx->chunk * sizeof(float4)), "cudaMallocHost");
            ZCUDA(cudaMallocHost(&ctx->h_r[s], ctx->chunk * sizeof(RayRecord)), "cudaMallocHost");
        }
    }
    const uint64_t nchunks = (n + ctx->chunk - 1) / ctx->chunk;
    int launches = 0;
    // pageable callers: the copy-out of chunk k-K is drained just before slot reuse
    auto drain = [&](uint64_t k) -> cudaError_t {
        const int s = (int)(k % K);
        cudaError_t e = cudaStreamSynchronize(ctx->streams[s]);
        if (e != cudaSuccess) return e;
        if (!direct) {
            const uint64_t b = k * ctx->chunk, m = (n - b < ctx->chunk) ? n - b : ctx->chunk;
            std::memcpy(h_rays + b, ctx->h_r[s], m * sizeof(RayRecord));
        }
        return cudaSuccess;
    };
    for (uint64_t k = 0; k < nchunks; ++k) {
        const int s = (int)(k % K);
        if (k >= (uint64_t)K) ZCUDA(drain(k - K), "pipeline drain");
        const uint64_t b = k * ctx->chunk, m = (n - b < ctx->chunk) ? n - b : ctx->chunk;
        const float* src = h_samples + 4 * b;
        if (!direct) { std::memcpy(ctx->h_in[s], src, m * sizeof(float4)); src = (const float*)ctx->h_in[s]; }
        ZCUDA(cudaMemcpyAsync(ctx->d_in[s], src, m * sizeof(float4), cudaMemcpyHostToDevice, ctx->streams[s]), "H2D");
        Workspace ws = {nullptr, nullptr, 0};
        ZCUDA(get_workspace(ctx, ctx->streams[s], m, &ws), "workspace");
        cudaError_t e = launch_generate(ctx->host.state, ctx->mode, ctx->d_in[s], m, first_index + b, rng_seed, ctx->d_r[s],
                                        ctx->d_stats, ctx->streams[s], ws, &launches);
        if (e != cudaSuccess) { count_launches(launches); return cuda_fail(e, "zoicb_generate_host launch"); }
        void* dst = direct ? (void*)(h_rays + b) : (void*)ctx->h_r[s];
        ZCUDA(cudaMemcpyAsync(dst, ctx->d_r[s], m * sizeof(RayRecord), cudaMemcpyDeviceToHost, ctx->streams[s]), "D2H");
    }
    count_launches(launches);
    for (uint64_t k = (nchunks > (uint64_t)K ? nchunks - K : 0); k < nchunks; ++k) ZCUDA(drain(k), "pipeline drain");
    return ZOICB_OK;
}

namespace {
struct OneShot {  // per-thread resources of zoicb_generate_one
    int device = -1;
    cudaStream_t stream = nullptr;
    float4* h = nullptr;  // pinned: [0] sample, [2..3] the ray record (32-byte aligned)
    float4* d = nullptr;
    ~OneShot() {
        if (device >= 0) {
            cudaSetDevice(device);
            if (stream) cudaStreamDestroy(stream);
            if (h) cudaFreeHost(h);
            if (d) cudaFree(d);
        }
    }
};
thread_local OneShot t_one;
}  // namespace

zoicb_status zoicb_generate_one(zoicb_ctx* ctx, const float* sample, uint64_t sample_index, uint64_t rng_seed,
                                zoicb_ray* ray) {
    if (!ctx || !sample || !ray) return fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_generate_one: null argument");
    ZCUDA(cudaSetDevice(ctx->device), "cudaSetDevice");
    OneShot& r = t_one;
    if (r.device != ctx->device) {
        if (r.device >= 0) return fail(ZOICB_ERR_UNSUPPORTED, "zoicb_generate_one: one device per calling thread");
        ZCUDA(cudaStreamCreateWithFlags(&r.stream, cudaStreamNonBlocking), "cudaStreamCreate");
        ZCUDA(cudaMallocHost(&r.h, 4 * sizeof(float4)), "cudaMallocHost");
        ZCUDA(cudaMalloc(&r.d, 4 * sizeof(float4)), "cudaMalloc");
        r.device = ctx->device;
    }
    std::memcpy(&r.h[0], sample, sizeof(float4));
    ZCUDA(cudaMemcpyAsync(&r.d[0], &r.h[0], sizeof(float4), cudaMemcpyHostToDevice, r.stream), "H2D");
    int launches = 0;
    const Workspace ws = {nullptr, nullptr, 0};
    cudaError_t e = launch_generate(ctx->host.state, ZOICB_MODE_EXACT, &r.d[0], 1, sample_index, rng_seed,
                                    reinterpret_cast<RayRecord*>(&r.d[2]), ctx->d_stats, r.stream, ws, &launches);
    count_launches(launches);
    if (e != cudaSuccess) return cuda_fail(e, "zoicb_generate_one launch");
    ZCUDA(cudaMemcpyAsync(&r.h[2], &r.d[2], 2 * sizeof(float4), cudaMemcpyDeviceToHost, r.stream), "D2H");
    ZCUDA(cudaStreamSynchronize(r.stream), "cudaStreamSynchronize");
    std::memcpy(ray, &r.h[2], sizeof(zoicb_ray));
    return ZOICB_OK;
}

zoicb_status zoicb_write_draw_file(zoicb_ctx* ctx, const char* path, const float* h_samples, uint32_t n,
                                   const uint64_t* h_indices, uint64_t first_index, uint64_t rng_seed) {
    if (!ctx || !path) return fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_write_draw_file: null argument");
    if (n && !h_samples) return fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_write_draw_file: null samples");
    if (ctx->host.state.lens_model != ZOICB_RAYTRACED)
        return fail(ZOICB_ERR_UNSUPPORTED, "zoicb_write_draw_file: raytraced lens model only");
    if (n > 65536) return fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_write_draw_file: at most 65536 samples");
    const zoicb_constants& C = ctx->host.constants;
    const uint32_t cap = (uint32_t)(kMaxTries + 2) * (uint32_t)(C.lensCount + 1);
    std::vector<float4> quads((size_t)n * cap);
    std::vector<uint8_t> kinds((size_t)n * cap);
    std::vector<uint32_t> counts(n);
    if (n) {
        ZCUDA(cudaSetDevice(ctx->device), "cudaSetDevice");
        float4 *d_s = nullptr, *d_q = nullptr;
        uint8_t* d_k = nullptr;
        uint32_t* d_c = nullptr;
        unsigned long long* d_i = nullptr;
        cudaError_t e = cudaSuccess;
        int launches = 0;
        do {
            if ((e = cudaMalloc(&d_s, n * sizeof(float4))) != cudaSuccess) break;
            if ((e = cudaMalloc(&d_q, quads.size() * sizeof(float4))) != cudaSuccess) break;
            if ((e = cudaMalloc(&d_k, kinds.size())) != cudaSuccess) break;
            if ((e = cudaMalloc(&d_c, n * sizeof(uint32_t))) != cudaSuccess) break;
            if ((e = cudaMemcpy(d_s, h_samples, n * sizeof(float4), cudaMemcpyHostToDevice)) != cudaSuccess) break;
            if (h_indices) {
                if ((e = cudaMalloc(&d_i, n * sizeof(unsigned long long))) != cudaSuccess) break;
                if ((e = cudaMemcpy(d_i, h_indices, n * sizeof(unsigned long long), cudaMemcpyHostToDevice)) != cudaSuccess) break;
            }
            if ((e = launch_draw_paths(ctx->host.state, d_s, n, d_i, first_index, rng_seed, d_q, d_k, d_c, cap, nullptr, &launches)) != cudaSuccess) break;
            if ((e = cudaMemcpy(quads.data(), d_q, quads.size() * sizeof(float4), cudaMemcpyDeviceToHost)) != cudaSuccess) break;
            if ((e = cudaMemcpy(kinds.data(), d_k, kinds.size(), cudaMemcpyDeviceToHost)) != cudaSuccess) break;
            e = cudaMemcpy(counts.data(), d_c, n * sizeof(uint32_t), cudaMemcpyDeviceToHost);
        } while (0);
        cudaFree(d_s); cudaFree(d_q); cudaFree(d_k); cudaFree(d_c); cudaFree(d_i);
        count_launches(launches);
        if (e != cudaSuccess) return cuda_fail(e, "zoicb_write_draw_file");
    }
    FILE* f = std::fopen(path, "w");
    if (!f) return fail(ZOICB_ERR_INVALID_ARGUMENT, std::string("zoicb_write_draw_file: cannot open ") + path);
    // header: src/zoic.cpp:1618 and writeToFile :1240-1293 (std::fixed << std::setprecision(10) == "%.10f")
    std::fprintf(f, "LENSMODEL{KOLB}\nLENSES{");
    const double deg = (double)(180 / 3.14159265358979323846f);   // 180 / AI_PI is a float division
    float max_ap = 0.0f;
    for (int i = 0; i < C.lensCount; ++i) {
        const double ang = std::asin(((double)C.aperture[i] * 0.5) / (double)C.curvature[i]) * deg;
        std::fprintf(f, "%.10f %.10f %.10f ", (double)-C.center[i], (double)-C.curvature[i], ang);
        if (C.aperture[i] > max_ap) max_ap = C.aperture[i];
    }
    std::fprintf(f, "}\nIOR{");
    for (int i = 0; i < C.lensCount; ++i) std::fprintf(f, "%.10f ", (double)C.ior[i]);
    std::fprintf(f, "}\nAPERTUREELEMENT{%d}\n", C.apertureElement);
    std::fprintf(f, "APERTUREDISTANCE{%.10f}\n", (double)-C.apertureDistance);
    std::fprintf(f, "APERTURE{%.10f}\n", (double)C.userApertureRadius);
    std::fprintf(f, "APERTUREMAX{%.10f}\n", (double)max_ap);
    std::fprintf(f, "FOCUSDISTANCE{%.10f}\n", (double)-ctx->host.params.focalDistance);
    std::fprintf(f, "IMAGEDISTANCE{%.10f}\n", (double)-C.originShift);
    std::fprintf(f, "SENSORHEIGHT{%.10f}\nRAYS{", 1.7);
    for (uint32_t i = 0; i < n; ++i) {
        for (uint32_t k = 0; k < counts[i] && k < cap; ++k) {
            const float4 q = quads[(size_t)i * cap + k];
            if (kinds[(size_t)i * cap + k] == 0)   // :1121-1128
                std::fprintf(f, "%.10f %.10f %.10f %.10f ", (double)-q.x, (double)-q.y, (double)-q.z, (double)-q.w);
            else                                   // :1146-1153: float + float * -10000.0 evaluated in double
                std::fprintf(f, "%.10f %.10f %.10f %.10f ", (double)-q.x, (double)-q.y, (double)q.x + (double)q.z * -10000.0,
                             (double)q.y + (double)q.w * -10000.0);
        }
    }
    std::fprintf(f, "}");
    std::fclose(f);
    return ZOICB_OK;
}

zoicb_status zoicb_transform_rays(zoicb_ctx* ctx, const zoicb_ray* d_rays, uint64_t n, const float* m3x4, zoicb_ray* d_out,
                                  void* stream) {
    if (!ctx) return fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_transform_rays: null context");
    if (n == 0) return ZOICB_OK;
    if (!d_rays || !d_out || !m3x4) return fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_transform_rays: null argument");
    if (((uintptr_t)d_rays & 31u) || ((uintptr_t)d_out & 31u))
        return fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_transform_rays: rays must be 32-byte aligned");
    ZCUDA(cudaSetDevice(ctx->device), "cudaSetDevice");
    int launches = 0;
    cudaError_t e = launch_transform(m3x4, (const RayRecord*)d_rays, n, (RayRecord*)d_out, (cudaStream_t)stream, &launches);
    count_launches(launches);
    if (e != cudaSuccess) return cuda_fail(e, "zoicb_transform_rays launch");
    return ZOICB_OK;
}

zoicb_status zoicb_synth_samples(zoicb_ctx* ctx, uint32_t W, uint32_t H, uint32_t spp, uint64_t seed, uint64_t first_index,
                                 uint64_t n, void* d_samples, void* stream) {
    if (!ctx) return fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_synth_samples: null context");
    if (!W || !H || !spp) return fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_synth_samples: zero dimension");
    if (n == 0) return ZOICB_OK;
    if (!d_samples) return fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_synth_samples: null buffer");
    ZCUDA(cudaSetDevice(ctx->device), "cudaSetDevice");
    int launches = 0;
    cudaError_t e = launch_synth(W, H, spp, seed, first_index, n, (float4*)d_samples, (cudaStream_t)stream, &launches);
    count_launches(launches);
    if (e != cudaSuccess) return cuda_fail(e, "zoicb_synth_samples launch");
    return ZOICB_OK;
}

zoicb_status zoicb_get_stats(zoicb_ctx* ctx, zoicb_stats* out) {
    if (!ctx || !out) return fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_get_stats: null argument");
    ZCUDA(cudaSetDevice(ctx->device), "cudaSetDevice");
    ZCUDA(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
    DeviceStats h;
    ZCUDA(cudaMemcpy(&h, ctx->d_stats, sizeof h, cudaMemcpyDeviceToHost), "cudaMemcpy(stats)");
    out->rays = h.rays; out->success = h.success; out->vignetted = h.vignetted;
    out->total_internal_reflection = h.tir; out->attempts = h.attempts; out->element_visits = h.element_visits;
    out->exact_reruns = h.exact_reruns;
    return ZOICB_OK;
}

zoicb_status zoicb_reset_stats(zoicb_ctx* ctx) {
    if (!ctx) return fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_reset_stats: null context");
    ZCUDA(cudaSetDevice(ctx->device), "cudaSetDevice");
    ZCUDA(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
    ZCUDA(cudaMemset(ctx->d_stats, 0, sizeof(DeviceStats)), "cudaMemset(stats)");
    return ZOICB_OK;
}

zoicb_status zoicb_get_constants(const zoicb_ctx* ctx, zoicb_constants* out) {
    if (!ctx || !out) return fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_get_constants: null argument");
    *out = ctx->host.constants;
    return ZOICB_OK;
}

zoicb_status zoicb_get_bokeh_tables(const zoicb_ctx* ctx, float* cdfRow, int32_t* rowIndices, float* cdfColumn,
                                    int32_t* columnIndices) {
    if (!ctx) return fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_get_bokeh_tables: null context");
    const HostBokeh& hb = ctx->host.bokeh;
    if (!hb.valid()) return fail(ZOICB_ERR_BOKEH_IMAGE, "zoicb_get_bokeh_tables: camera has no bokeh image");
    const size_t np = (size_t)hb.w * hb.h;
    if (cdfRow) std::memcpy(cdfRow, hb.cdf_row.data(), hb.h * sizeof(float));
    if (rowIndices) std::memcpy(rowIndices, hb.row_indices.data(), hb.h * sizeof(int32_t));
    if (cdfColumn) std::memcpy(cdfColumn, hb.cdf_column.data(), np * sizeof(float));
    if (columnIndices) std::memcpy(columnIndices, hb.column_indices.data(), np * sizeof(int32_t));
    return ZOICB_OK;
}

zoicb_status zoicb_setup_host_only(const zoicb_params* params, const float* rgb, int width, int height, int nch,
                                   zoicb_constants* out, float* cdfRow, int32_t* rowIndices, float* cdfColumn,
                                   int32_t* columnIndices) {
    if (!params || !out) return fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_setup_host_only: null argument");
    HostCamera hc;
    std::string err;
    zoicb_status rc = build_camera(*params, rgb, width, height, nch, &hc, &err, nullptr, nullptr);
    if (rc != ZOICB_OK) return fail(rc, "zoicb_setup_host_only: " + err);
    *out = hc.constants;
    const HostBokeh& hb = hc.bokeh;
    if (hb.valid()) {
        const size_t np = (size_t)hb.w * hb.h;
        if (cdfRow) std::memcpy(cdfRow, hb.cdf_row.data(), hb.h * sizeof(float));
        if (rowIndices) std::memcpy(rowIndices, hb.row_indices.data(), hb.h * sizeof(int32_t));
        if (cdfColumn) std::memcpy(cdfColumn, hb.cdf_column.data(), np * sizeof(float));
        if (columnIndices) std::memcpy(columnIndices, hb.column_indices.data(), np * sizeof(int32_t));
    }
    return ZOICB_OK;
}

zoicb_status zoicb_build_bokeh_tables(int device, const float* rgb, int width, int height, int nch, float* cdfRow,
                                      int32_t* rowIndices, float* cdfColumn, int32_t* columnIndices, float* ms) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(ZOICB_ERR_CUDA, "zoicb_build_bokeh_tables: no CUDA device (libzoicb has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_build_bokeh_tables: bad device index");
    ZCUDA(cudaSetDevice(device), "cudaSetDevice");
    std::string err;
    zoicb_status rc = check_bokeh_image(rgb, width, height, nch, &err);
    if (rc != ZOICB_OK) return fail(rc, "zoicb_build_bokeh_tables: " + err);
    zoicb_ctx* c = new zoicb_ctx();
    c->device = device;
    HostBokeh hb;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    c->bokeh_ev0 = e0; c->bokeh_ev1 = e1;
    const bool ok = bokeh_build_gpu(c, rgb, width, height, nch, &hb);
    if (ok && ms) cudaEventElapsedTime(ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    free_ctx(c);
    if (!ok) return fail(ZOICB_ERR_CUDA, "zoicb_build_bokeh_tables: device build failed");
    const size_t np = (size_t)width * height;
    if (cdfRow) std::memcpy(cdfRow, hb.cdf_row.data(), height * sizeof(float));
    if (rowIndices) std::memcpy(rowIndices, hb.row_indices.data(), height * sizeof(int32_t));
    if (cdfColumn) std::memcpy(cdfColumn, hb.cdf_column.data(), np * sizeof(float));
    if (columnIndices) std::memcpy(columnIndices, hb.column_indices.data(), np * sizeof(int32_t));
    return ZOICB_OK;
}

zoicb_status zoicb_debug_sort_orders(const float* values, int32_t n, int32_t* restated, int32_t* library) {
    if (n < 0 || (n > 0 && !values)) return fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_debug_sort_orders: bad argument");
    if (restated) {
        for (int32_t i = 0; i < n; ++i) restated[i] = i;
        gnusort::sort(restated, (long)n, gnusort::Before<int32_t>{values});
    }
    if (library) std_sort_desc(values, n, library);
    return ZOICB_OK;
}

zoicb_status zoicb_measure_fp32_peak(int device, double* tflops) {
    if (!tflops) return fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_measure_fp32_peak: null argument");
    ZCUDA(cudaSetDevice(device), "cudaSetDevice");
    int launches = 0;
    cudaError_t e = measure_fp32_peak(tflops, &launches);
    count_launches(launches);
    if (e != cudaSuccess) return cuda_fail(e, "zoicb_measure_fp32_peak");
    return ZOICB_OK;
}

uint64_t zoicb_kernel_launches(void) { return g_launches.load(std::memory_order_relaxed); }
const char* zoicb_last_error(void) { return g_last_error.c_str(); }
const char* zoicb_version(void) { return "zoicb 0.1 (sm_100a)"; }

}  // extern "C"
