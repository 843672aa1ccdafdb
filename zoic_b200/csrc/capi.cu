// capi.cu -- the C ABI declared in include/zoicb.h.
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <string>
#include <vector>

#include "capi_internal.h"
#include "gnu_sort.h"
#include "lens_math.cuh"

using namespace zoicb;

static_assert(sizeof(zoicb_ray) == 32 && sizeof(RayRecord) == 32, "a ray is one 32-byte record");
static_assert(sizeof(zoicb_ray_diff) == 48, "differentials are 12 floats");

namespace {
thread_local std::string g_last_error;
std::atomic<uint64_t> g_launches{0};
double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
}  // namespace

namespace zoicb {
zoicb_status api_fail(zoicb_status code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
zoicb_status api_cuda_fail(cudaError_t e, const char* what) {
    return api_fail(ZOICB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
void api_count_launches(int k) { if (k > 0) g_launches.fetch_add((uint64_t)k, std::memory_order_relaxed); }
}  // namespace zoicb

namespace {

// exit-pupil LUT on the GPU: the 3.2 M candidates are classified by lut_trace_kernel and folded into
// the 32 bounding boxes by lut_bbox_kernel (in-order fold, re-arm quirk included); only 32 x 4 floats come back.
int lut_trace_gpu(void* user, const LensState& lens, const float* film_x, int n_film, const uint32_t* draws,
                  int per_film, uint8_t* accept, float* boxes) {
    zoicb_ctx* c = static_cast<zoicb_ctx*>(user);
    (void)accept;
    const double t0 = now_ms();
    const size_t total = (size_t)n_film * per_film;
    float* d_film = nullptr;
    uint32_t* d_draws = nullptr;
    uint8_t* d_acc = nullptr;
    float4* d_boxes = nullptr;
    bool ok = false;
    int launches = 0;
    do {
        if (cudaMalloc(&d_film, n_film * sizeof(float)) != cudaSuccess) break;
        if (cudaMalloc(&d_draws, total * 2 * sizeof(uint32_t)) != cudaSuccess) break;
        if (cudaMalloc(&d_acc, total) != cudaSuccess) break;
        if (cudaMalloc(&d_boxes, n_film * sizeof(float4)) != cudaSuccess) break;
        if (cudaMemcpy(d_film, film_x, n_film * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) break;
        if (cudaMemcpy(d_draws, draws, total * 2 * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess) break;
        if (launch_lut_trace(lens, d_film, n_film, per_film, d_draws, d_acc, nullptr, &launches) != cudaSuccess) break;
        if (launch_lut_bbox(d_draws, d_acc, n_film, per_film, lens.first_aperture, d_boxes, nullptr, &launches) != cudaSuccess) break;
        if (cudaMemcpy(boxes, d_boxes, n_film * sizeof(float4), cudaMemcpyDeviceToHost) != cudaSuccess) break;
        ok = true;
    } while (0);
    api_count_launches(launches);
    cudaFree(d_film);
    cudaFree(d_draws);
    cudaFree(d_acc);
    cudaFree(d_boxes);
    if (!ok) cudaGetLastError();
    if (c) c->lut_ms = now_ms() - t0;
    return ok ? 2 : 0;
}

// Image-based aperture tables built on the GPU (bokeh_build.cu).  The tables stay on the device in
// the context; a copy comes back for zoicb_get_bokeh_tables.
bool bokeh_build_gpu(void* user, const float* rgb, int w, int h, int nch, HostBokeh* out) {
    zoicb_ctx* c = static_cast<zoicb_ctx*>(user);
    const size_t np = (size_t)w * h;
    const int row_shift = out->row_shift, col_shift = out->col_shift;   // set by build_camera
    const size_t ng_row = ((size_t)1 << row_shift) + 2, ng_col = (size_t)h * (((size_t)1 << col_shift) + 2);
    const double t0 = now_ms();
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
                               c->d_cdf_col, c->d_rel_col, row_shift, col_shift, c->d_row_guide, c->d_col_guide, nullptr, &launches) != cudaSuccess) break;
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
    api_count_launches(launches);
    cudaFree(d_rgb); cudaFree(d_work); cudaFree(d_scratch); cudaFree(d_total); cudaFree(d_row_mass);
    if (!ok) { cudaGetLastError(); out->w = out->h = 0; }
    c->bokeh_ms = now_ms() - t0;
    return ok;
}

}  // namespace

// Scratch for the guarded mode on `st`, large enough for n samples: room for n/24 undecided samples
// (measured rates are 1e-4 .. 1.7e-2); anything beyond the capacity is settled inline by the kernel.
// The caller holds ctx->gen_mu until its launches are enqueued; a larger queue replaces the old one only after the
// stream has drained, so no kernel in flight and no other caller still uses the freed buffer.
cudaError_t zoicb::api_get_workspace(zoicb_ctx* c, cudaStream_t st, uint64_t n, Workspace* out) {
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
        if ((e = cudaMalloc(&w.queue, want * sizeof(QueueRecord))) != cudaSuccess) return e;
        w.capacity = want;
    }
    *out = w;
    return cudaSuccess;
}

namespace {

void free_ctx(zoicb_ctx* c) {
    if (!c) return;
    DeviceGuard guard(c->device);
    if (c->job_cache && c->job_cache_free) c->job_cache_free(c->job_cache);
    for (auto& kv : c->workspaces) { cudaFree(kv.second.counters); cudaFree(kv.second.queue); }
    cudaFree(c->d_cdf_row); cudaFree(c->d_row_idx); cudaFree(c->d_cdf_col); cudaFree(c->d_rel_col);
    cudaFree(c->d_row_guide); cudaFree(c->d_col_guide); cudaFree(c->d_dxy); cudaFree(c->d_compact);
    cudaFree(c->d_stats);
    for (int s = 0; s < zoicb_ctx::kSlots; ++s) {
        if (c->streams[s]) cudaStreamDestroy(c->streams[s]);
        cudaFree(c->d_in[s]); cudaFree(c->d_r[s]); cudaFree(c->d_p[s]);
        if (c->h_in[s]) cudaFreeHost(c->h_in[s]);
        if (c->h_r[s]) cudaFreeHost(c->h_r[s]);
        if (c->h_p[s]) cudaFreeHost(c->h_p[s]);
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
    if (!params || !out) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_create: null argument");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return api_fail(ZOICB_ERR_CUDA, "zoicb_create: no CUDA device (libzoicb has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_create: bad device index");
    ZGUARD(device);
    cudaDeviceProp prop;
    ZCUDA(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties");
    if (prop.major != 10) return api_fail(ZOICB_ERR_CUDA, "zoicb_create: kernels are built for sm_100a only; device is " + std::string(prop.name));

    const double t_create = now_ms();
    zoicb_ctx* c = new zoicb_ctx();
    c->device = device;
    std::string err;
    zoicb_status rc = build_camera(*params, rgb, width, height, nch, &c->host, &err, lut_trace_gpu, c, bokeh_build_gpu, c);
    if (rc != ZOICB_OK) { free_ctx(c); return api_fail(rc, "zoicb_create: " + err); }

    auto bail = [&](cudaError_t ce, const char* what) { free_ctx(c); return api_cuda_fail(ce, what); };
    if ((e = cudaMalloc(&c->d_stats, sizeof(DeviceStats))) != cudaSuccess) return bail(e, "cudaMalloc(stats)");
    if ((e = cudaMemset(c->d_stats, 0, sizeof(DeviceStats))) != cudaSuccess) return bail(e, "cudaMemset(stats)");
    const HostBokeh& hb = c->host.bokeh;
    BokehTables& bt = c->host.state.bokeh;
    if (hb.valid()) {   // tables were built on the device by bokeh_build_gpu and stayed there
        bt.cdf_row = c->d_cdf_row; bt.row_indices = c->d_row_idx; bt.cdf_column = c->d_cdf_col; bt.rel_column = c->d_rel_col;
        bt.row_guide = c->d_row_guide; bt.col_guide = c->d_col_guide;
        bt.dx_of_col = c->d_dxy; bt.dy_of_row = c->d_dxy + hb.w;
        bt.w = hb.w; bt.h = hb.h; bt.row_shift = hb.row_shift; bt.col_shift = hb.col_shift;
        if (params->lensModel == ZOICB_THINLENS && hb.w <= kCompactMaxWidth && hb.h <= kCompactMaxRows) {
            // byte-wide copies of the two big column tables for the thin-lens retry kernel (camera_state.h: BokehCompact)
            const size_t ng = hb.col_guide.size(), np = (size_t)hb.w * hb.h;
            std::vector<uint8_t> narrow(ng + np);
            for (size_t i = 0; i < ng; ++i) narrow[i] = (uint8_t)hb.col_guide[i];
            for (int r = 0; r < hb.h; ++r)
                for (int k = 0; k < hb.w; ++k) narrow[ng + (size_t)r * hb.w + k] = (uint8_t)(hb.column_indices[(size_t)r * hb.w + k] - r * hb.w);
            if ((e = cudaMalloc(&c->d_compact, narrow.size())) != cudaSuccess) return bail(e, "cudaMalloc(bokeh)");
            if ((e = cudaMemcpy(c->d_compact, narrow.data(), narrow.size(), cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e, "cudaMemcpy(bokeh)");
            c->host.state.compact.col_guide8 = c->d_compact;
            c->host.state.compact.rel_column8 = c->d_compact + ng;
        }
    } else if (hb.degenerate) {
        // An image the reference accepts but treats as invalid (fewer than 3 channels, src/zoic.cpp:135-137): every
        // bokehSample answers the lens centre (0, 0) (:420-425).  The kernels keep their one code path: a 1 x 1 table whose
        // CDF entry is +inf (every u, NaN included after the clamp, lands on entry 0) and whose lens coordinates are +0.
        const float inf = std::numeric_limits<float>::infinity();
        const float cdf1[1] = {inf}, zero2[2] = {0.0f, 0.0f};
        const int32_t idx1[1] = {0};
        const uint16_t rel1[1] = {0}, guide3[3] = {0, 0, 0};   // shift 0: G = 1 cell, G + 2 = 3 entries
        if ((e = cudaMalloc(&c->d_cdf_row, sizeof cdf1)) != cudaSuccess) return bail(e, "cudaMalloc(bokeh)");
        if ((e = cudaMalloc(&c->d_row_idx, sizeof idx1)) != cudaSuccess) return bail(e, "cudaMalloc(bokeh)");
        if ((e = cudaMalloc(&c->d_cdf_col, sizeof cdf1)) != cudaSuccess) return bail(e, "cudaMalloc(bokeh)");
        if ((e = cudaMalloc(&c->d_rel_col, 4)) != cudaSuccess) return bail(e, "cudaMalloc(bokeh)");
        if ((e = cudaMalloc(&c->d_row_guide, 8)) != cudaSuccess) return bail(e, "cudaMalloc(bokeh)");
        if ((e = cudaMalloc(&c->d_col_guide, 8)) != cudaSuccess) return bail(e, "cudaMalloc(bokeh)");
        if ((e = cudaMalloc(&c->d_dxy, sizeof zero2)) != cudaSuccess) return bail(e, "cudaMalloc(bokeh)");
        cudaMemcpy(c->d_cdf_row, cdf1, sizeof cdf1, cudaMemcpyHostToDevice);
        cudaMemcpy(c->d_row_idx, idx1, sizeof idx1, cudaMemcpyHostToDevice);
        cudaMemcpy(c->d_cdf_col, cdf1, sizeof cdf1, cudaMemcpyHostToDevice);
        cudaMemcpy(c->d_rel_col, rel1, sizeof rel1, cudaMemcpyHostToDevice);
        cudaMemcpy(c->d_row_guide, guide3, sizeof guide3, cudaMemcpyHostToDevice);
        cudaMemcpy(c->d_col_guide, guide3, sizeof guide3, cudaMemcpyHostToDevice);
        if ((e = cudaMemcpy(c->d_dxy, zero2, sizeof zero2, cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e, "cudaMemcpy(bokeh)");
        bt.cdf_row = c->d_cdf_row; bt.row_indices = c->d_row_idx; bt.cdf_column = c->d_cdf_col; bt.rel_column = c->d_rel_col;
        bt.row_guide = c->d_row_guide; bt.col_guide = c->d_col_guide;
        bt.dx_of_col = c->d_dxy; bt.dy_of_row = c->d_dxy + 1;
        bt.w = 1; bt.h = 1; bt.row_shift = 0; bt.col_shift = 0;
    }
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) return bail(e, "zoicb_create");
    c->create_ms = now_ms() - t_create;
    *out = c;
    return ZOICB_OK;
}

void zoicb_destroy(zoicb_ctx* ctx) { free_ctx(ctx); }

zoicb_status zoicb_set_mode(zoicb_ctx* ctx, int mode) {
    if (!ctx) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_set_mode: null context");
    if (mode != ZOICB_MODE_EXACT && mode != ZOICB_MODE_GUARDED) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_set_mode: unknown mode");
    ctx->mode = mode;
    return ZOICB_OK;
}
int zoicb_get_mode(const zoicb_ctx* ctx) { return ctx ? ctx->mode : -1; }

zoicb_status zoicb_set_guard_scale(zoicb_ctx* ctx, float scale) {
    if (!ctx || !(scale >= 0.0f)) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_set_guard_scale: bad argument");
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
    if (!ctx) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_generate: null context");
    if (n == 0) return ZOICB_OK;
    if (!d_samples || !d_rays) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_generate: null buffer");
    if (((uintptr_t)d_rays & 31u) || ((uintptr_t)d_samples & 15u))
        return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_generate: samples must be 16-byte and rays 32-byte aligned");
    if (n > (1ull << 40)) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_generate: at most 2^40 samples per call (split the batch; first_index keeps the streams)");
    ZGUARD(ctx->device);
    int launches = 0;
    Workspace ws = {nullptr, nullptr, 0};
    std::lock_guard<std::mutex> lock(ctx->gen_mu);   // workspace lookup + the whole enqueue sequence (see capi_internal.h)
    ZCUDA(api_get_workspace(ctx, (cudaStream_t)stream, n, &ws), "workspace");
    cudaError_t e = launch_generate(ctx->host.state, ctx->mode, (const float4*)d_samples, n, first_index, rng_seed,
                                    (RayRecord*)d_rays, ctx->d_stats, (cudaStream_t)stream, ws, &launches);
    api_count_launches(launches);
    if (e != cudaSuccess) return api_cuda_fail(e, "zoicb_generate launch");
    return ZOICB_OK;
}

}  // extern "C"

namespace {

// zoicb_generate_host / zoicb_generate_host_planar: host samples in, host rays out, through a 3-slot pipeline of
// upload / kernels / (pack) / download.  Exactly one of h_rays (32-byte records) and planes (25 bytes per ray) is set.
zoicb_status generate_host_impl(zoicb_ctx* ctx, const float* h_samples, uint64_t n, uint64_t first_index, uint64_t rng_seed,
                                zoicb_ray* h_rays, const zoicb_ray_planes* planes) {
    std::lock_guard<std::mutex> lock(ctx->host_mu);
    ZGUARD(ctx->device);
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
    const uint64_t C = ctx->chunk;
    // planar: device chunk = six float planes of C entries, then C bytes
    if (planes && !ctx->d_p[0])
        for (int s = 0; s < K; ++s) ZCUDA(cudaMalloc(&ctx->d_p[s], C * 25), "cudaMalloc(staging)");
    void* const plane_ptr[7] = {planes ? planes->origin[0] : nullptr, planes ? planes->origin[1] : nullptr,
                                planes ? planes->origin[2] : nullptr, planes ? planes->dir[0] : nullptr,
                                planes ? planes->dir[1] : nullptr, planes ? planes->dir[2] : nullptr,
                                planes ? (void*)planes->flags : nullptr};
    bool direct = is_pinned_host(h_samples);
    if (planes) { for (void* q : plane_ptr) direct = direct && is_pinned_host(q); }
    else direct = direct && is_pinned_host(h_rays);
    if (!direct) {
        for (int s = 0; s < K; ++s) {
            if (!ctx->h_in[s]) ZCUDA(cudaMallocHost(&ctx->h_in[s], C * sizeof(float4)), "cudaMallocHost");
            if (!planes && !ctx->h_r[s]) ZCUDA(cudaMallocHost(&ctx->h_r[s], C * sizeof(RayRecord)), "cudaMallocHost");
            if (planes && !ctx->h_p[s]) ZCUDA(cudaMallocHost(&ctx->h_p[s], C * 25), "cudaMallocHost");
        }
    }
    const uint64_t nchunks = (n + C - 1) / C;
    int launches = 0;
    auto plane_bytes = [](int j) -> uint64_t { return j < 6 ? 4u : 1u; };
    // pageable callers: the copy-out of chunk k-K is drained just before slot reuse
    auto drain = [&](uint64_t k) -> cudaError_t {
        const int s = (int)(k % K);
        cudaError_t e = cudaStreamSynchronize(ctx->streams[s]);
        if (e != cudaSuccess) return e;
        if (!direct) {
            const uint64_t b = k * C, m = (n - b < C) ? n - b : C;
            if (planes) {
                for (int j = 0; j < 7; ++j)
                    std::memcpy(static_cast<char*>(plane_ptr[j]) + b * plane_bytes(j), ctx->h_p[s] + (uint64_t)j * 4u * C, m * plane_bytes(j));
            } else {
                std::memcpy(h_rays + b, ctx->h_r[s], m * sizeof(RayRecord));
            }
        }
        return cudaSuccess;
    };
    // On any failure every pipeline stream is drained before returning, so that no copy still in flight writes into
    // the caller's buffers after the error return.
    const char* what = "";
    auto run = [&]() -> cudaError_t {
        cudaError_t e;
        for (uint64_t k = 0; k < nchunks; ++k) {
            const int s = (int)(k % K);
            what = "pipeline drain";
            if (k >= (uint64_t)K && (e = drain(k - K)) != cudaSuccess) return e;
            const uint64_t b = k * C, m = (n - b < C) ? n - b : C;
            const float* src = h_samples + 4 * b;
            if (!direct) { std::memcpy(ctx->h_in[s], src, m * sizeof(float4)); src = (const float*)ctx->h_in[s]; }
            what = "H2D";
            if ((e = cudaMemcpyAsync(ctx->d_in[s], src, m * sizeof(float4), cudaMemcpyHostToDevice, ctx->streams[s])) != cudaSuccess) return e;
            {
                std::lock_guard<std::mutex> gen_lock(ctx->gen_mu);
                Workspace ws = {nullptr, nullptr, 0};
                what = "workspace";
                if ((e = api_get_workspace(ctx, ctx->streams[s], m, &ws)) != cudaSuccess) return e;
                what = "zoicb_generate_host launch";
                if ((e = launch_generate(ctx->host.state, ctx->mode, ctx->d_in[s], m, first_index + b, rng_seed, ctx->d_r[s],
                                         ctx->d_stats, ctx->streams[s], ws, &launches)) != cudaSuccess) return e;
            }
            what = "D2H";
            if (planes) {
                if ((e = launch_pack_planar(ctx->d_r[s], m, C, ctx->d_p[s], ctx->streams[s], &launches)) != cudaSuccess) return e;
                for (int j = 0; j < 7; ++j) {   // plane j of the chunk starts at byte 4 j C (the byte plane last)
                    void* dst = direct ? (void*)(static_cast<char*>(plane_ptr[j]) + b * plane_bytes(j)) : (void*)(ctx->h_p[s] + (uint64_t)j * 4u * C);
                    if ((e = cudaMemcpyAsync(dst, ctx->d_p[s] + (uint64_t)j * 4u * C, m * plane_bytes(j), cudaMemcpyDeviceToHost, ctx->streams[s])) != cudaSuccess) return e;
                }
            } else {
                void* dst = direct ? (void*)(h_rays + b) : (void*)ctx->h_r[s];
                if ((e = cudaMemcpyAsync(dst, ctx->d_r[s], m * sizeof(RayRecord), cudaMemcpyDeviceToHost, ctx->streams[s])) != cudaSuccess) return e;
            }
        }
        what = "pipeline drain";
        for (uint64_t k = (nchunks > (uint64_t)K ? nchunks - K : 0); k < nchunks; ++k)
            if ((e = drain(k)) != cudaSuccess) return e;
        return cudaSuccess;
    };
    const cudaError_t e = run();
    api_count_launches(launches);
    if (e != cudaSuccess) {
        for (int s = 0; s < K; ++s) cudaStreamSynchronize(ctx->streams[s]);
        cudaGetLastError();
        return api_cuda_fail(e, what);
    }
    return ZOICB_OK;
}

}  // namespace

extern "C" {

zoicb_status zoicb_generate_host(zoicb_ctx* ctx, const float* h_samples, uint64_t n, uint64_t first_index,
                                 uint64_t rng_seed, zoicb_ray* h_rays) {
    if (!ctx) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_generate_host: null context");
    if (n == 0) return ZOICB_OK;
    if (!h_samples || !h_rays) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_generate_host: null buffer");
    return generate_host_impl(ctx, h_samples, n, first_index, rng_seed, h_rays, nullptr);
}

zoicb_status zoicb_generate_host_planar(zoicb_ctx* ctx, const float* h_samples, uint64_t n, uint64_t first_index,
                                        uint64_t rng_seed, const zoicb_ray_planes* out, float* live_weight) {
    if (!ctx) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_generate_host_planar: null context");
    if (live_weight) *live_weight = ctx->host.state.weight_scale;   // weight = {1, 0} * the exposure scale, in every kernel
    if (n == 0) return ZOICB_OK;
    if (!h_samples || !out) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_generate_host_planar: null buffer");
    for (int k = 0; k < 3; ++k)
        if (!out->origin[k] || !out->dir[k]) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_generate_host_planar: null plane");
    if (!out->flags) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_generate_host_planar: null plane");
    return generate_host_impl(ctx, h_samples, n, first_index, rng_seed, nullptr, out);
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
    if (!ctx || !sample || !ray) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_generate_one: null argument");
    ZGUARD(ctx->device);
    OneShot& r = t_one;
    if (r.device != ctx->device) {
        if (r.device >= 0) return api_fail(ZOICB_ERR_UNSUPPORTED, "zoicb_generate_one: one device per calling thread");
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
    api_count_launches(launches);
    if (e != cudaSuccess) return api_cuda_fail(e, "zoicb_generate_one launch");
    ZCUDA(cudaMemcpyAsync(&r.h[2], &r.d[2], 2 * sizeof(float4), cudaMemcpyDeviceToHost, r.stream), "D2H");
    ZCUDA(cudaStreamSynchronize(r.stream), "cudaStreamSynchronize");
    std::memcpy(ray, &r.h[2], sizeof(zoicb_ray));
    return ZOICB_OK;
}

zoicb_status zoicb_write_draw_file(zoicb_ctx* ctx, const char* path, const float* h_samples, uint32_t n,
                                   const uint64_t* h_indices, uint64_t first_index, uint64_t rng_seed) {
    if (!ctx || !path) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_write_draw_file: null argument");
    if (n && !h_samples) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_write_draw_file: null samples");
    if (ctx->host.state.lens_model != ZOICB_RAYTRACED)
        return api_fail(ZOICB_ERR_UNSUPPORTED, "zoicb_write_draw_file: raytraced lens model only");
    if (n > 65536) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_write_draw_file: at most 65536 samples");
    const zoicb_constants& C = ctx->host.constants;
    const uint32_t cap = (uint32_t)(kMaxTries + 2) * (uint32_t)(C.lensCount + 1);
    std::vector<float4> quads((size_t)n * cap);
    std::vector<uint8_t> kinds((size_t)n * cap);
    std::vector<uint32_t> counts(n);
    if (n) {
        ZGUARD(ctx->device);
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
        api_count_launches(launches);
        if (e != cudaSuccess) return api_cuda_fail(e, "zoicb_write_draw_file");
    }
    FILE* f = std::fopen(path, "w");
    if (!f) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, std::string("zoicb_write_draw_file: cannot open ") + path);
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
    if (!ctx) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_transform_rays: null context");
    if (n == 0) return ZOICB_OK;
    if (!d_rays || !d_out || !m3x4) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_transform_rays: null argument");
    if (((uintptr_t)d_rays & 31u) || ((uintptr_t)d_out & 31u))
        return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_transform_rays: rays must be 32-byte aligned");
    ZGUARD(ctx->device);
    int launches = 0;
    cudaError_t e = launch_transform(m3x4, (const RayRecord*)d_rays, n, (RayRecord*)d_out, (cudaStream_t)stream, &launches);
    api_count_launches(launches);
    if (e != cudaSuccess) return api_cuda_fail(e, "zoicb_transform_rays launch");
    return ZOICB_OK;
}

zoicb_status zoicb_differentials(zoicb_ctx* ctx, const void* d_samples, uint64_t n, uint64_t first_index, uint64_t rng_seed,
                                 float dsx, float dsy, const zoicb_ray* d_rays, zoicb_ray_diff* d_out, void* stream) {
    if (!ctx) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_differentials: null context");
    if (n == 0) return ZOICB_OK;
    if (!d_samples || !d_rays || !d_out) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_differentials: null buffer");
    if (((uintptr_t)d_rays & 31u) || ((uintptr_t)d_samples & 15u) || ((uintptr_t)d_out & 15u))
        return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_differentials: samples / differentials must be 16-byte and rays 32-byte aligned");
    ZGUARD(ctx->device);
    int launches = 0;
    cudaError_t e = launch_differentials(ctx->host.state, (const float4*)d_samples, n, first_index, rng_seed, dsx, dsy,
                                         (const RayRecord*)d_rays, (float4*)d_out, (cudaStream_t)stream, &launches);
    api_count_launches(launches);
    if (e != cudaSuccess) return api_cuda_fail(e, "zoicb_differentials launch");
    return ZOICB_OK;
}

zoicb_status zoicb_transform_differentials(zoicb_ctx* ctx, const zoicb_ray_diff* d_diffs, uint64_t n, const float* m3x4,
                                           zoicb_ray_diff* d_out, void* stream) {
    if (!ctx) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_transform_differentials: null context");
    if (n == 0) return ZOICB_OK;
    if (!d_diffs || !d_out || !m3x4) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_transform_differentials: null argument");
    if (((uintptr_t)d_diffs & 15u) || ((uintptr_t)d_out & 15u))
        return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_transform_differentials: differentials must be 16-byte aligned");
    ZGUARD(ctx->device);
    int launches = 0;
    cudaError_t e = launch_transform_diffs(m3x4, (const float4*)d_diffs, n, (float4*)d_out, (cudaStream_t)stream, &launches);
    api_count_launches(launches);
    if (e != cudaSuccess) return api_cuda_fail(e, "zoicb_transform_differentials launch");
    return ZOICB_OK;
}

zoicb_status zoicb_synth_samples(zoicb_ctx* ctx, uint32_t W, uint32_t H, uint32_t spp, uint64_t seed, uint64_t first_index,
                                 uint64_t n, void* d_samples, void* stream) {
    if (!ctx) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_synth_samples: null context");
    if (!W || !H || !spp) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_synth_samples: zero dimension");
    if (n == 0) return ZOICB_OK;
    if (!d_samples) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_synth_samples: null buffer");
    ZGUARD(ctx->device);
    int launches = 0;
    cudaError_t e = launch_synth(W, H, spp, seed, first_index, n, (float4*)d_samples, (cudaStream_t)stream, &launches);
    api_count_launches(launches);
    if (e != cudaSuccess) return api_cuda_fail(e, "zoicb_synth_samples launch");
    return ZOICB_OK;
}

zoicb_status zoicb_get_stats(zoicb_ctx* ctx, zoicb_stats* out) {
    if (!ctx || !out) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_get_stats: null argument");
    ZGUARD(ctx->device);
    ZCUDA(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
    DeviceStats h;
    ZCUDA(cudaMemcpy(&h, ctx->d_stats, sizeof h, cudaMemcpyDeviceToHost), "cudaMemcpy(stats)");
    out->rays = h.rays; out->success = h.success; out->vignetted = h.vignetted;
    out->total_internal_reflection = h.tir; out->attempts = h.attempts; out->element_visits = h.element_visits;
    out->exact_reruns = h.exact_reruns;
    return ZOICB_OK;
}

zoicb_status zoicb_reset_stats(zoicb_ctx* ctx) {
    if (!ctx) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_reset_stats: null context");
    ZGUARD(ctx->device);
    ZCUDA(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
    ZCUDA(cudaMemset(ctx->d_stats, 0, sizeof(DeviceStats)), "cudaMemset(stats)");
    return ZOICB_OK;
}

zoicb_status zoicb_get_constants(const zoicb_ctx* ctx, zoicb_constants* out) {
    if (!ctx || !out) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_get_constants: null argument");
    *out = ctx->host.constants;
    return ZOICB_OK;
}

zoicb_status zoicb_get_bokeh_tables(const zoicb_ctx* ctx, float* cdfRow, int32_t* rowIndices, float* cdfColumn,
                                    int32_t* columnIndices) {
    if (!ctx) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_get_bokeh_tables: null context");
    const HostBokeh& hb = ctx->host.bokeh;
    if (!hb.valid()) return api_fail(ZOICB_ERR_BOKEH_IMAGE, "zoicb_get_bokeh_tables: camera has no bokeh tables (no image, or fewer than 3 channels: the reference builds none)");
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
    if (!params || !out) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_setup_host_only: null argument");
    HostCamera hc;
    std::string err;
    zoicb_status rc = build_camera(*params, rgb, width, height, nch, &hc, &err, nullptr, nullptr);
    if (rc != ZOICB_OK) return api_fail(rc, "zoicb_setup_host_only: " + err);
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
        return api_fail(ZOICB_ERR_CUDA, "zoicb_build_bokeh_tables: no CUDA device (libzoicb has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_build_bokeh_tables: bad device index");
    ZGUARD(device);
    std::string err;
    zoicb_status rc = check_bokeh_image(rgb, width, height, nch, &err);
    if (rc != ZOICB_OK) return api_fail(rc, "zoicb_build_bokeh_tables: " + err);
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
    if (!ok) return api_fail(ZOICB_ERR_CUDA, "zoicb_build_bokeh_tables: device build failed");
    const size_t np = (size_t)width * height;
    if (cdfRow) std::memcpy(cdfRow, hb.cdf_row.data(), height * sizeof(float));
    if (rowIndices) std::memcpy(rowIndices, hb.row_indices.data(), height * sizeof(int32_t));
    if (cdfColumn) std::memcpy(cdfColumn, hb.cdf_column.data(), np * sizeof(float));
    if (columnIndices) std::memcpy(columnIndices, hb.column_indices.data(), np * sizeof(int32_t));
    return ZOICB_OK;
}

zoicb_status zoicb_debug_sort_orders(const float* values, int32_t n, int32_t* restated, int32_t* library) {
    if (n < 0 || (n > 0 && !values)) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_debug_sort_orders: bad argument");
    if (restated) {
        for (int32_t i = 0; i < n; ++i) restated[i] = i;
        gnusort::sort(restated, (long)n, gnusort::Before<int32_t>{values});
    }
    if (library) std_sort_desc(values, n, library);
    return ZOICB_OK;
}

zoicb_status zoicb_debug_lut_boxes(int device, const uint32_t* draws, const uint8_t* accept, int32_t n_film, int32_t per_film,
                                   float first_aperture, float* boxes_device, float* boxes_host) {
    if (!draws || !accept || n_film <= 0 || per_film <= 0)
        return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_debug_lut_boxes: bad argument");
    if (boxes_host) lut_fold_boxes_host(draws, accept, n_film, per_film, first_aperture, boxes_host);
    if (!boxes_device) return ZOICB_OK;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        cudaGetLastError();
        return api_fail(ZOICB_ERR_CUDA, "zoicb_debug_lut_boxes: no such CUDA device (libzoicb has no CPU fallback)");
    }
    ZGUARD(device);
    const size_t total = (size_t)n_film * per_film;
    uint32_t* d_draws = nullptr;
    uint8_t* d_acc = nullptr;
    float4* d_boxes = nullptr;
    cudaError_t e = cudaSuccess;
    int launches = 0;
    do {
        if ((e = cudaMalloc(&d_draws, total * 2 * sizeof(uint32_t))) != cudaSuccess) break;
        if ((e = cudaMalloc(&d_acc, total)) != cudaSuccess) break;
        if ((e = cudaMalloc(&d_boxes, n_film * sizeof(float4))) != cudaSuccess) break;
        if ((e = cudaMemcpy(d_draws, draws, total * 2 * sizeof(uint32_t), cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if ((e = cudaMemcpy(d_acc, accept, total, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if ((e = launch_lut_bbox(d_draws, d_acc, n_film, per_film, first_aperture, d_boxes, nullptr, &launches)) != cudaSuccess) break;
        e = cudaMemcpy(boxes_device, d_boxes, n_film * sizeof(float4), cudaMemcpyDeviceToHost);
    } while (0);
    cudaFree(d_draws); cudaFree(d_acc); cudaFree(d_boxes);
    api_count_launches(launches);
    if (e != cudaSuccess) return api_cuda_fail(e, "zoicb_debug_lut_boxes");
    return ZOICB_OK;
}

float zoicb_debug_sqrt_threshold(float radius) { return sqrt_threshold(radius); }

zoicb_status zoicb_get_create_times(const zoicb_ctx* ctx, double* total_ms, double* lut_ms, double* bokeh_ms) {
    if (!ctx) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_get_create_times: null context");
    if (total_ms) *total_ms = ctx->create_ms;
    if (lut_ms) *lut_ms = ctx->lut_ms;
    if (bokeh_ms) *bokeh_ms = ctx->bokeh_ms;
    return ZOICB_OK;
}

zoicb_status zoicb_debug_check_normalize_factor(int device, uint64_t* mismatches, uint32_t* first_bad) {
    if (!mismatches || !first_bad) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_debug_check_normalize_factor: null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        cudaGetLastError();
        return api_fail(ZOICB_ERR_CUDA, "zoicb_debug_check_normalize_factor: no such CUDA device (libzoicb has no CPU fallback)");
    }
    ZGUARD(device);
    unsigned long long bad = 0;
    unsigned first = 0;
    int launches = 0;
    const cudaError_t e = check_normalize_factor(&bad, &first, &launches);
    api_count_launches(launches);
    if (e != cudaSuccess) return api_cuda_fail(e, "zoicb_debug_check_normalize_factor");
    *mismatches = bad;
    *first_bad = first;
    return ZOICB_OK;
}

zoicb_status zoicb_measure_fp32_peak(int device, double* tflops) {
    if (!tflops) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_measure_fp32_peak: null argument");
    ZGUARD(device);
    int launches = 0;
    cudaError_t e = measure_fp32_peak(tflops, &launches);
    api_count_launches(launches);
    if (e != cudaSuccess) return api_cuda_fail(e, "zoicb_measure_fp32_peak");
    return ZOICB_OK;
}

uint64_t zoicb_kernel_launches(void) { return g_launches.load(std::memory_order_relaxed); }
const char* zoicb_last_error(void) { return g_last_error.c_str(); }
const char* zoicb_version(void) { return "zoicb 0.2 (sm_100a)"; }

}  // extern "C"
