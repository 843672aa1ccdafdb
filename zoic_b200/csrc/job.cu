// job.cu -- the tiled job runner behind zoicb_run_job (include/zoicb.h): a whole W x H x spp job of camera_create_ray
// (reference src/zoic.cpp:1752-1990, one call per sample there) streamed through rotating tile buffers on the device.
//
// Per tile: synthesise the samples on the device (synth_samples_kernel), generate the rays (the kernels behind
// zoicb_generate), hand the finished tile to a CONSUMER -- here a checksum kernel standing in for the renderer -- and
// recycle the buffers, so jobs far larger than HBM run at full size (BASELINE config 5 writes 1.09 TB of rays per
// lens).  Three streams overlap the stages: tile k+1 is synthesised and tile k-1 consumed while tile k is generated.
// Optional per tile: the parity CENSUS (the tile generated again in EXACT mode and every record compared on the
// device) and the capture of a few WINDOWS of records for comparison with the CPU oracle.
// With a zoicb_gather (gather.cu) the finished tiles of all ranks land in the consumer rank's buffers over NVLink.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "capi_internal.h"
#include "gather.h"
#include "kernel_common.cuh"

using namespace zoicb;

namespace zoicb {

// ------------------------------------------------------------------------------------------------
// consumer: order-independent checksum of a tile of records + the counts a renderer would see
//   checksum = sum over records of sum_j word_j * K_j  (mod 2^64), K_j odd 32-bit constants
// ------------------------------------------------------------------------------------------------
struct ConsumeTotals { unsigned long long checksum, zero_weight, tries_sum, records; };

__global__ void __launch_bounds__(256)
consume_rays_kernel(const RayRecord* __restrict__ rays, uint64_t n, ConsumeTotals* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long sum = 0, zero = 0, tries = 0, cnt = 0;
    auto load = [&](uint64_t i, unsigned* w) {
        asm volatile("ld.global.cs.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                     : "l"(rays + i));
    };
    auto eat = [&](const unsigned* w) {
        sum += (unsigned long long)w[0] * 0x9E3779B1u + (unsigned long long)w[1] * 0x85EBCA77u +
               (unsigned long long)w[2] * 0xC2B2AE3Du + (unsigned long long)w[3] * 0x27D4EB2Fu +
               (unsigned long long)w[4] * 0x165667B1u + (unsigned long long)w[5] * 0xD3A2646Du +
               (unsigned long long)w[6] * 0xFD7046C5u + (unsigned long long)w[7] * 0xB55A4F09u;
        zero += (__uint_as_float(w[3]) == 0.0f) ? 1u : 0u;
        tries += (unsigned long long)__uint_as_float(w[7]);
        cnt += 1;
    };
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + stride < n; i += 2 * stride) {   // two records in flight per thread
        unsigned a[8], b[8];
        load(i, a);
        load(i + stride, b);
        eat(a);
        eat(b);
    }
    if (i < n) {
        unsigned a[8];
        load(i, a);
        eat(a);
    }
    __shared__ unsigned long long acc[4];
    if (threadIdx.x < 4) acc[threadIdx.x] = 0ull;
    __syncthreads();
    unsigned long long v[4] = {sum, zero, tries, cnt};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        unsigned long long x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(&acc[k], x);
    }
    __syncthreads();
    if (threadIdx.x < 4) atomicAdd(&out->checksum + threadIdx.x, acc[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------------
// parity census: records of the GUARDED mode against the EXACT mode's, every record of a tile.
//   flip       : weight or tries differ (another accept / reject sequence)
//   out_of_tol : a live ray (weight != 0) whose origin moved by more than tol * max(|origin|, 1 cm) or whose direction
//                moved by more than tol (the north-star tolerance, per vector), or turned non-finite on one side only
//   max_rel_origin / max_dir : the largest such distances among live rays (float bits through atomicMax: non-negative)
// Zero-weight rays carry no ray (the reference leaves the half-traced state of the last failed attempt there).
// ------------------------------------------------------------------------------------------------
struct CensusTotals { unsigned long long rays, flips, out_of_tol, live; unsigned max_rel_origin_bits, max_dir_bits, pad0, pad1; };

__global__ void __launch_bounds__(256)
census_kernel(const RayRecord* __restrict__ fast, const RayRecord* __restrict__ exact, uint64_t n, float tol,
              CensusTotals* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned flips = 0, bad = 0, live = 0, cnt = 0;
    float worst_o = 0.0f, worst_d = 0.0f;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float a[8], b[8];
        asm volatile("ld.global.cs.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(a[0]), "=f"(a[1]), "=f"(a[2]), "=f"(a[3]), "=f"(a[4]), "=f"(a[5]), "=f"(a[6]), "=f"(a[7])
                     : "l"(fast + i));
        asm volatile("ld.global.cs.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(b[0]), "=f"(b[1]), "=f"(b[2]), "=f"(b[3]), "=f"(b[4]), "=f"(b[5]), "=f"(b[6]), "=f"(b[7])
                     : "l"(exact + i));
        cnt++;
        const bool flip = !(a[3] == b[3]) || !(a[7] == b[7]);
        if (flip) { flips++; continue; }
        if (b[3] == 0.0f) continue;
        live++;
        // the lens-centre sample (0/0 in the concentric map) is a NaN ray with weight 1 in the reference and in both
        // modes: identical NaN patterns are equal
        bool same_nan = true, any_nan = false;
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            if (k == 3) continue;
            const bool na = a[k] != a[k], nb = b[k] != b[k];
            any_nan |= na || nb;
            same_nan &= na == nb;
        }
        if (any_nan) { if (!same_nan) bad++; continue; }
        const float dox = a[0] - b[0], doy = a[1] - b[1], doz = a[2] - b[2];
        const float ddx = a[4] - b[4], ddy = a[5] - b[5], ddz = a[6] - b[6];
        const float d_o = sqrtf(dox * dox + doy * doy + doz * doz);
        const float d_d = sqrtf(ddx * ddx + ddy * ddy + ddz * ddz);
        const float scale = fmaxf(sqrtf(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]), 1.0f);
        const float rel = d_o / scale;
        if (!(rel <= tol) || !(d_d <= tol)) bad++;
        if (rel == rel) worst_o = fmaxf(worst_o, rel);
        if (d_d == d_d) worst_d = fmaxf(worst_d, d_d);
    }
    unsigned v[4] = {cnt, flips, bad, live};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const unsigned s = __reduce_add_sync(0xffffffffu, v[k]);
        if ((threadIdx.x & 31) == 0 && s) atomicAdd(&out->rays + k, (unsigned long long)s);
    }
    const unsigned mo = __reduce_max_sync(0xffffffffu, __float_as_uint(worst_o));
    const unsigned md = __reduce_max_sync(0xffffffffu, __float_as_uint(worst_d));
    if ((threadIdx.x & 31) == 0) {
        if (mo) atomicMax(&out->max_rel_origin_bits, mo);
        if (md) atomicMax(&out->max_dir_bits, md);
    }
}

static unsigned stream_grid(uint64_t n, int ctas_per_sm) {
    const uint64_t want = (n + 255) / 256, cap = (uint64_t)sm_count() * ctas_per_sm;
    return (unsigned)(want < cap ? (want ? want : 1) : cap);
}

cudaError_t launch_consume(const RayRecord* rays, uint64_t n, void* d_totals, cudaStream_t st, int* launches) {
    if (n == 0) return cudaSuccess;
    static const bool skip = [] { const char* v = getenv("ZOICB_JOB_NO_CONSUME"); return v && atoi(v) != 0; }();   // diagnosis only
    if (skip) return cudaSuccess;
    // 2 CTAs per SM: the consumer runs NEXT to the persistent generate kernel of the following tile, in the registers
    // and warp slots that kernel leaves free
    consume_rays_kernel<<<stream_grid(n, 2), 256, 0, st>>>(rays, n, static_cast<ConsumeTotals*>(d_totals));
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace zoicb

namespace {

// Everything a job allocates, kept in the context between jobs and grown on demand.
struct JobBuffers {
    uint64_t cap_samples = 0, cap_rays = 0, cap_exact = 0, cap_queue = 0;
    int n_samples = 0, n_rays = 0;
    float4* samples[2] = {nullptr, nullptr};
    RayRecord* rays[3] = {nullptr, nullptr, nullptr};
    RayRecord* exact = nullptr;
    Workspace ws = {nullptr, nullptr, 0};   // guarded-mode scratch of the job's own generate stream (private: no context lock)
    DeviceStats* stats = nullptr;        // [0] the job's counters, [1] the census pass's
    ConsumeTotals* consume = nullptr;
    CensusTotals* census = nullptr;
    cudaStream_t s_syn = nullptr, s_gen = nullptr, s_con = nullptr;
    std::vector<cudaEvent_t> plain, timed;   // event pools, handed out per job
    size_t next_plain = 0, next_timed = 0;
    ~JobBuffers() {
        for (auto p : samples) cudaFree(p);
        for (auto p : rays) cudaFree(p);
        cudaFree(exact); cudaFree(stats); cudaFree(consume); cudaFree(census);
        cudaFree(ws.counters); cudaFree(ws.queue);
        for (auto s : {s_syn, s_gen, s_con}) if (s) cudaStreamDestroy(s);
        for (auto e : plain) cudaEventDestroy(e);
        for (auto e : timed) cudaEventDestroy(e);
    }
    void begin_job() { next_plain = next_timed = 0; }
    cudaEvent_t event(bool timing = false) {
        std::vector<cudaEvent_t>& pool = timing ? timed : plain;
        size_t& next = timing ? next_timed : next_plain;
        if (next == pool.size()) {
            cudaEvent_t e = nullptr;
            cudaEventCreateWithFlags(&e, timing ? cudaEventDefault : cudaEventDisableTiming);
            pool.push_back(e);
        }
        return pool[next++];
    }
    // device memory for a job: `nsmp` sample tiles and `nray` ray tiles of `cap` entries (+ one for the census)
    cudaError_t ensure(uint64_t cap, int nsmp, int nray, bool census_pass) {
        cudaError_t e;
        if (!s_gen) {
            if ((e = cudaStreamCreateWithFlags(&s_syn, cudaStreamNonBlocking)) != cudaSuccess) return e;
            if ((e = cudaStreamCreateWithFlags(&s_gen, cudaStreamNonBlocking)) != cudaSuccess) return e;
            if ((e = cudaStreamCreateWithFlags(&s_con, cudaStreamNonBlocking)) != cudaSuccess) return e;
            if ((e = cudaMalloc(&stats, 2 * sizeof(DeviceStats))) != cudaSuccess) return e;
            if ((e = cudaMalloc(&consume, sizeof(ConsumeTotals))) != cudaSuccess) return e;
            if ((e = cudaMalloc(&census, sizeof(CensusTotals))) != cudaSuccess) return e;
            if ((e = cudaMalloc(&ws.counters, 4 * sizeof(unsigned long long))) != cudaSuccess) return e;
        }
        if (cap > cap_samples || nsmp > n_samples) {
            for (auto& p : samples) { cudaFree(p); p = nullptr; }
            cap_samples = 0;
            const uint64_t c = cap > cap_samples ? cap : cap_samples;
            for (int i = 0; i < nsmp; ++i) if ((e = cudaMalloc(&samples[i], c * sizeof(float4))) != cudaSuccess) return e;
            cap_samples = c; n_samples = nsmp;
        }
        if (nray > 0 && (cap > cap_rays || nray > n_rays)) {
            for (auto& p : rays) { cudaFree(p); p = nullptr; }
            cap_rays = 0;
            for (int i = 0; i < nray; ++i) if ((e = cudaMalloc(&rays[i], cap * sizeof(RayRecord))) != cudaSuccess) return e;
            cap_rays = cap; n_rays = nray;
        }
        if (census_pass && cap > cap_exact) {
            cudaFree(exact); exact = nullptr; cap_exact = 0;
            if ((e = cudaMalloc(&exact, cap * sizeof(RayRecord))) != cudaSuccess) return e;
            cap_exact = cap;
        }
        const uint64_t want = cap / 24 + 4096;   // room for the undecided samples of a tile (capi.cu: api_get_workspace)
        if (want > cap_queue) {
            cudaFree(ws.queue); ws.queue = nullptr; cap_queue = 0;
            if ((e = cudaMalloc(&ws.queue, want * sizeof(QueueRecord))) != cudaSuccess) return e;
            cap_queue = want;
        }
        ws.capacity = cap_queue;
        return cudaSuccess;
    }
};

void free_job_buffers(void* p) { delete static_cast<JobBuffers*>(p); }

}  // namespace

extern "C" zoicb_status zoicb_run_job(zoicb_ctx* ctx, const zoicb_job* job, zoicb_job_result* res) {
    if (!ctx || !job || !res) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_run_job: null argument");
    std::memset(res, 0, sizeof *res);
    if (!job->W || !job->H || !job->spp_per_pass) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_run_job: zero dimension");
    if (job->n_windows < 0 || (job->n_windows > 0 && (!job->window_first || !job->d_windows || !job->window_count)))
        return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_run_job: bad window arguments");
    zoicb_gather* g = job->gather;
    const bool gathered = g != nullptr;
    if (job->count == 0 && !gathered) return ZOICB_OK;   // (a rank with an empty share still walks the rounds of a gathered job)
    ZGUARD(ctx->device);
    uint64_t tile = job->tile ? job->tile : (1ull << 28);
    if (gathered) {
        if (gather_device(g) != ctx->device) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_run_job: gather and camera live on different devices");
        if (job->census) return api_fail(ZOICB_ERR_UNSUPPORTED, "zoicb_run_job: the census runs on ungathered jobs");
        tile = gather_tile_rays(g);   // every rank contributes `tile` records per round
    }
    tile = std::min<uint64_t>(tile, 1ull << 30);
    const uint64_t ntiles = (job->count + tile - 1) / tile;
    // all ranks of a gathered job run the same number of rounds (the consumer waits for every rank in every round)
    if (gathered && !job->gather_counts) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_run_job: a gathered job needs gather_counts (every rank's sample count)");
    if (gathered && job->gather_counts[gather_rank(g)] != job->count)
        return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_run_job: gather_counts[this rank] differs from the job's count");
    const uint64_t rounds = gathered ? std::max<uint64_t>(ntiles, gather_rounds(g, job->gather_counts)) : ntiles;

    std::lock_guard<std::mutex> job_lock(ctx->job_mu);
    if (!ctx->job_cache) { ctx->job_cache = new JobBuffers(); ctx->job_cache_free = free_job_buffers; }
    JobBuffers& B = *static_cast<JobBuffers*>(ctx->job_cache);
    B.begin_job();
    const int nray = gathered ? 0 : (job->serial ? 1 : 2);
    const int nsmp = job->serial ? 1 : 2;
    const uint64_t cap = std::min<uint64_t>(tile, job->count);
    {
        const cudaError_t ea = B.ensure(cap, nsmp, nray, job->census != 0);
        if (ea != cudaSuccess) {   // drop the cache: a later, smaller job may still fit
            cudaGetLastError();
            delete &B;
            ctx->job_cache = nullptr;
            return api_cuda_fail(ea, "zoicb_run_job: device memory for the tile buffers");
        }
    }
    ZCUDA(cudaMemsetAsync(B.stats, 0, 2 * sizeof(DeviceStats), B.s_gen), "cudaMemset");
    ZCUDA(cudaMemsetAsync(B.consume, 0, sizeof(ConsumeTotals), B.s_gen), "cudaMemset");
    ZCUDA(cudaMemsetAsync(B.census, 0, sizeof(CensusTotals), B.s_gen), "cudaMemset");
    cudaStream_t s_syn = job->serial ? B.s_gen : B.s_syn, s_gen = B.s_gen, s_con = job->serial ? B.s_gen : B.s_con;

    // events: synthesised[slot], generated[slot], consumed[slot] (a slot is reused two tiles later)
    cudaEvent_t ev_syn[2] = {B.event(), B.event()}, ev_gen[3] = {B.event(), B.event(), B.event()};
    cudaEvent_t ev_con[3] = {B.event(), B.event(), B.event()}, ev_used[2] = {B.event(), B.event()};
    cudaEvent_t t0 = B.event(true), t1 = B.event(true);
    std::vector<cudaEvent_t> g0(ntiles), g1(ntiles);   // around the generate kernels of every tile, on their stream
    for (uint64_t k = 0; k < ntiles; ++k) { g0[k] = B.event(true); g1[k] = B.event(true); }
    double gen_ms = 0.0;
    int launches = 0;
    cudaError_t e = cudaSuccess;
    const char* what = "";

    if (gathered && (e = gather_begin(g, job->gather_counts, s_gen, job->serial != 0)) != cudaSuccess) return api_cuda_fail(e, "zoicb_run_job: gather begin");
    ZCUDA(cudaEventRecord(t0, s_gen), "cudaEventRecord");
    if (!job->serial) {
        ZCUDA(cudaStreamWaitEvent(s_syn, t0, 0), "cudaStreamWaitEvent");
        ZCUDA(cudaStreamWaitEvent(s_con, t0, 0), "cudaStreamWaitEvent");
    }

    auto run = [&]() -> cudaError_t {
        const Workspace ws = B.ws, ws2 = B.ws;   // the census pass runs on the same stream, after the guarded pass
        for (uint64_t k = 0; k < rounds; ++k) {
            const bool have = k < ntiles;
            const uint64_t b = k * tile, m = have ? std::min<uint64_t>(tile, job->count - b) : 0;
            const int ss = (int)(k % nsmp), rs = nray ? (int)(k % nray) : 0;
            RayRecord* dst = nullptr;
            if (gathered) {
                what = "gather acquire";
                if ((e = gather_acquire(g, k, s_gen, &dst)) != cudaSuccess) return e;
            } else {
                dst = B.rays[rs];
            }
            if (have) {
                // samples of tile k (slot ss was last read by the generate of tile k - nsmp)
                what = "synth";
                if (!job->serial && k >= (uint64_t)nsmp && (e = cudaStreamWaitEvent(s_syn, ev_used[ss], 0)) != cudaSuccess) return e;
                if ((e = launch_synth(job->W, job->H, job->spp_per_pass, job->sample_seed, job->first + b, m, B.samples[ss], s_syn, &launches)) != cudaSuccess) return e;
                if (!job->serial) {
                    if ((e = cudaEventRecord(ev_syn[ss], s_syn)) != cudaSuccess) return e;
                    if ((e = cudaStreamWaitEvent(s_gen, ev_syn[ss], 0)) != cudaSuccess) return e;
                    // ray slot rs was last read by the consumer of tile k - nray
                    if (!gathered && k >= (uint64_t)nray && (e = cudaStreamWaitEvent(s_gen, ev_con[rs], 0)) != cudaSuccess) return e;
                }
                what = "generate";
                if ((e = cudaEventRecord(g0[k], s_gen)) != cudaSuccess) return e;
                if ((e = launch_generate(ctx->host.state, ctx->mode, B.samples[ss], m, job->first + b, job->rng_seed, dst, B.stats, s_gen, ws, &launches)) != cudaSuccess) return e;
                if ((e = cudaEventRecord(g1[k], s_gen)) != cudaSuccess) return e;
                if (job->census) {
                    what = "census";
                    if ((e = launch_generate(ctx->host.state, ZOICB_MODE_EXACT, B.samples[ss], m, job->first + b, job->rng_seed, B.exact, B.stats + 1, s_gen, ws2, &launches)) != cudaSuccess) return e;
                    census_kernel<<<stream_grid(m, 4), 256, 0, s_gen>>>(dst, B.exact, m, job->census_tol > 0.0f ? job->census_tol : 1e-5f, B.census);
                    ++launches;
                    if ((e = cudaGetLastError()) != cudaSuccess) return e;
                }
                if (!job->serial && (e = cudaEventRecord(ev_used[ss], s_gen)) != cudaSuccess) return e;
                // windows of records for the oracle comparison
                for (int w = 0; w < job->n_windows; ++w) {
                    const uint64_t wf = job->window_first[w], wl = wf + job->window_count;   // global sample indices
                    const uint64_t lo = std::max(wf, job->first + b), hi = std::min(wl, job->first + b + m);
                    if (lo >= hi) continue;
                    what = "window copy";
                    if ((e = cudaMemcpyAsync(job->d_windows + (size_t)w * job->window_count + (lo - wf), dst + (lo - job->first - b),
                                             (hi - lo) * sizeof(RayRecord), cudaMemcpyDefault, s_gen)) != cudaSuccess) return e;
                }
            }
            if (gathered) {
                what = "gather commit";
                if ((e = gather_commit(g, k, m, s_gen, B.consume, &launches)) != cudaSuccess) return e;
            } else {
                what = "consume";
                if (!job->serial) {
                    if ((e = cudaEventRecord(ev_gen[rs], s_gen)) != cudaSuccess) return e;
                    if ((e = cudaStreamWaitEvent(s_con, ev_gen[rs], 0)) != cudaSuccess) return e;
                }
                if ((e = launch_consume(dst, m, B.consume, s_con, &launches)) != cudaSuccess) return e;
                if (!job->serial && (e = cudaEventRecord(ev_con[rs], s_con)) != cudaSuccess) return e;
            }
            // bound the host's run-ahead to a few tiles (queued work, not correctness: an event re-recorded here was
            // captured by its waiters when they were enqueued)
            what = "tile sync";
            if (k >= 4 && k - 4 < ntiles && (e = cudaEventSynchronize(g1[k - 4])) != cudaSuccess) return e;
        }
        what = "gather end";
        if (gathered && (e = gather_end(g, s_gen)) != cudaSuccess) return e;
        what = "job drain";
        if (!job->serial) {
            cudaEvent_t done = B.event();
            if ((e = cudaEventRecord(done, s_con)) != cudaSuccess) return e;
            if ((e = cudaStreamWaitEvent(s_gen, done, 0)) != cudaSuccess) return e;
        }
        if ((e = cudaEventRecord(t1, s_gen)) != cudaSuccess) return e;
        if ((e = cudaEventSynchronize(t1)) != cudaSuccess) return e;
        what = "generate timing";
        for (uint64_t k = 0; k < ntiles; ++k) {   // the generate kernels' own time, tile by tile
            float ms = 0.0f;
            if ((e = cudaEventElapsedTime(&ms, g0[k], g1[k])) != cudaSuccess) return e;
            gen_ms += ms;
        }
        return cudaSuccess;
    };
    e = run();
    api_count_launches(launches);
    if (e != cudaSuccess) {
        cudaDeviceSynchronize();
        cudaGetLastError();
        return api_cuda_fail(e, what);
    }
    if (gathered && gather_failed(g)) return api_fail(ZOICB_ERR_CUDA, "zoicb_run_job: a gather flag wait timed out (a peer rank fell behind or died)");

    float ms = 0.0f;
    ZCUDA(cudaEventElapsedTime(&ms, t0, t1), "cudaEventElapsedTime");
    ConsumeTotals ct;
    CensusTotals cs;
    DeviceStats st[2];
    ZCUDA(cudaMemcpy(&ct, B.consume, sizeof ct, cudaMemcpyDeviceToHost), "cudaMemcpy(job totals)");
    ZCUDA(cudaMemcpy(&cs, B.census, sizeof cs, cudaMemcpyDeviceToHost), "cudaMemcpy(job totals)");
    ZCUDA(cudaMemcpy(st, B.stats, sizeof st, cudaMemcpyDeviceToHost), "cudaMemcpy(job stats)");
    res->rays = job->count;
    res->tiles = ntiles;
    res->launches = (uint64_t)launches;
    res->device_ms = ms;
    res->generate_ms = (float)gen_ms;
    res->checksum = ct.checksum; res->zero_weight = ct.zero_weight; res->tries_sum = ct.tries_sum; res->consumed = ct.records;
    res->census_rays = cs.rays; res->census_flips = cs.flips; res->census_out_of_tol = cs.out_of_tol; res->census_live = cs.live;
    std::memcpy(&res->census_max_rel_origin, &cs.max_rel_origin_bits, 4);
    std::memcpy(&res->census_max_dir, &cs.max_dir_bits, 4);
    auto put = [](zoicb_stats* o, const DeviceStats& h) {
        o->rays = h.rays; o->success = h.success; o->vignetted = h.vignetted; o->total_internal_reflection = h.tir;
        o->attempts = h.attempts; o->element_visits = h.element_visits; o->exact_reruns = h.exact_reruns;
    };
    put(&res->stats, st[0]);
    put(&res->census_stats, st[1]);
    return ZOICB_OK;
}

// The parity census of two resident ray buffers (zoicb_run_job runs it tile by tile; this entry point serves callers who
// hold both buffers, e.g. the tests): totals are ADDED to *res's census fields.
extern "C" zoicb_status zoicb_census(zoicb_ctx* ctx, const zoicb_ray* d_fast, const zoicb_ray* d_exact, uint64_t n, float tol,
                                     zoicb_job_result* res, void* stream) {
    if (!ctx || !res) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_census: null argument");
    if (n == 0) return ZOICB_OK;
    if (!d_fast || !d_exact) return api_fail(ZOICB_ERR_INVALID_ARGUMENT, "zoicb_census: null buffer");
    ZGUARD(ctx->device);
    CensusTotals* d = nullptr;
    ZCUDA(cudaMalloc(&d, sizeof(CensusTotals)), "cudaMalloc");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(d, 0, sizeof(CensusTotals), st);
    if (e == cudaSuccess) {
        census_kernel<<<stream_grid(n, 8), 256, 0, st>>>((const RayRecord*)d_fast, (const RayRecord*)d_exact, n, tol > 0.0f ? tol : 1e-5f, d);
        api_count_launches(1);
        e = cudaGetLastError();
    }
    CensusTotals cs;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&cs, d, sizeof cs, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d);
    if (e != cudaSuccess) return api_cuda_fail(e, "zoicb_census");
    res->census_rays += cs.rays; res->census_flips += cs.flips; res->census_out_of_tol += cs.out_of_tol; res->census_live += cs.live;
    float mo, md;
    std::memcpy(&mo, &cs.max_rel_origin_bits, 4);
    std::memcpy(&md, &cs.max_dir_bits, 4);
    res->census_max_rel_origin = std::max(res->census_max_rel_origin, mo);
    res->census_max_dir = std::max(res->census_max_dir, md);
    return ZOICB_OK;
}
