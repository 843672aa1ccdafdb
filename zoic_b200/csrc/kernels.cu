// kernels.cu -- sm_100a kernels for batched camera_create_ray.
//
// Behavioural reference: camera_create_ray, reference src/zoic.cpp:1752-1990 (thin-lens branch
// :1771-1848, raytraced branch :1850-1964, tail :1974-1987) and its callees.
//
// Data layout in HBM: samples float4 (sx, sy, lensx, lensy); outputs two float4
// arrays (origin.xyz, weight) and (dir.xyz, tries), index = sample index, so a warp reads 512 B and
// writes 2 x 512 B contiguous.  Camera constants travel as a __grid_constant__ kernel parameter.
#include <cuda_runtime.h>
#include <math_constants.h>

#include <cstdlib>

#include "kernels.h"
#include "lens_math.cuh"

namespace zoicb {

// ------------------------------------------------------------------------------------------------
// image-based aperture sampling (reference imageData::bokehSample, src/zoic.cpp:420-485)
// ------------------------------------------------------------------------------------------------
// std::upper_bound over a[0..n): first index whose value is greater than u, with libstdc++'s probe sequence
// (first/len halving).  The loop runs a warp-uniform number of rounds (bit length of n) with predicated
// updates instead of a per-lane trip count: no divergence, and -- the reason it is written this way -- no
// lane leaves the loop early.  (With a data-dependent trip count ptxas 12.9 let the early lanes run ahead and
// re-use the uniform registers that hold the table pointers while the late lanes were still reading them.)
template <typename Load>
__device__ __forceinline__ int upper_bound_rounds(int n, float u, Load load) {
    int first = 0, len = n;
    const int rounds = 32 - __clz(n);  // len halves every round: n -> 0 in at most bit_length(n) rounds
    for (int it = 0; it < rounds; ++it) {
        const int half = len >> 1;
        const int mid = first + half;
        const float v = load(mid < n ? mid : n - 1);
        const bool live = len > 0;
        const bool left = u < v;
        first = (live && !left) ? mid + 1 : first;
        len = live ? (left ? half : len - half - 1) : 0;
    }
    return first;
}

// The row tables (cdfRow, rowIndices: 8 bytes per image row) are always staged in dynamic shared memory --
// s_rows[0..h) holds the CDF, s_rows[h..2h) the row indices -- and addressed as shared memory (no generic
// pointers); the per-row column tables stay in global memory (L1/L2 resident).
extern __shared__ float s_rows[];

struct BokehView {
    const float* cdf_col;     // global
    const uint16_t* rel_col;
    int w, h;
};

__device__ __forceinline__ void bokeh_sample(const BokehView& b, float u_row, float u_col, float* dx, float* dy) {
    int r = upper_bound_rounds(b.h, u_row, [&](int i) { return s_rows[i]; });
    if (r >= b.h) r = b.h - 1;
    const int row = __float_as_int(s_rows[b.h + r]);
    const int rrow = row - ((b.w - 1) / 2);  // centred with the WIDTH (:441)
    const int start = row * b.w;
    const float* __restrict__ col = b.cdf_col + start;
    int c = upper_bound_rounds(b.w, u_col, [&](int i) { return __ldg(col + i); });
    if (c >= b.w) c = b.w - 1;
    const int rel = (int)__ldg(b.rel_col + start + c);
    const int rcol = rel - ((b.h - 1) / 2);  // centred with the HEIGHT (:466)
    const float fr = (float)rcol;
    const float fc = xmul((float)rrow, -1.0f);
    *dx = xmul(xdiv(fr, (float)b.w), 2.0f);
    *dy = xmul(xdiv(fc, (float)b.h), 2.0f);
}

template <bool kImage>
__device__ __forceinline__ void lens_sample(const BokehView& b, float u, float v, float* lx, float* ly) {
    if (kImage) bokeh_sample(b, u, v, lx, ly);
    else concentric_disk(u, v, lx, ly);
}

// two draws of the per-sample stream; the FIRST draw feeds the SECOND parameter (g++ evaluates the
// reference's argument lists right to left; pinned in tests/test_oracle_port_vs_ref.py)
__device__ __forceinline__ void draw_pair(Xor128& rng, float* first_param, float* second_param) {
    uint32_t k1 = xor128_next(rng);
    uint32_t k2 = xor128_next(rng);
    *second_param = u32_to_unit(k1);
    *first_param = u32_to_unit(k2);
}

__device__ __forceinline__ BokehView stage_bokeh(const CameraState& cam) {
    BokehView b;
    b.w = cam.bokeh.w; b.h = cam.bokeh.h;
    b.cdf_col = cam.bokeh.cdf_column;
    b.rel_col = cam.bokeh.rel_column;
    for (int i = threadIdx.x; i < b.h; i += blockDim.x) {
        s_rows[i] = cam.bokeh.cdf_row[i];
        s_rows[b.h + i] = __int_as_float(cam.bokeh.row_indices[i]);
    }
    __syncthreads();
    return b;
}

// ------------------------------------------------------------------------------------------------
// per-block counter reduction: warp shuffle -> shared -> one atomicAdd per counter per block
// ------------------------------------------------------------------------------------------------
struct LocalStats { unsigned rays, success, vignetted, tir, attempts, visits, reruns; };

__device__ __forceinline__ void flush_stats(const LocalStats& ls, DeviceStats* g) {
    __shared__ unsigned long long s_acc[7];
    if (threadIdx.x < 7) s_acc[threadIdx.x] = 0ull;
    __syncthreads();
    unsigned v[7] = {ls.rays, ls.success, ls.vignetted, ls.tir, ls.attempts, ls.visits, ls.reruns};
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        unsigned s = __reduce_add_sync(0xffffffffu, v[k]);
        if ((threadIdx.x & 31) == 0 && s) atomicAdd(&s_acc[k], (unsigned long long)s);
    }
    __syncthreads();
    if (threadIdx.x < 7 && s_acc[threadIdx.x]) {
        unsigned long long* dst = &g->rays + threadIdx.x;
        atomicAdd(dst, s_acc[threadIdx.x]);
    }
}

// ------------------------------------------------------------------------------------------------
// EXACT thin lens, one sample (src/zoic.cpp:1771-1848, :1297-1305)
// ------------------------------------------------------------------------------------------------
template <bool kImage>
__device__ __forceinline__ void thin_exact_sample(const CameraState& cam, const BokehView& bk, float4 s, uint64_t gidx,
                                                  uint64_t seed, float4* o4, float4* d4, LocalStats& ls) {
    const ThinState& T = cam.thin;
    Vec3 p = vmake(xmul(s.x, T.tan_fov), xmul(s.y, T.tan_fov), 1.0f);
    const Vec3 dir0 = vnormalize(p);  // p - origin0 with origin0 = 0
    Vec3 origin = vmake(0.0f, 0.0f, 0.0f);
    Vec3 dir = dir0;
    int tries = 0;
    float weight = 1.0f;
    ls.rays++;
    ls.attempts++;
    if (T.use_dof) {
        float lx, ly;
        lens_sample<kImage>(bk, s.z, s.w, &lx, &ly);
        const float inter = fabsf(xdiv(T.focal_distance, dir0.z));
        const Vec3 focus = vscale(dir0, inter);
        origin = vmake(xmul(lx, T.aperture_radius), xmul(ly, T.aperture_radius), 0.0f);
        dir = vnormalize(vsub(focus, origin));
        if (T.use_ov) {
            Xor128 rng = sample_stream(seed, gidx);
            while (tries <= kMaxTries) {
                // empericalOpticalVignetting
                float qx = xsub(xmul(dir.x, T.ov_distance), origin.x);
                float qy = xsub(xmul(dir.y, T.ov_distance), origin.y);
                float hyp = xsqrt(xadd(xmul(qx, qx), xmul(qy, qy)));
                if (fabsf(hyp) < T.ov_radius_true) break;
                float u, v;
                draw_pair(rng, &u, &v);
                lens_sample<kImage>(bk, u, v, &lx, &ly);
                origin = vmake(xmul(lx, T.aperture_radius), xmul(ly, T.aperture_radius), 0.0f);
                dir = vnormalize(vsub(focus, origin));
                ++tries;
                ls.attempts++;
            }
        }
        if (tries > kMaxTries) { weight = 0.0f; ls.vignetted++; }
        else ls.success++;
    }
    dir.z = -dir.z;
    weight = xmul(weight, cam.weight_scale);
    *o4 = make_float4(origin.x, origin.y, origin.z, weight);
    *d4 = make_float4(dir.x, dir.y, dir.z, (float)tries);
}

// ------------------------------------------------------------------------------------------------
// EXACT raytraced lens, one sample (src/zoic.cpp:1850-1964, :1099-1158)
// ------------------------------------------------------------------------------------------------
template <bool kImage, bool kLut>
__device__ __forceinline__ void kolb_exact_sample(const CameraState& cam, const BokehView& bk, float4 s, uint64_t gidx,
                                                  uint64_t seed, float4* o4, float4* d4, LocalStats& ls) {
    const LensState& L = cam.lens;
    const KolbSampleState k = kolb_sample_setup<kLut, true>(L, s.x, s.y);
    float lx, ly;
    lens_sample<kImage>(bk, s.z, s.w, &lx, &ly);
    Ray r;
    r.o = vmake(k.fx, k.fy, L.origin_shift);
    r.d = kolb_aim<kLut>(L, k, lx, ly, false);
    int tries = 0;
    Xor128 rng = sample_stream(seed, gidx);
    ls.rays++;
    for (;;) {
        int visited;
        const int rc = exact_march(L, r, &visited);
        ls.attempts++;
        ls.visits += visited;
        if (rc == kTir) ls.tir++;
        if (rc == kPass || tries > kMaxTries) break;
        float u, v;
        draw_pair(rng, &u, &v);
        lens_sample<kImage>(bk, u, v, &lx, &ly);
        r.o = vmake(k.fx, k.fy, L.origin_shift);
        r.d = kolb_aim<kLut>(L, k, lx, ly, true);
        ++tries;
    }
    float weight = 1.0f;
    if (tries > kMaxTries) { weight = 0.0f; ls.vignetted++; }
    else ls.success++;
    weight = xmul(weight, cam.weight_scale);
    // flip to look down -Z (:1960-1961)
    *o4 = make_float4(-r.o.x, -r.o.y, -r.o.z, weight);
    *d4 = make_float4(-r.d.x, -r.d.y, -r.d.z, (float)tries);
}

// ------------------------------------------------------------------------------------------------
// EXACT kernels: one thread per sample, grid-stride
// ------------------------------------------------------------------------------------------------
template <int kModel, bool kImage, bool kLut>
__global__ void __launch_bounds__(256)
exact_kernel(const __grid_constant__ CameraState cam, const float4* __restrict__ samples, uint64_t n,
             uint64_t first_index, uint64_t seed, float4* __restrict__ origin_w, float4* __restrict__ dir_tries,
             DeviceStats* stats, int stage_rows) {
    BokehView bk;
    if (kImage) bk = stage_bokeh(cam);
    (void)stage_rows;
    LocalStats ls = {0, 0, 0, 0, 0, 0, 0};
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float4 s = __ldcs(samples + i);
        float4 o4, d4;
        if (kModel == 0) thin_exact_sample<kImage>(cam, bk, s, first_index + i, seed, &o4, &d4, ls);
        else kolb_exact_sample<kImage, kLut>(cam, bk, s, first_index + i, seed, &o4, &d4, ls);
        __stcs(origin_w + i, o4);
        __stcs(dir_tries + i, d4);
    }
    flush_stats(ls, stats);
}

// Re-run of the samples the guarded kernels could not decide: one thread per queued sample index.
template <int kModel, bool kImage, bool kLut>
__global__ void __launch_bounds__(256)
rerun_kernel(const __grid_constant__ CameraState cam, const float4* __restrict__ samples, uint64_t first_index,
             uint64_t seed, float4* __restrict__ origin_w, float4* __restrict__ dir_tries, DeviceStats* stats,
             int stage_rows, const unsigned long long* __restrict__ queue, const unsigned long long* __restrict__ count,
             unsigned long long capacity) {
    BokehView bk;
    if (kImage) bk = stage_bokeh(cam);
    (void)stage_rows;
    LocalStats ls = {0, 0, 0, 0, 0, 0, 0};
    unsigned long long m = *count;
    if (m > capacity) m = capacity;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < m; q += stride) {
        const uint64_t i = queue[q];
        const float4 s = samples[i];
        float4 o4, d4;
        if (kModel == 0) thin_exact_sample<kImage>(cam, bk, s, first_index + i, seed, &o4, &d4, ls);
        else kolb_exact_sample<kImage, kLut>(cam, bk, s, first_index + i, seed, &o4, &d4, ls);
        origin_w[i] = o4;
        dir_tries[i] = d4;
        ls.reruns++;
    }
    flush_stats(ls, stats);
}

// ------------------------------------------------------------------------------------------------
// GUARDED fast path
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float approx_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float approx_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float approx_rsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

enum { kUndecided = 3 };

// The element stack with fused arithmetic (one function so that the compiler sees straight-line code).
// `u` is kept unit length across surfaces (Snell's law maps unit vectors to unit vectors), the intersection
// uses the cancellation-free root of the quadratic, the normal is (c - hit)/R, and the refraction is
// computed unconditionally (eta = 1 at the stop gives u' = u up to 1e-8).  Every accept/reject test carries
// a margin; inside the margin the result is kUndecided.  A lane leaves the stack at the surface that stops
// it, holding the state the reference would hold there (o, u of the last completed surface; the raw aim
// vector if the very first surface stops it; the new origin if it is total internal reflection).
//
// kN > 0: compile-time element count, fully unrolled, element constants become immediate operands;
// kN == 0: run-time count.
template <int kN>
__device__ __forceinline__ int fast_march(const LensState& L, float gscale, Vec3& o, Vec3& u, int* visited) {
    float ox = o.x, oy = o.y, oz = o.z;
    const float dx = u.x, dy = u.y, dz0 = u.z;
    // unit direction: rsqrt + one Newton step
    const float q = fmaf(dx, dx, fmaf(dy, dy, dz0 * dz0));
    float y = approx_rsqrt(q);
    y = y * fmaf(-0.5f * q * y, y, 1.5f);
    float ux = dx * y, uy = dy * y, uz = dz0 * y;
    const float tir_hi = fmaf(1e-4f, gscale, 1.0f), tir_lo = fmaf(-1e-4f, gscale, 1.0f);
    int n = 0, rc = kPass;
    const int count = kN > 0 ? kN : L.count;
#pragma unroll
    for (int i = 0; i < (kN > 0 ? kN : kMaxElements); ++i) {
        if (kN == 0 && i >= count) break;
        const Element& e = L.e[i];
        ++n;
        const float dz = e.vertex - oz;
        const float Lz = e.center - oz;
        const float b = fmaf(ox, ux, oy * uy);
        const float tca = fmaf(Lz, uz, -b);
        // C = |o - c|^2 - radius2 without forming the two large squares: (dz - R)^2 - R^2 = dz (dz - 2R)
        const float C = fmaf(dz, dz - 2.0f * e.radius, fmaf(ox, ox, oy * oy)) + e.r2_corr;
        const float disc = fmaf(tca, tca, -C);
        const float tiny = 1e-5f * gscale * e.radius2;
        const float s = e.sgn * approx_sqrt(fmaxf(disc, 0.0f));
        // t = tca + s; when the two terms cancel use the conjugate root C / (tca - s)
        const float t_conj = C * approx_rcp(tca - s);
        const float t = (tca * s < 0.0f) ? t_conj : tca + s;
        const float hx = fmaf(ux, t, ox), hy = fmaf(uy, t, oy), hz = fmaf(uz, t, oz);
        const float h2 = fmaf(hx, hx, hy * hy);
        const float w = fmaf(hx, ux, hy * uy);
        const float margin = h2 - e.rim2;
        const float guard = fmaf(fabsf(w), e.dt_guard, e.rim2_guard);
        // clean miss, or outside the rim / stop (every grazing hit lands far outside the rim)
        const bool blocked = (disc < -tiny) || (margin > guard);
        const bool unsure = (margin > -guard) || (disc < tiny);
        if (blocked || unsure) {
            rc = blocked ? kBlocked : kUndecided;
            if (i == 0) { ux = dx; uy = dy; uz = dz0; }
            break;
        }
        const float nzr = e.center - hz;
        const float c1 = (w - uz * nzr) * e.inv_radius;       // -(u . n), n = (c - hit)/R
        const float cs2 = fmaf(-e.eta2 * c1, c1, e.eta2);     // <= eta^2: can exceed 1 only where ior_i > ior_next
        const float k = fmaf(e.eta, c1, -approx_sqrt(fabsf(1.0f - cs2)));
        const float kk = k * e.inv_radius;
        ox = hx; oy = hy; oz = hz;                             // the reference moves the origin before Snell (:1130)
        if (cs2 > tir_lo) {
            rc = cs2 > tir_hi ? kTir : kUndecided;
            if (i == 0) { ux = dx; uy = dy; uz = dz0; }
            break;
        }
        ux = fmaf(kk, -hx, e.eta * ux);
        uy = fmaf(kk, -hy, e.eta * uy);
        uz = fmaf(kk, nzr, e.eta * uz);
    }
    o = vmake(ox, oy, oz);
    u = vmake(ux, uy, uz);
    *visited = n;
    return rc;
}

// approximate-division variant of the concentric map (same branch decisions: a, b are computed exactly)
__device__ __forceinline__ void concentric_disk_fast(float ox, float oy, float* lx, float* ly) {
    const float a = two_x_minus_one(ox);
    const float b = two_x_minus_one(oy);
    const bool first = xmul(a, a) > xmul(b, b);
    const float num = first ? b : a, den = first ? a : b;
    const float qt = num * approx_rcp(den);
    const float r = first ? a : b;
    const float phi = first ? 0.78539816339f * qt : 1.57079632679489661923f - 0.78539816339f * qt;
    // phi in [-pi/4, 3pi/4]: phi + pi < 2pi always; (phi + pi/2) + pi may pass 2pi once
    const float two_pi = ZOICB_PI_F * 2.0f;
    const float xs = xsub(xadd(phi, ZOICB_PI_F), ZOICB_PI_F);
    float vc = xadd(xadd(phi, ZOICB_PI_F * 0.5f), ZOICB_PI_F);
    vc = vc >= two_pi ? xsub(vc, two_pi) : vc;
    const float xc = xsub(vc, ZOICB_PI_F);
    *lx = r * parabola_sin(xc);
    *ly = r * parabola_sin(xs);
}

template <bool kImage>
__device__ __forceinline__ void lens_sample_fast(const BokehView& b, float u, float v, float* lx, float* ly) {
    if (kImage) bokeh_sample(b, u, v, lx, ly);
    else concentric_disk_fast(u, v, lx, ly);
}

// Persistent warps with per-lane ray regeneration: a lane whose sample is finished (passed, exhausted its
// retries, or was handed to the exact re-run queue) immediately takes the next unprocessed sample of the
// warp's current chunk, so every lane runs an attempt in every iteration; chunks of kChunk samples are
// handed out by a global counter.  Outputs go straight to their sample index (sector-merged in L2).
constexpr int kChunk = 2048;

template <int kN, bool kImage, bool kLut>
__global__ void __launch_bounds__(256, 3)
kolb_guarded_kernel(const __grid_constant__ CameraState cam, const float4* __restrict__ samples, uint64_t n,
                    uint64_t first_index, uint64_t seed, float4* __restrict__ origin_w, float4* __restrict__ dir_tries,
                    DeviceStats* stats, int stage_rows, unsigned long long* chunk_counter, unsigned long long* queue,
                    unsigned long long* queue_count, unsigned long long capacity) {
    BokehView bk;
    if (kImage) bk = stage_bokeh(cam);
    (void)stage_rows;
    const LensState& L = cam.lens;
    const unsigned lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    LocalStats ls = {0, 0, 0, 0, 0, 0, 0};
    uint64_t cur = 0, end = 0;  // warp-uniform cursor over the current chunk
    bool exhausted = false;     // warp-uniform: the global counter ran past n
    bool have = false;          // this lane holds a sample
    bool fresh = false;         // ... whose first attempt has not run yet
    uint64_t idx = 0;
    KolbSampleState k;
    k.fx = k.fy = k.max_scale = k.translation = k.sn = 0.0f; k.cs = 1.0f;
    Xor128 rng = {0, 0, 0, 0};
    int tries = 0;
    unsigned s_attempts = 0, s_visits = 0, s_tir = 0;  // counters of the sample in flight
    float ua = 0.0f, ub = 0.0f;                        // unit-square point of the next attempt

    for (;;) {
        // ---- phase 1a: lanes without a sample take the next ones of the chunk
        const unsigned need = __ballot_sync(0xffffffffu, !have);
        if (need) {
            if (cur == end && !exhausted) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(chunk_counter, (unsigned long long)kChunk);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base >= n) { exhausted = true; }
                else { cur = base; end = (base + kChunk < n) ? base + kChunk : n; }
            }
            const unsigned avail = (unsigned)(end - cur);
            const unsigned want = __popc(need);
            const unsigned take = want < avail ? want : avail;
            const unsigned rank = __popc(need & lt_mask);
            if (!have && rank < take) {
                idx = cur + rank;
                const float4 s = __ldcs(samples + idx);
                k = kolb_sample_setup<kLut, false>(L, s.x, s.y);
                ua = s.z;
                ub = s.w;
                tries = 0;
                s_attempts = s_visits = s_tir = 0;
                have = true;
                fresh = true;
            }
            cur += take;
        }
        if (!__any_sync(0xffffffffu, have)) {
            if (exhausted) break;
            continue;
        }
        // ---- phase 1b: lanes whose last attempt failed draw the next lens point (lazy stream seeding)
        if (have && !fresh) {
            if (tries == 0) rng = sample_stream(seed, first_index + idx);
            draw_pair(rng, &ua, &ub);
            ++tries;
        }
        // ---- phase 2: one attempt per lane, all lanes together
        float lx, ly;
        lens_sample_fast<kImage>(bk, ua, ub, &lx, &ly);
        Vec3 o = vmake(k.fx, k.fy, L.origin_shift);
        Vec3 u = kolb_aim<kLut>(L, k, lx, ly, !fresh);
        int visited = 0, rc = kBlocked;
        if (have) rc = fast_march<kN>(L, cam.guard_scale, o, u, &visited);
        fresh = false;
        // ---- phase 3: outcome
        if (have) {
            s_attempts++;
            s_visits += visited;
            if (rc == kTir) s_tir++;
        }
        const bool undecided = have && rc == kUndecided;
        const unsigned umask = __ballot_sync(0xffffffffu, undecided);
        if (umask) {
            unsigned long long base = 0;
            const int leader = __ffs(umask) - 1;
            if ((int)lane == leader) base = atomicAdd(queue_count, (unsigned long long)__popc(umask));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (undecided) {
                const unsigned long long pos = base + __popc(umask & lt_mask);
                if (pos < capacity) {
                    queue[pos] = idx;
                } else {  // queue full: settle it here, exactly
                    float4 o4, d4;
                    kolb_exact_sample<kImage, kLut>(cam, bk, samples[idx], first_index + idx, seed, &o4, &d4, ls);
                    __stcs(origin_w + idx, o4);
                    __stcs(dir_tries + idx, d4);
                    ls.reruns++;
                }
                have = false;
            }
        }
        if (have && (rc == kPass || tries > kMaxTries)) {
            float weight = 1.0f;
            if (tries > kMaxTries) { weight = 0.0f; ls.vignetted++; }
            else ls.success++;
            weight *= cam.weight_scale;
            __stcs(origin_w + idx, make_float4(-o.x, -o.y, -o.z, weight));
            __stcs(dir_tries + idx, make_float4(-u.x, -u.y, -u.z, (float)tries));
            ls.rays++;
            ls.attempts += s_attempts;
            ls.visits += s_visits;
            ls.tir += s_tir;
            have = false;
        }
    }
    flush_stats(ls, stats);
}

// ------------------------------------------------------------------------------------------------
// GUARDED kernel, two-stage schedule
//
// Lanes are stateless workers; the samples in flight live in a per-warp pool of kPoolSlots slots in shared
// memory.  Stage A = (draw lens point, aim, surfaces [0, split)), stage B = surfaces [split, N).  After each
// stage the warp sorts the slots it just worked on into three stacks with ballot + popc prefix sums -- rays
// that still need an attempt (A), rays that survived stage A (B), free slots (F) -- and the next pass takes
// 32 slots from whichever stack is full enough, so both stages run with (nearly) all lanes busy no matter how
// many attempts die at the rear rim or at the stop.
// ------------------------------------------------------------------------------------------------
constexpr int kPoolSlots = 96;   // 3 x 32: one of the three stacks always holds a full pass (pigeonhole)
constexpr int kWarpsPerCta = 8;

struct alignas(16) WarpPool {
    float4 film[kPoolSlots];   // fx, fy, max_scale, translation
    float4 rot[kPoolSlots];    // sn, cs, first lens point (ua, ub)
    uint4 rng[kPoolSlots];     // per-sample xorshift128 state
    float4 ray0[kPoolSlots];   // stage A -> B: ox, oy, oz, ux
    float4 ray1[kPoolSlots];   //               uy, uz, sample index (bits), packed counters (bits)
    unsigned char qa[kPoolSlots], qb[kPoolSlots], qf[kPoolSlots];
};
// packed counters: tries [0..7] | fresh [8] | tir [9..15] | surface visits [16..31]
__device__ __forceinline__ unsigned pk_tries(unsigned p) { return p & 0xffu; }
__device__ __forceinline__ bool pk_fresh(unsigned p) { return (p >> 8) & 1u; }
__device__ __forceinline__ unsigned pk_tir(unsigned p) { return (p >> 9) & 0x7fu; }
__device__ __forceinline__ unsigned pk_visits(unsigned p) { return p >> 16; }

// surfaces [from, to) of the fused march (same arithmetic as fast_march); from/to are warp-uniform.
// A ray that is stopped leaves with its state DEAD (nothing after the loop reads o/u of a stopped ray), which
// lets the compiler keep the unrolled surfaces in straight-line SSA form without copies at the exits.
template <int kN>
__device__ __forceinline__ int fast_march_range(const LensState& L, float gscale, int from, int to, float& ox, float& oy,
                                                float& oz, float& ux, float& uy, float& uz, int* visited) {
    const float tir_band = 1e-4f * gscale;
    int last = to - 1, rc = kPass;   // index of the last surface entered
    float px = ox, py = oy, pz = oz, vx = ux, vy = uy, vz = uz;
#pragma unroll
    for (int i = 0; i < (kN > 0 ? kN : kMaxElements); ++i) {
        if (i < from) continue;   // warp-uniform
        if (i >= to) break;       // warp-uniform
        const Element& e = L.e[i];
        const float dz = e.vertex - pz;
        const float m2 = e.vertex_m2r - pz;                              // dz - 2R
        const float Lz = e.center - pz;
        const float tca = fmaf(Lz, vz, -fmaf(px, vx, py * vy));
        // C = |o - c|^2 - radius2 = dz (dz - 2R) + ox^2 + oy^2 + (R^2 - fl(R^2))
        const float C = fmaf(dz, m2, fmaf(px, px, fmaf(py, py, e.r2_corr)));
        const float disc = fmaf(tca, tca, -C);
        const float s = e.sgn * approx_sqrt(fmaxf(disc, 0.0f));
        const float t = (tca * s < 0.0f) ? C * approx_rcp(tca - s) : tca + s;   // conjugate root when tca + s cancels
        const float hx = fmaf(vx, t, px), hy = fmaf(vy, t, py), hz = fmaf(vz, t, pz);
        const float w = fmaf(hx, vx, hy * vy);
        const float margin = fmaf(hx, hx, fmaf(hy, hy, -e.rim2));
        const float guard = fmaf(fabsf(w), e.dt_guard, e.rim2_guard);
        if (margin > -guard || disc < e.miss_guard) {   // stopped here, or too close to call
            rc = ((disc < -e.miss_guard) || (margin > guard)) ? kBlocked : kUndecided;
            last = i;
            break;
        }
        const float nzr = e.center - hz;
        const float c1 = (w - vz * nzr) * e.inv_radius;
        const float rad = fmaf(e.eta2 * c1, c1, e.one_m_eta2);   // 1 - cs2, cs2 = eta^2 (1 - c1^2); negative => TIR
        if (rad < tir_band) {
            rc = rad < -tir_band ? kTir : kUndecided;
            last = i;
            break;
        }
        const float kk = fmaf(e.eta, c1, -approx_sqrt(rad)) * e.inv_radius;
        vx = fmaf(kk, -hx, e.eta * vx);
        vy = fmaf(kk, -hy, e.eta * vy);
        vz = fmaf(kk, nzr, e.eta * vz);
        px = hx; py = hy; pz = hz;
    }
    if (rc == kPass) { ox = px; oy = py; oz = pz; ux = vx; uy = vy; uz = vz; }
    *visited = last - from + 1;
    return rc;
}

template <int kN, bool kImage, bool kLut>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 3)
kolb_pool_kernel(const __grid_constant__ CameraState cam, const float4* __restrict__ samples, uint32_t n,
                 uint64_t first_index, uint64_t seed, float4* __restrict__ origin_w, float4* __restrict__ dir_tries,
                 DeviceStats* stats, unsigned long long* chunk_counter, unsigned long long* queue,
                 unsigned long long* queue_count, unsigned long long capacity, uint64_t queue_base) {
    // dynamic shared memory: [bokeh row tables (2h floats, 16-byte aligned)] [one WarpPool per warp]
    BokehView bk;
    if (kImage) bk = stage_bokeh(cam);
    const unsigned rows_bytes = kImage ? ((unsigned)cam.bokeh.h * 8u + 15u) & ~15u : 0u;
    WarpPool& P = reinterpret_cast<WarpPool*>(reinterpret_cast<char*>(s_rows) + rows_bytes)[threadIdx.x >> 5];
    const LensState& L = cam.lens;
    const int count = kN > 0 ? kN : L.count;
    const int split = L.split;
    const unsigned lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    LocalStats ls = {0, 0, 0, 0, 0, 0, 0};
    P.qf[lane] = (unsigned char)lane;
    P.qf[lane + 32] = (unsigned char)(lane + 32);
    P.qf[lane + 64] = (unsigned char)(lane + 64);
    __syncwarp();
    int nA = 0, nB = 0, nF = kPoolSlots;   // warp-uniform stack heights
    uint32_t cur = 0, end = 0;
    bool exhausted = false;

    // push `slot` of every lane with `p` set onto a stack; returns the new height
    auto push = [&](unsigned char* stack, int height, bool p, int slot) {
        const unsigned m = __ballot_sync(0xffffffffu, p);
        if (p) stack[height + __popc(m & lt_mask)] = (unsigned char)slot;
        return height + __popc(m);
    };
    // a finished or abandoned sample: counters, outputs, exact re-run queue.  A sample that ran out of retries
    // gets weight 0 and -- its half-traced state being meaningless in the reference too --
    // the film point as origin and the optical axis as direction.
    auto finish = [&](bool done, bool undecided, uint32_t idx, unsigned packed, float ox, float oy, float oz, float ux,
                      float uy, float uz) {
        if (done) {
            const unsigned tries = pk_tries(packed);
            float weight = 1.0f;
            if (tries > (unsigned)kMaxTries) { weight = 0.0f; ls.vignetted++; ux = 0.0f; uy = 0.0f; uz = 1.0f; }
            else ls.success++;
            weight *= cam.weight_scale;
            __stcs(origin_w + idx, make_float4(-ox, -oy, -oz, weight));
            __stcs(dir_tries + idx, make_float4(-ux, -uy, -uz, (float)tries));
            ls.rays++;
            ls.attempts += tries + 1;
            ls.visits += pk_visits(packed);
            ls.tir += pk_tir(packed);
        }
        const unsigned um = __ballot_sync(0xffffffffu, undecided);
        if (um) {
            unsigned long long base = 0;
            const int leader = __ffs(um) - 1;
            if ((int)lane == leader) base = atomicAdd(queue_count, (unsigned long long)__popc(um));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (undecided) {
                const unsigned long long pos = base + __popc(um & lt_mask);
                if (pos < capacity) {
                    queue[pos] = queue_base + idx;
                } else {  // queue full: settle it here, exactly
                    float4 o4, d4;
                    kolb_exact_sample<kImage, kLut>(cam, bk, samples[idx], first_index + idx, seed, &o4, &d4, ls);
                    __stcs(origin_w + idx, o4);
                    __stcs(dir_tries + idx, d4);
                    ls.reruns++;
                }
            }
        }
    };

    for (;;) {
        // ---------------- pick the next pass: a full warp of work from one of the stacks whenever there is one
        const bool more = !exhausted || cur < end;
        int mode, m = 32;   // mode 0: stage B, 1: stage A, 2: take new samples
        if (nB >= 32) mode = 0;
        else if (nA >= 32) mode = 1;
        else if (more && nF >= 32) mode = 2;
        else if (nB > 0) { mode = 0; m = nB; }      // the tail of the launch: partial passes
        else if (nA > 0) { mode = 1; m = nA; }
        else if (more) mode = 2;
        else break;

        if (mode == 2) {
            // ---------------- new samples: per-sample set-up (film point, LUT, rotation, retry stream) into free slots
            if (cur == end) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(chunk_counter, (unsigned long long)kChunk);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base >= n) { exhausted = true; continue; }
                cur = (uint32_t)base;
                end = (base + kChunk < n) ? (uint32_t)(base + kChunk) : n;
            }
            int take = nF < 32 ? nF : 32;
            if (take > (int)(end - cur)) take = (int)(end - cur);
            if ((int)lane < take) {
                const int slot = P.qf[nF - 1 - lane];
                const uint32_t idx = cur + lane;
                const float4 s = __ldcs(samples + idx);
                const KolbSampleState k = kolb_sample_setup<kLut, false>(L, s.x, s.y);
                const Xor128 g = sample_stream(seed, first_index + idx);
                P.film[slot] = make_float4(k.fx, k.fy, k.max_scale, k.translation);
                P.rot[slot] = make_float4(k.sn, k.cs, s.z, s.w);
                P.rng[slot] = make_uint4(g.x, g.y, g.z, g.w);
                P.ray1[slot] = make_float4(0.0f, 0.0f, __uint_as_float(idx), __uint_as_float(1u << 8));  // fresh, tries 0
                P.qa[nA + lane] = (unsigned char)slot;
            }
            nF -= take;
            nA += take;
            cur += take;
            __syncwarp();
        } else if (mode == 0) {
            // ---------------- stage B: surfaces [split, count) for survivors of stage A
            const bool act = (int)lane < m;
            const int slot = act ? P.qb[nB - 1 - lane] : 0;
            nB -= m;
            float4 r0 = make_float4(0, 0, 0, 0), r1 = make_float4(0, 0, 1, 0);
            if (act) { r0 = P.ray0[slot]; r1 = P.ray1[slot]; }
            float ox = r0.x, oy = r0.y, oz = r0.z, ux = r0.w, uy = r1.x, uz = r1.y;
            const uint32_t idx = __float_as_uint(r1.z);
            unsigned packed = __float_as_uint(r1.w);
            int visited = 0, rc = kPass;
            if (act) {
                rc = fast_march_range<kN>(L, cam.guard_scale, split, count, ox, oy, oz, ux, uy, uz, &visited);
                packed += (unsigned)visited << 16;
                if (rc == kTir) packed += 1u << 9;
            }
            const bool failed = act && (rc == kBlocked || rc == kTir);
            const bool again = failed && pk_tries(packed) <= (unsigned)kMaxTries;
            const bool done = act && (rc == kPass || (failed && !again));
            const bool undecided = act && rc == kUndecided;
            if (again) P.ray1[slot].w = __uint_as_float(packed);
            finish(done, undecided, idx, packed, ox, oy, oz, ux, uy, uz);
            nA = push(P.qa, nA, again, slot);
            nF = push(P.qf, nF, done || undecided, slot);
            __syncwarp();
        } else {
            // ---------------- stage A: lens point, aim, surfaces [0, split)
            const bool act = (int)lane < m;
            const int slot = act ? P.qa[nA - 1 - lane] : 0;
            nA -= m;
            float4 f = make_float4(0, 0, 1, 0), rt = make_float4(0, 1, 0.5f, 0.25f), r1 = make_float4(0, 0, 0, 0);
            uint4 g4 = make_uint4(1, 2, 3, 4);
            if (act) { f = P.film[slot]; rt = P.rot[slot]; g4 = P.rng[slot]; r1 = P.ray1[slot]; }
            const uint32_t idx = __float_as_uint(r1.z);
            unsigned packed = __float_as_uint(r1.w);
            bool fresh = pk_fresh(packed);
            packed &= ~(1u << 8);
            float ua = rt.z, ub = rt.w;
            KolbSampleState k;
            k.fx = f.x; k.fy = f.y; k.max_scale = f.z; k.translation = f.w; k.sn = rt.x; k.cs = rt.y;
            Xor128 g = {g4.x, g4.y, g4.z, g4.w};
            float ox = k.fx, oy = k.fy, oz = L.origin_shift, ux = 0.0f, uy = 0.0f, uz = 1.0f;
            int rc = kPass;
            bool todo = act;   // lanes that still owe an attempt in this pass
            // While at least half the warp was stopped inside stage A, those lanes re-sample right here instead of
            // going round through the stacks (the cheap path for cameras whose attempts mostly die at the rear rim).
            for (;;) {
                if (todo) {
                    if (!fresh) { draw_pair(g, &ua, &ub); packed += 1u; }   // ++tries
                    float lx, ly;
                    lens_sample_fast<kImage>(bk, ua, ub, &lx, &ly);
                    const Vec3 d = kolb_aim<kLut>(L, k, lx, ly, !fresh);
                    const float q = fmaf(d.x, d.x, fmaf(d.y, d.y, d.z * d.z));
                    float y = approx_rsqrt(q);
                    y = y * fmaf(-0.5f * q * y, y, 1.5f);
                    ox = k.fx; oy = k.fy; oz = L.origin_shift; ux = d.x * y; uy = d.y * y; uz = d.z * y;
                    int visited = 0;
                    rc = fast_march_range<kN>(L, cam.guard_scale, 0, split, ox, oy, oz, ux, uy, uz, &visited);
                    packed += (unsigned)visited << 16;
                    if (rc == kTir) packed += 1u << 9;
                    fresh = false;
                }
                todo = todo && (rc == kBlocked || rc == kTir) && pk_tries(packed) <= (unsigned)kMaxTries;
                if (__popc(__ballot_sync(0xffffffffu, todo)) < 16) break;
            }
            g4 = make_uint4(g.x, g.y, g.z, g.w);
            const bool failed = act && (rc == kBlocked || rc == kTir);
            const bool again = failed && pk_tries(packed) <= (unsigned)kMaxTries;
            const bool onward = act && rc == kPass;
            const bool done = failed && !again;
            const bool undecided = act && rc == kUndecided;
            if (again || onward) {
                P.rng[slot] = g4;
                if (onward) P.ray0[slot] = make_float4(ox, oy, oz, ux);
                P.ray1[slot] = make_float4(uy, uz, __uint_as_float(idx), __uint_as_float(packed));
            }
            finish(done, undecided, idx, packed, ox, oy, oz, ux, uy, uz);
            nA = push(P.qa, nA, again, slot);
            nB = push(P.qb, nB, onward, slot);
            nF = push(P.qf, nF, done || undecided, slot);
            __syncwarp();
        }
    }
    flush_stats(ls, stats);
}

// ------------------------------------------------------------------------------------------------
// Thin lens with optical vignetting: the same persistent-warp / per-lane regeneration schedule, EXACT
// arithmetic (the thin-lens attempt has no double-precision step, so exactness costs little): results are
// bit-identical to thin_exact_sample, only the order of work differs.
// ------------------------------------------------------------------------------------------------
template <bool kImage>
__global__ void __launch_bounds__(256, 4)
thin_persistent_kernel(const __grid_constant__ CameraState cam, const float4* __restrict__ samples, uint64_t n,
                       uint64_t first_index, uint64_t seed, float4* __restrict__ origin_w, float4* __restrict__ dir_tries,
                       DeviceStats* stats, int stage_rows, unsigned long long* chunk_counter) {
    BokehView bk;
    if (kImage) bk = stage_bokeh(cam);
    (void)stage_rows;
    const ThinState& T = cam.thin;
    const unsigned lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    LocalStats ls = {0, 0, 0, 0, 0, 0, 0};
    uint64_t cur = 0, end = 0;
    bool exhausted = false, have = false, fresh = false;
    uint64_t idx = 0;
    Vec3 focus = vmake(0.0f, 0.0f, 0.0f);
    Xor128 rng = {0, 0, 0, 0};
    int tries = 0;
    float ua = 0.0f, ub = 0.0f;

    for (;;) {
        const unsigned need = __ballot_sync(0xffffffffu, !have);
        if (need) {
            if (cur == end && !exhausted) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(chunk_counter, (unsigned long long)kChunk);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base >= n) { exhausted = true; }
                else { cur = base; end = (base + kChunk < n) ? base + kChunk : n; }
            }
            const unsigned avail = (unsigned)(end - cur);
            const unsigned want = __popc(need);
            const unsigned take = want < avail ? want : avail;
            const unsigned rank = __popc(need & lt_mask);
            if (!have && rank < take) {
                idx = cur + rank;
                const float4 s = __ldcs(samples + idx);
                const Vec3 dir0 = vnormalize(vmake(xmul(s.x, T.tan_fov), xmul(s.y, T.tan_fov), 1.0f));
                focus = vscale(dir0, fabsf(xdiv(T.focal_distance, dir0.z)));
                ua = s.z;
                ub = s.w;
                tries = 0;
                have = true;
                fresh = true;
            }
            cur += take;
        }
        if (!__any_sync(0xffffffffu, have)) {
            if (exhausted) break;
            continue;
        }
        if (have && !fresh) {
            if (tries == 0) rng = sample_stream(seed, first_index + idx);
            draw_pair(rng, &ua, &ub);
            ++tries;
        }
        fresh = false;
        float lx, ly;
        lens_sample<kImage>(bk, ua, ub, &lx, &ly);
        const Vec3 origin = vmake(xmul(lx, T.aperture_radius), xmul(ly, T.aperture_radius), 0.0f);
        const Vec3 dir = vnormalize(vsub(focus, origin));
        const float qx = xsub(xmul(dir.x, T.ov_distance), origin.x);
        const float qy = xsub(xmul(dir.y, T.ov_distance), origin.y);
        const float hyp = xsqrt(xadd(xmul(qx, qx), xmul(qy, qy)));
        const bool pass = fabsf(hyp) < T.ov_radius_true;
        if (have) {
            ls.attempts++;
            // the reference stops sampling once tries has passed maxtries, whatever the last test said
            if (pass || tries > kMaxTries) {
                float weight = 1.0f;
                if (tries > kMaxTries) { weight = 0.0f; ls.vignetted++; }
                else ls.success++;
                weight = xmul(weight, cam.weight_scale);
                __stcs(origin_w + idx, make_float4(origin.x, origin.y, origin.z, weight));
                __stcs(dir_tries + idx, make_float4(dir.x, dir.y, -dir.z, (float)tries));
                ls.rays++;
                have = false;
            }
        }
    }
    flush_stats(ls, stats);
}

// ------------------------------------------------------------------------------------------------
// synthetic samples
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
synth_samples_kernel(uint32_t W, uint32_t H, uint32_t spp, uint64_t seed, uint64_t first_index, uint64_t n,
                     float4* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        const uint64_t i = first_index + j;
        const uint64_t pix = i / spp;
        const uint32_t px = (uint32_t)(pix % W), py = (uint32_t)((pix / W) % H);
        const uint64_t g0 = mix64((seed ^ 0xA5A5A5A55A5A5A5Aull) + ZOICB_GOLDEN * (i + 1));
        const uint64_t g1 = mix64(g0 + ZOICB_GOLDEN);
        const float inv24 = 1.0f / 16777216.0f;
        const float u0 = xmul((float)(uint32_t)(g0 & 0xFFFFFF), inv24);
        const float u1 = xmul((float)(uint32_t)((g0 >> 32) & 0xFFFFFF), inv24);
        const float u2 = xmul((float)(uint32_t)(g1 & 0xFFFFFF), inv24);
        const float u3 = xmul((float)(uint32_t)((g1 >> 32) & 0xFFFFFF), inv24);
        const float fx = xadd((float)px, u0);
        const float fy = xadd((float)py, u1);
        float4 o;
        o.x = xsub(xdiv(xmul(2.0f, fx), (float)W), 1.0f);
        o.y = xmul(xsub(1.0f, xdiv(xmul(2.0f, fy), (float)H)), xdiv((float)H, (float)W));
        o.z = u2;
        o.w = u3;
        out[j] = o;
    }
}

// ------------------------------------------------------------------------------------------------
// exit-pupil LUT candidates (src/zoic.cpp:1409-1421): classify n_film x per_film rays
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
lut_trace_kernel(const __grid_constant__ LensState L, const float* __restrict__ film_x, int n_film, int per_film,
                 const uint32_t* __restrict__ draws, uint8_t* __restrict__ accept) {
    const size_t total = (size_t)n_film * per_film;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const int f = (int)(idx / per_film);
        const uint2 k = reinterpret_cast<const uint2*>(draws)[idx];
        const float U = xsub(xmul(u32_to_unit(k.x), 2.0f), 1.0f);
        const float V = xsub(xmul(u32_to_unit(k.y), 2.0f), 1.0f);
        Ray r;
        r.o = vmake(film_x[f], 0.0f, L.origin_shift);
        r.d = vmake(xsub(xmul(U, L.first_aperture), r.o.x), xsub(xmul(V, L.first_aperture), r.o.y), L.neg_first_thickness);
        int visited;
        accept[idx] = exact_march(L, r, &visited) == kPass ? 1 : 0;
    }
}

// ------------------------------------------------------------------------------------------------
// fp32 peak probe: 8 independent FFMA chains per thread
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ffma_peak_kernel(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 123.456f) out[0] = s;
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
static int g_sm_count = 0;
static int sm_count() {
    if (!g_sm_count) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (g_sm_count <= 0) g_sm_count = 148;
    }
    return g_sm_count;
}

static unsigned grid_for(uint64_t n, int threads, int ctas_per_sm) {
    uint64_t want = (n + threads - 1) / threads;
    uint64_t cap = (uint64_t)sm_count() * ctas_per_sm;  // whole waves of resident CTAs, grid-stride beyond
    return (unsigned)(want < cap ? (want ? want : 1) : cap);
}

template <int kModel, bool kImage, bool kLut>
static cudaError_t launch_variant(const CameraState& cam, int mode, const float4* samples, uint64_t n, uint64_t first_index,
                                  uint64_t seed, float4* origin_w, float4* dir_tries, DeviceStats* stats, cudaStream_t st,
                                  const Workspace& ws, size_t smem, int stage, int* launches) {
    const int threads = 256;
    if (mode == 1 && kModel == 1) {  // guarded fast path + exact re-run of the undecided samples
        cudaError_t e = cudaMemsetAsync(ws.counters, 0, 2 * sizeof(unsigned long long), st);
        if (e != cudaSuccess) return e;
        const unsigned grid = (unsigned)sm_count() * 3;  // persistent: 3 CTAs of 8 warps per SM
        static const bool use_v2 = getenv("ZOICB_KOLB_V2") != nullptr;  // A/B switch for the single-stage schedule
        if (!use_v2) {
            // two-stage pool kernel; 32-bit sample offsets inside a launch, so very large batches go in slices
            const uint64_t slice = 1ull << 31;
            for (uint64_t b = 0; b < n; b += slice) {
                const uint32_t m = (uint32_t)((n - b < slice) ? n - b : slice);
                if (b) {
                    e = cudaMemsetAsync(ws.counters, 0, sizeof(unsigned long long), st);  // chunk cursor only
                    if (e != cudaSuccess) return e;
                }
#define ZP(N)                                                                                                            \
    do {                                                                                                                 \
        const size_t pool_smem = ((smem + 15) & ~(size_t)15) + kWarpsPerCta * sizeof(WarpPool);                            \
        cudaFuncSetAttribute(kolb_pool_kernel<N, kImage, kLut>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pool_smem);      \
        kolb_pool_kernel<N, kImage, kLut><<<grid, threads, pool_smem, st>>>(cam, samples + b, m, first_index + b, seed, origin_w + b, \
                                                                      dir_tries + b, stats, ws.counters, ws.queue,       \
                                                                      ws.counters + 1, ws.capacity, b);                  \
    } while (0)
                switch (cam.lens.count) {
                    case 7: ZP(7); break;
                    case 8: ZP(8); break;
                    case 9: ZP(9); break;
                    case 11: ZP(11); break;
                    case 12: ZP(12); break;
                    default: ZP(0); break;
                }
#undef ZP
                if (launches) *launches += 1;
            }
            rerun_kernel<kModel, kImage, kLut><<<(unsigned)sm_count() * 2, threads, smem, st>>>(
                cam, samples, first_index, seed, origin_w, dir_tries, stats, stage, ws.queue, ws.counters + 1, ws.capacity);
            if (launches) *launches += 1;
            return cudaGetLastError();
        }
#define ZG(N) kolb_guarded_kernel<N, kImage, kLut><<<grid, threads, smem, st>>>(cam, samples, n, first_index, seed, origin_w, \
                                                                         dir_tries, stats, stage, ws.counters, ws.queue,   \
                                                                         ws.counters + 1, ws.capacity)
        switch (cam.lens.count) {  // unrolled instantiations for the element counts of the shipped lens tables
            case 7: ZG(7); break;
            case 8: ZG(8); break;
            case 9: ZG(9); break;
            case 11: ZG(11); break;
            case 12: ZG(12); break;
            default: ZG(0); break;
        }
#undef ZG
        rerun_kernel<kModel, kImage, kLut><<<(unsigned)sm_count() * 2, threads, smem, st>>>(
            cam, samples, first_index, seed, origin_w, dir_tries, stats, stage, ws.queue, ws.counters + 1, ws.capacity);
        if (launches) *launches += 2;
        return cudaGetLastError();
    }
    if (mode == 1 && kModel == 0 && cam.thin.use_dof && cam.thin.use_ov) {  // retry loop present: persistent schedule
        cudaError_t e = cudaMemsetAsync(ws.counters, 0, 2 * sizeof(unsigned long long), st);
        if (e != cudaSuccess) return e;
        thin_persistent_kernel<kImage><<<(unsigned)sm_count() * 4, threads, smem, st>>>(cam, samples, n, first_index, seed, origin_w,
                                                                                      dir_tries, stats, stage, ws.counters);
        if (launches) *launches += 1;
        return cudaGetLastError();
    }
    exact_kernel<kModel, kImage, kLut><<<grid_for(n, threads, 8), threads, smem, st>>>(cam, samples, n, first_index, seed,
                                                                                     origin_w, dir_tries, stats, stage);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_generate(const CameraState& cam, int mode, const float4* samples, uint64_t n, uint64_t first_index,
                            uint64_t seed, float4* origin_w, float4* dir_tries, DeviceStats* stats, cudaStream_t st,
                            const Workspace& ws, int* launches) {
    if (n == 0) return cudaSuccess;
    if (n < 8192) mode = 0;  // a persistent grid does not pay for a handful of samples; the exact kernel is also bit-exact
    const bool image = cam.use_image != 0;
    size_t smem = 0;
    int stage = 0;
    if (image) {  // row tables staged in shared memory; zoicb_create rejects images with more than kMaxBokehRows rows
        smem = (size_t)cam.bokeh.h * 8;
        stage = 1;
    }
#define ZL(M, I, U) launch_variant<M, I, U>(cam, mode, samples, n, first_index, seed, origin_w, dir_tries, stats, st, ws, smem, stage, launches)
    if (cam.lens_model == 0) return image ? ZL(0, true, false) : ZL(0, false, false);
    const bool lut = cam.lens.use_lut != 0;
    if (image) return lut ? ZL(1, true, true) : ZL(1, true, false);
    return lut ? ZL(1, false, true) : ZL(1, false, false);
#undef ZL
}

cudaError_t launch_synth(uint32_t W, uint32_t H, uint32_t spp, uint64_t seed, uint64_t first_index, uint64_t n,
                         float4* out, cudaStream_t st, int* launches) {
    if (n == 0) return cudaSuccess;
    synth_samples_kernel<<<grid_for(n, 256, 8), 256, 0, st>>>(W, H, spp, seed, first_index, n, out);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_lut_trace(const LensState& L, const float* d_film_x, int n_film, int per_film, const uint32_t* d_draws,
                             uint8_t* d_accept, cudaStream_t st, int* launches) {
    lut_trace_kernel<<<grid_for((uint64_t)n_film * per_film, 256, 8), 256, 0, st>>>(L, d_film_x, n_film, per_film, d_draws, d_accept);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t measure_fp32_peak(double* tflops, int* launches) {
    float* d = nullptr;
    cudaError_t e = cudaMalloc(&d, 4);
    if (e != cudaSuccess) return e;
    const int iters = 4096, blocks = sm_count() * 8, threads = 256;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(a);
        ffma_peak_kernel<<<blocks, threads>>>(d, iters, 1.0000001f, 1e-7f);
        cudaEventRecord(b);
        e = cudaEventSynchronize(b);
        if (e != cudaSuccess) break;
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        double flops = 2.0 * 8 * 16 * (double)iters * blocks * threads;
        double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
        if (launches) *launches += 1;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d);
    *tflops = best;
    return e;
}

}  // namespace zoicb
