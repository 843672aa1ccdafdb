// kernels.cu -- sm_100a kernels for batched camera_create_ray.
//
// Behavioural reference: camera_create_ray, reference src/zoic.cpp:1752-1990 (thin-lens branch
// :1771-1848, raytraced branch :1850-1964, tail :1974-1987) and its callees.
//
// Data layout in HBM: samples float4 (sx, sy, lensx, lensy); outputs ONE 32-byte record
// per sample (origin.xyz, weight, dir.xyz, tries; index = sample index), written with a single 256-bit store.
// Camera constants travel as a __grid_constant__ kernel parameter.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdlib.h>

#include "kernel_common.cuh"

namespace zoicb {

// ------------------------------------------------------------------------------------------------
// EXACT kernels: one thread per sample, grid-stride
// ------------------------------------------------------------------------------------------------
template <int kModel, bool kImage, bool kLut>
__global__ void __launch_bounds__(256)
exact_kernel(const __grid_constant__ CameraState cam, const float4* __restrict__ samples, uint64_t n,
             uint64_t first_index, uint64_t seed, RayRecord* __restrict__ rays,
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
        store_ray(rays, i, o4, d4);
    }
    flush_stats(ls, stats);
}

// ------------------------------------------------------------------------------------------------
// EXACT raytraced lens on the persistent-warp / per-lane regeneration schedule.  Same arithmetic and same
// results as kolb_exact_sample (bit-identical); a lane that finishes its sample takes the next work item at
// once instead of idling until the slowest sample of its warp has used up its retries.  Work items are either
// the samples [0, m) themselves or the entries of the undecided-sample queue of the guarded kernel.
// ------------------------------------------------------------------------------------------------
template <bool kImage, bool kLut, bool kQueued>
__global__ void __launch_bounds__(256, 3)
kolb_exact_persistent_kernel(const __grid_constant__ CameraState cam, const float4* __restrict__ samples, uint64_t m_direct,
                             uint64_t first_index, uint64_t seed, RayRecord* __restrict__ rays,
                             DeviceStats* stats, unsigned long long* cursor,
                             const QueueRecord* __restrict__ queue, const unsigned long long* __restrict__ queue_count,
                             unsigned long long capacity) {
    BokehView bk;
    if (kImage) bk = stage_bokeh(cam);
    const LensState& L = cam.lens;
    const unsigned lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    LocalStats ls = {0, 0, 0, 0, 0, 0, 0};
    unsigned long long m = m_direct;
    if (kQueued) { m = *queue_count; if (m > capacity) m = capacity; }
    constexpr unsigned kGrab = kQueued ? 32 : 256;   // work items per grab of the global cursor (the queue is short: spread it)
    uint64_t cur = 0, end = 0;
    bool exhausted = false, have = false, fresh = false;
    uint64_t idx = 0;
    KolbSampleState k;
    k.fx = k.fy = k.max_scale = k.translation = k.sn = 0.0f; k.cs = 1.0f;
    Xor128 rng = {0, 0, 0, 0};
    int tries = 0;
    float ua = 0.0f, ub = 0.0f;
    for (;;) {
        const unsigned need = __ballot_sync(0xffffffffu, !have);
        if (need) {
            if (cur == end && !exhausted) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(cursor, (unsigned long long)kGrab);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base >= m) { exhausted = true; }
                else { cur = base; end = (base + kGrab < m) ? base + kGrab : m; }
            }
            const unsigned avail = (unsigned)(end - cur);
            const unsigned want = __popc(need);
            const unsigned take = want < avail ? want : avail;
            const unsigned rank = __popc(need & lt_mask);
            if (!have && rank < take) {
                QueueRecord q = 0;
                if (kQueued) {
                    q = queue[cur + rank];
                    idx = queue_index(q);
                } else {
                    idx = cur + rank;
                }
                const float4 s = samples[idx];
                k = kolb_sample_setup<kLut, true>(L, s.x, s.y);
                rng = sample_stream(seed, first_index + idx);
                ua = s.z;
                ub = s.w;
                tries = 0;
                if (kQueued) {
                    // resume at the attempt the fast path could not decide: its earlier attempts were stopped for certain,
                    // so only their draws and their counters are needed (kolb_pool2.cu: enqueue2)
                    tries = (int)queue_tries(q);
                    for (int t = 0; t < tries; ++t) draw_pair(rng, &ua, &ub);
                    ls.attempts += (unsigned)tries;
                    ls.tir += queue_tir(q);
                    ls.visits += queue_visits(q);
                }
                have = true;
                fresh = true;   // the lens sample of this attempt is in (ua, ub) already
                ls.rays++;
                if (kQueued) ls.reruns++;
            }
            cur += take;
        }
        if (!__any_sync(0xffffffffu, have)) {
            if (exhausted) break;
            continue;
        }
        if (have) {
            if (!fresh) { draw_pair(rng, &ua, &ub); ++tries; }
            float lx, ly;
            lens_sample<kImage>(bk, ua, ub, &lx, &ly);
            Ray r;
            r.o = vmake(k.fx, k.fy, L.origin_shift);
            r.d = kolb_aim<kLut>(L, k, lx, ly, tries > 0);   // the retry arithmetic from the first re-sample on (:1933)
            fresh = false;
            int visited;
            const int rc = exact_march(L, r, &visited);
            ls.attempts++;
            ls.visits += visited;
            if (rc == kTir) ls.tir++;
            if (rc == kPass || tries > kMaxTries) {
                float weight = 1.0f;
                if (tries > kMaxTries) { weight = 0.0f; ls.vignetted++; }
                else ls.success++;
                weight = xmul(weight, cam.weight_scale);
                store_ray(rays, idx, make_float4(-r.o.x, -r.o.y, -r.o.z, weight), make_float4(-r.d.x, -r.d.y, -r.d.z, (float)tries));
                have = false;
            }
        }
    }
    flush_stats(ls, stats);
}

// ------------------------------------------------------------------------------------------------
// Thin lens with optical vignetting: the same persistent-warp / per-lane regeneration schedule, EXACT
// arithmetic (the thin-lens attempt has no double-precision step, so exactness costs little): results are
// bit-identical to thin_exact_sample, only the order of work differs.
// ------------------------------------------------------------------------------------------------
#ifndef ZOICB_THIN_CTAS
#define ZOICB_THIN_CTAS 6   // resident CTAs of 8 warps per SM: 48 warps at <= 40 registers (4 -> 5 -> 6: 14.7 -> 16.3 -> 17.2 Grays/s on
                           // config 3, profiles/r01b_ab.txt; 8 spills)
#endif
#ifndef ZOICB_THIN_CTAS_COMPACT
#define ZOICB_THIN_CTAS_COMPACT 8   // the kernel with byte-wide tables fits 32 registers without a spill: 64 warps per SM
                                   // (6 / 7 / 8 CTAs: 25.4 / 25.6 / 26.1 Grays/s at 32 spp, profiles/r02_ab.txt call 30)
#endif
#ifndef ZOICB_THIN_MERGED_NORM
#define ZOICB_THIN_MERGED_NORM 1
#endif
#if ZOICB_THIN_MERGED_NORM
#define ZOICB_THIN_NORMALIZE vnormalize_merged   // lens_math.cuh: same floats, one range check instead of three branches
#else
#define ZOICB_THIN_NORMALIZE vnormalize
#endif
// kCompact: byte-wide column tables and the rows' final CDF values in shared memory (camera_state.h: BokehCompact)
template <bool kImage, bool kCompact>
__global__ void __launch_bounds__(256, kCompact ? ZOICB_THIN_CTAS_COMPACT : ZOICB_THIN_CTAS)
thin_persistent_kernel(const __grid_constant__ CameraState cam, const float4* __restrict__ samples, uint64_t n,
                       uint64_t first_index, uint64_t seed, RayRecord* __restrict__ rays,
                       DeviceStats* stats, int stage_rows, unsigned long long* chunk_counter) {
    BokehView bk;
    if (kImage) bk = kCompact ? stage_bokeh_compact(cam) : stage_bokeh(cam);
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
                const Vec3 dir0 = ZOICB_THIN_NORMALIZE(vmake(xmul(s.x, T.tan_fov), xmul(s.y, T.tan_fov), 1.0f));
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
        lens_sample<kImage, kCompact>(bk, ua, ub, &lx, &ly);
        const Vec3 origin = vmake(xmul(lx, T.aperture_radius), xmul(ly, T.aperture_radius), 0.0f);
        const Vec3 dir = ZOICB_THIN_NORMALIZE(vsub(focus, origin));
        const float qx = xsub(xmul(dir.x, T.ov_distance), origin.x);
        const float qy = xsub(xmul(dir.y, T.ov_distance), origin.y);
        // the reference's sqrt(s) < ov_radius_true, decided on s itself (camera_state.h: ov_s_threshold)
        const bool pass = xadd(xmul(qx, qx), xmul(qy, qy)) < T.ov_s_threshold;
        if (have) {
            ls.attempts++;
            // the reference stops sampling once tries has passed maxtries, whatever the last test said
            if (pass || tries > kMaxTries) {
                float weight = 1.0f;
                if (tries > kMaxTries) { weight = 0.0f; ls.vignetted++; }
                else ls.success++;
                weight = xmul(weight, cam.weight_scale);
                store_ray(rays, idx, make_float4(origin.x, origin.y, origin.z, weight), make_float4(dir.x, dir.y, -dir.z, (float)tries));
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
// The pixel of sample i is (i / spp) mod (W H): three 64-bit divisions per sample made this kernel compute-bound (0.82 ms
// per 2^27 samples, 2.6 TB/s of stores).  A thread walks its samples with a fixed stride, so it divides ONCE and then
// carries (sample-in-pixel, px, py) forward with 32-bit adds: stride = (a1 W H' ... ) decomposed on entry.
__global__ void __launch_bounds__(256)
synth_samples_kernel(uint32_t W, uint32_t H, uint32_t spp, uint64_t seed, uint64_t first_index, uint64_t n,
                     float4* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    // where this thread starts
    uint64_t i = first_index + j;
    const uint64_t pix0 = i / spp;
    uint32_t sp = (uint32_t)(i - pix0 * spp);
    uint32_t px = (uint32_t)(pix0 % W), py = (uint32_t)((pix0 / W) % H);
    // stride = ((d_py * W) + d_px) * spp + d_sp  (mod W H spp)
    const uint64_t spix = stride / spp;
    const uint32_t d_sp = (uint32_t)(stride - spix * spp);
    const uint32_t d_px = (uint32_t)(spix % W), d_py = (uint32_t)((spix / W) % H);
    const float inv24 = 1.0f / 16777216.0f;
    const float fW = (float)W, fH = (float)H, aspect = xdiv(fH, fW);
    for (;;) {
        const uint64_t g0 = mix64((seed ^ 0xA5A5A5A55A5A5A5Aull) + ZOICB_GOLDEN * (i + 1));
        const uint64_t g1 = mix64(g0 + ZOICB_GOLDEN);
        const float u0 = xmul((float)(uint32_t)(g0 & 0xFFFFFF), inv24);
        const float u1 = xmul((float)(uint32_t)((g0 >> 32) & 0xFFFFFF), inv24);
        const float u2 = xmul((float)(uint32_t)(g1 & 0xFFFFFF), inv24);
        const float u3 = xmul((float)(uint32_t)((g1 >> 32) & 0xFFFFFF), inv24);
        const float fx = xadd((float)px, u0);
        const float fy = xadd((float)py, u1);
        float4 o;
        o.x = xsub(xdiv(xmul(2.0f, fx), fW), 1.0f);
        o.y = xmul(xsub(1.0f, xdiv(xmul(2.0f, fy), fH)), aspect);
        o.z = u2;
        o.w = u3;
        out[j] = o;
        j += stride;
        if (j >= n) break;
        i += stride;
        sp += d_sp;
        uint32_t carry = 0;
        if (sp >= spp) { sp -= spp; carry = 1; }
        px += d_px + carry;
        carry = 0;
        if (px >= W) { px -= W; carry = 1; }
        py += d_py + carry;
        if (py >= H) py -= H;
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
// exit-pupil LUT bounding boxes (src/zoic.cpp:1421-1440).  The reference folds the accepted
// candidates of a film position into its bounding box IN ORDER, and the fold is not a plain min/max: whenever the
// running min.x + min.y is exactly 0 -- always for the first accepted candidate (the box starts at the origin),
// and again later if the two minima happen to cancel -- the box is RE-ARMED to the current point (:1423).  One warp
// per film position walks its candidates 32 at a time: warp prefix minima / maxima by shuffles give every lane the
// box BEFORE its candidate; if no accepted lane sees the re-arm condition the chunk is an ordinary min/max (and a
// re-arm inside the chunk would have been seen by its own lane, whose prefix is still exact), otherwise lane 0 replays
// the 32 candidates with the reference's statement.  box = (min.x, min.y, max.x, max.y) per film position.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
lut_bbox_kernel(const uint32_t* __restrict__ draws, const uint8_t* __restrict__ accept, int per_film, float ap,
                float4* __restrict__ boxes) {
    const int f = blockIdx.x;
    const unsigned lane = threadIdx.x;
    const uint2* d = reinterpret_cast<const uint2*>(draws) + (size_t)f * per_film;
    const uint8_t* acc = accept + (size_t)f * per_film;
    float minx = 0.0f, miny = 0.0f, maxx = 0.0f, maxy = 0.0f;   // warp-uniform running box
    const float inf = __int_as_float(0x7f800000);
    for (int base = 0; base < per_film; base += 32) {
        const int s = base + (int)lane;
        bool a = false;
        float px = 0.0f, py = 0.0f;
        if (s < per_film) {
            a = acc[s] != 0;
            const uint2 k = d[s];
            px = xmul(xsub(xmul(u32_to_unit(k.x), 2.0f), 1.0f), ap);
            py = xmul(xsub(xmul(u32_to_unit(k.y), 2.0f), 1.0f), ap);
        }
        if (!__any_sync(0xffffffffu, a)) continue;
        // inclusive prefix min / max over the accepted candidates of the chunk (identity for the others)
        float lminx = a ? px : inf, lminy = a ? py : inf, lmaxx = a ? px : -inf, lmaxy = a ? py : -inf;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float ax = __shfl_up_sync(0xffffffffu, lminx, o), ay = __shfl_up_sync(0xffffffffu, lminy, o);
            const float bx = __shfl_up_sync(0xffffffffu, lmaxx, o), by = __shfl_up_sync(0xffffffffu, lmaxy, o);
            // (earlier, later) -> the later value replaces the earlier one only when strictly smaller / larger, as in the
            // reference's sequential `if (p < min) min = p`
            if (lane >= (unsigned)o) {
                lminx = lminx < ax ? lminx : ax; lminy = lminy < ay ? lminy : ay;
                lmaxx = lmaxx > bx ? lmaxx : bx; lmaxy = lmaxy > by ? lmaxy : by;
            }
        }
        // the box before this lane's candidate: running box folded with the candidates of the lower lanes
        float ex = __shfl_up_sync(0xffffffffu, lminx, 1), ey = __shfl_up_sync(0xffffffffu, lminy, 1);
        if (lane == 0) { ex = inf; ey = inf; }
        const float before_minx = ex < minx ? ex : minx, before_miny = ey < miny ? ey : miny;
        const bool rearm = a && xadd(before_minx, before_miny) == 0.0f;
        if (__any_sync(0xffffffffu, rearm)) {
            // rare: replay the chunk in order with the reference's statement
            for (int j = 0; j < 32; ++j) {
                const bool aj = __shfl_sync(0xffffffffu, a ? 1 : 0, j) != 0;
                const float qx = __shfl_sync(0xffffffffu, px, j), qy = __shfl_sync(0xffffffffu, py, j);
                if (!aj) continue;
                if (xadd(minx, miny) == 0.0f) { minx = maxx = qx; miny = maxy = qy; }
                if (qx > maxx) maxx = qx;
                if (qy > maxy) maxy = qy;
                if (qx < minx) minx = qx;
                if (qy < miny) miny = qy;
            }
        } else {
            const float cx = __shfl_sync(0xffffffffu, lminx, 31), cy = __shfl_sync(0xffffffffu, lminy, 31);
            const float dx = __shfl_sync(0xffffffffu, lmaxx, 31), dy = __shfl_sync(0xffffffffu, lmaxy, 31);
            minx = cx < minx ? cx : minx; miny = cy < miny ? cy : miny;
            maxx = dx > maxx ? dx : maxx; maxy = dy > maxy ? dy : maxy;
        }
    }
    if (lane == 0) boxes[f] = make_float4(minx, miny, maxx, maxy);
}

// ------------------------------------------------------------------------------------------------
// draw.zoic ray paths: what the reference's -D_DRAW build writes from inside
// traceThroughLensElements (src/zoic.cpp:1121-1128, :1146-1153) for a drawn sample -- per surface whose rim test
// passed, the ray origin and the hit point (z, y); after the last surface, the hit point and the exit direction --
// for EVERY attempt of the sample, with the draw build's conventions (film point x = 0 :1859, direction x = 0
// :1877/:1925/:1944).  One thread per sample, exact arithmetic; the host formats the numbers.
// quads: per sample `cap` records of 4 floats; kinds: 0 = (o.z, o.y, hit.z, hit.y), 1 = (hit.z, hit.y, dir.z, dir.y).
// ------------------------------------------------------------------------------------------------
template <bool kImage, bool kLut>
__global__ void __launch_bounds__(128)
draw_paths_kernel(const __grid_constant__ CameraState cam, const float4* __restrict__ samples, uint32_t n,
                  const unsigned long long* __restrict__ indices, uint64_t first_index, uint64_t seed, float4* __restrict__ quads, uint8_t* __restrict__ kinds, uint32_t* __restrict__ counts,
                  uint32_t cap) {
    BokehView bk;
    if (kImage) bk = stage_bokeh(cam);
    const LensState& L = cam.lens;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    // all lanes run the (warp-uniform round count) table searches; lanes past the end work on sample n - 1 and store nothing
    const bool live = i < n;
    const float4 s = samples[live ? i : n - 1];
    const KolbSampleState k = kolb_sample_setup<kLut, true>(L, 0.0f, s.y);   // origin.x = 0 before the LUT lookup
    float4* q = quads + (size_t)(live ? i : 0) * cap;
    uint8_t* kd = kinds + (size_t)(live ? i : 0) * cap;
    uint32_t nq = 0;
    const uint32_t j = live ? i : n - 1;
    Xor128 rng = sample_stream(seed, indices ? (uint64_t)indices[j] : first_index + j);
    float u = s.z, v = s.w;
    for (int tries = 0;; ++tries) {
        float lx, ly;
        lens_sample<kImage>(bk, u, v, &lx, &ly);
        Ray r;
        r.o = vmake(k.fx, k.fy, L.origin_shift);
        r.d = kolb_aim<kLut>(L, k, lx, ly, tries > 0);
        r.d.x = 0.0f;
        int rc = kPass;
        for (int e = 0; e < L.count; ++e) {
            const Vec3 o_old = r.o;
            rc = exact_surface(L.e[e], r);
            if (rc == kBlocked) break;   // missed the sphere or the rim: nothing is written for this surface
            if (live && nq < cap) { q[nq] = make_float4(o_old.z, o_old.y, r.o.z, r.o.y); kd[nq] = 0; ++nq; }
            if (rc == kTir) break;
            if (e == L.count - 1 && live && nq < cap) { q[nq] = make_float4(r.o.z, r.o.y, r.d.z, r.d.y); kd[nq] = 1; ++nq; }
        }
        if (rc == kPass || tries > kMaxTries) break;
        draw_pair(rng, &u, &v);
    }
    if (live) counts[i] = nq;
}

// ------------------------------------------------------------------------------------------------
// camera -> world epilogue: the step the renderer applies to every ray after
// camera_create_ray.  origin' = M (origin, 1), dir' = M3x3 dir with M a row-major 3x4 matrix; weight and tries
// pass through.  One 32-byte record in, one out (in place allowed): HBM-bound, 64 bytes per ray.
// Arithmetic (the contract the oracle restates): each output component is one fma chain, innermost term first,
//   o'_r = fma(m[r][0], ox, fma(m[r][1], oy, fma(m[r][2], oz, m[r][3])))
//   d'_r = fma(m[r][0], dx, fma(m[r][1], dy, m[r][2] * dz))
// ------------------------------------------------------------------------------------------------
struct Xform { float m[12]; };

__device__ __forceinline__ void load_ray(const RayRecord* __restrict__ rays, uint64_t idx, float4* o, float4* d) {
    asm volatile("ld.global.cs.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(o->x), "=f"(o->y), "=f"(o->z), "=f"(o->w), "=f"(d->x), "=f"(d->y), "=f"(d->z), "=f"(d->w)
                 : "l"(rays + idx));
}

__global__ void __launch_bounds__(256)
transform_rays_kernel(const __grid_constant__ Xform X, const RayRecord* __restrict__ in, uint64_t n, RayRecord* __restrict__ out) {
    const float* m = X.m;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float4 o, d;
        load_ray(in, i, &o, &d);
        float4 oo, dd;
        oo.x = __fmaf_rn(m[0], o.x, __fmaf_rn(m[1], o.y, __fmaf_rn(m[2], o.z, m[3])));
        oo.y = __fmaf_rn(m[4], o.x, __fmaf_rn(m[5], o.y, __fmaf_rn(m[6], o.z, m[7])));
        oo.z = __fmaf_rn(m[8], o.x, __fmaf_rn(m[9], o.y, __fmaf_rn(m[10], o.z, m[11])));
        oo.w = o.w;
        dd.x = __fmaf_rn(m[0], d.x, __fmaf_rn(m[1], d.y, __fmul_rn(m[2], d.z)));
        dd.y = __fmaf_rn(m[4], d.x, __fmaf_rn(m[5], d.y, __fmul_rn(m[6], d.z)));
        dd.z = __fmaf_rn(m[8], d.x, __fmaf_rn(m[9], d.y, __fmul_rn(m[10], d.z)));
        dd.w = d.w;
        store_ray(out, i, oo, dd);
    }
}

// ------------------------------------------------------------------------------------------------
// records -> planes for the host link (include/zoicb.h: zoicb_ray_planes): the 32-byte record carries two floats that
// take two values (weight) and twenty-eight (tries); on the wire they are one byte
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pack_planar_kernel(const RayRecord* __restrict__ rays, uint64_t n, uint64_t stride, float* __restrict__ planes, uint8_t* __restrict__ flags) {
    const uint64_t step = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        float4 o, d;
        load_ray(rays, i, &o, &d);
        planes[i] = o.x; planes[stride + i] = o.y; planes[2 * stride + i] = o.z;
        planes[3 * stride + i] = d.x; planes[4 * stride + i] = d.y; planes[5 * stride + i] = d.z;
        flags[i] = (uint8_t)(((unsigned)(int)d.w & 0x7Fu) | (o.w == 0.0f ? 0x80u : 0u));
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
static unsigned grid_for(uint64_t n, int threads, int ctas_per_sm) {
    uint64_t want = (n + threads - 1) / threads;
    uint64_t cap = (uint64_t)sm_count() * ctas_per_sm;  // whole waves of resident CTAs, grid-stride beyond
    return (unsigned)(want < cap ? (want ? want : 1) : cap);
}

template <int kModel, bool kImage, bool kLut>
static cudaError_t launch_variant(const CameraState& cam, int mode, const float4* samples, uint64_t n, uint64_t first_index,
                                  uint64_t seed, RayRecord* rays, DeviceStats* stats, cudaStream_t st,
                                  const Workspace& ws, size_t smem, int stage, int* launches) {
    const int threads = 256;
    if (mode == 1 && kModel == 1) {  // guarded fast path + exact re-run of the undecided samples
        cudaError_t e = cudaMemsetAsync(ws.counters, 0, 4 * sizeof(unsigned long long), st);
        if (e != cudaSuccess) return e;
        // Flavours of the pool kernel (kolb_pool2.cu: two rays per lane, FFMA2), chosen from the camera's calibration
        // (host_setup.cpp: choose_split):
        //   plain               -- most attempts survive the first surfaces (e.g. the double Gauss);
        //   rim pre-test loop   -- most attempts die on the first surfaces (narrow-field lenses on a wide sensor:
        //                          15-22 attempts per ray; the fisheye).
        // ZOICB_POOL=2 / 3 forces plain / pre-test (A/B runs and the parity tests of both flavours on every lens).
        static const int force_pool = [] { const char* v = getenv("ZOICB_POOL"); return v ? atoi(v) : 0; }();
        CameraState c2 = cam;
        if (force_pool == 3) c2.lens.pretest = 1;
        if (force_pool == 2) c2.lens.pretest = 0;
        e = launch_kolb_pool2(c2, samples, n, first_index, seed, rays, stats, st, ws, smem, launches);
        if (e != cudaSuccess) return e;
        kolb_exact_persistent_kernel<kImage, kLut, true><<<(unsigned)sm_count() * 3, threads, smem, st>>>(
            cam, samples, 0, first_index, seed, rays, stats, ws.counters + 2, ws.queue, ws.counters + 1,
            ws.capacity);
        if (launches) *launches += 1;
        return cudaGetLastError();
    }
    if (kModel == 1 && ws.counters && n >= 65536) {  // EXACT mode on a big batch: same arithmetic, persistent schedule
        cudaError_t e = cudaMemsetAsync(ws.counters, 0, 4 * sizeof(unsigned long long), st);
        if (e != cudaSuccess) return e;
        kolb_exact_persistent_kernel<kImage, kLut, false><<<(unsigned)sm_count() * 3, threads, smem, st>>>(
            cam, samples, n, first_index, seed, rays, stats, ws.counters + 2, nullptr, nullptr, 0);
        if (launches) *launches += 1;
        return cudaGetLastError();
    }
    if (mode == 1 && kModel == 0 && cam.thin.use_dof && cam.thin.use_ov) {  // retry loop present: persistent schedule
        // (a slot-pool version of this kernel -- set-up pass / attempt pass over a shared-memory pool, like the raytraced
        // kernels -- was tried and dropped: the two table searches, not the divergent regeneration, are what an attempt
        // costs, and the pool's shared-memory traffic made it slower, 12.4-15.0 against 17.6 Grays/s; profiles/r01b_ab.txt)
        cudaError_t e = cudaMemsetAsync(ws.counters, 0, 4 * sizeof(unsigned long long), st);
        if (e != cudaSuccess) return e;
        // the column tables of an image-shaped aperture live in L1 / L2: ask for the smallest shared-memory carve-out that
        // still holds the resident CTAs' row tables, so that L1 gets the rest of the 256 KB
        static const int carve = [] { const char* v = getenv("ZOICB_THIN_CARVEOUT"); return v ? atoi(v) : -1; }();
        // byte-wide column tables when the camera has them (ZOICB_THIN_COMPACT=0 turns them off: A/B)
        static const bool allow_compact = [] { const char* v = getenv("ZOICB_THIN_COMPACT"); return !v || atoi(v) != 0; }();
        const bool compact = kImage && allow_compact && cam.compact.col_guide8 && cam.compact.rel_column8;
        const size_t smem_k = compact ? compact_smem_bytes(cam.bokeh.h) : smem;
        const int ctas = compact ? ZOICB_THIN_CTAS_COMPACT : ZOICB_THIN_CTAS;
        const size_t need = (size_t)ctas * (smem_k + (compact ? 1100 : 2200));   // + the 1 KB the system reserves per CTA and the counters
        int pct = (int)((need * 100 + 233471) / 233472);
        if (carve >= 0) pct = carve;
        if (pct > 100) pct = 100;
        // (prepared blocks -- 32 samples made ready by one dense pass and parked in registers (call 16) or in L2 scratch
        // (call 21) for finished lanes to adopt -- save 6-12 % of the warp instructions, stay bit-exact and lose to the
        // registers / the L2 round trip they cost: profiles/r02_ab.txt)
        const unsigned grid = (unsigned)sm_count() * (unsigned)ctas;
        if constexpr (kImage) {
            if (compact) {
                cudaFuncSetAttribute(thin_persistent_kernel<true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
                thin_persistent_kernel<true, true><<<grid, threads, smem_k, st>>>(cam, samples, n, first_index, seed, rays, stats, stage, ws.counters);
                if (launches) *launches += 1;
                return cudaGetLastError();
            }
            cudaFuncSetAttribute(thin_persistent_kernel<true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        }
        thin_persistent_kernel<kImage, false><<<grid, threads, smem_k, st>>>(cam, samples, n, first_index, seed, rays, stats, stage, ws.counters);
        if (launches) *launches += 1;
        return cudaGetLastError();
    }
    // grid = the CTAs that are resident at once (register-limited: 6 of 256 threads for the thin lens), so the grid-stride
    // loop runs as ONE wave; with a fixed 8 per SM the last 2 of every 8 CTAs ran as a second, mostly empty wave
    static const int resident = [] {
        int b = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, exact_kernel<kModel, kImage, kLut>, 256, 0) != cudaSuccess || b <= 0) {
            cudaGetLastError();
            b = 4;
        }
        return b;
    }();
    int ctas = resident;
    if (smem > 0) {   // row tables in dynamic shared memory can lower the residency: ask again for this size
        int b = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, exact_kernel<kModel, kImage, kLut>, threads, smem) == cudaSuccess && b > 0) ctas = b;
        else cudaGetLastError();
    }
    exact_kernel<kModel, kImage, kLut><<<grid_for(n, threads, ctas), threads, smem, st>>>(cam, samples, n, first_index, seed,
                                                                                        rays, stats, stage);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_generate(const CameraState& cam, int mode, const float4* samples, uint64_t n, uint64_t first_index,
                            uint64_t seed, RayRecord* rays, DeviceStats* stats, cudaStream_t st,
                            const Workspace& ws, int* launches) {
    if (n == 0) return cudaSuccess;
    if (n < 8192) mode = 0;  // a persistent grid does not pay for a handful of samples; the exact kernel is also bit-exact
    const bool image = cam.use_image != 0;
    size_t smem = 0;
    int stage = 0;
    if (image) {  // row tables staged in shared memory; zoicb_create rejects images with more than kMaxBokehRows rows
        smem = bokeh_smem_bytes(cam.bokeh.h);
        stage = 1;
    }
#define ZL(M, I, U) launch_variant<M, I, U>(cam, mode, samples, n, first_index, seed, rays, stats, st, ws, smem, stage, launches)
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

cudaError_t launch_draw_paths(const CameraState& cam, const float4* samples, uint32_t n, const unsigned long long* indices,
                              uint64_t first_index, uint64_t seed,
                              float4* quads, uint8_t* kinds, uint32_t* counts, uint32_t cap, cudaStream_t st, int* launches) {
    if (n == 0) return cudaSuccess;
    const unsigned grid = (n + 127) / 128;
    const bool image = cam.use_image != 0, lut = cam.lens.use_lut != 0;
    const size_t rows_smem = image ? bokeh_smem_bytes(cam.bokeh.h) : 0;
#define ZD(I, U) draw_paths_kernel<I, U><<<grid, 128, rows_smem, st>>>(cam, samples, n, indices, first_index, seed, quads, kinds, counts, cap)
    if (image) { if (lut) ZD(true, true); else ZD(true, false); }
    else { if (lut) ZD(false, true); else ZD(false, false); }
#undef ZD
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_transform(const float* m3x4, const RayRecord* in, uint64_t n, RayRecord* out, cudaStream_t st, int* launches) {
    if (n == 0) return cudaSuccess;
    Xform X;
    for (int i = 0; i < 12; ++i) X.m[i] = m3x4[i];
    transform_rays_kernel<<<grid_for(n, 256, 8), 256, 0, st>>>(X, in, n, out);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_lut_trace(const LensState& L, const float* d_film_x, int n_film, int per_film, const uint32_t* d_draws,
                             uint8_t* d_accept, cudaStream_t st, int* launches) {
    lut_trace_kernel<<<grid_for((uint64_t)n_film * per_film, 256, 8), 256, 0, st>>>(L, d_film_x, n_film, per_film, d_draws, d_accept);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_lut_bbox(const uint32_t* d_draws, const uint8_t* d_accept, int n_film, int per_film, float ap,
                            float4* d_boxes, cudaStream_t st, int* launches) {
    if (n_film <= 0) return cudaSuccess;
    lut_bbox_kernel<<<n_film, 32, 0, st>>>(d_draws, d_accept, per_film, ap, d_boxes);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_pack_planar(const RayRecord* rays, uint64_t n, uint64_t stride, uint8_t* planes, cudaStream_t st, int* launches) {
    if (n == 0) return cudaSuccess;
    pack_planar_kernel<<<grid_for(n, 256, 8), 256, 0, st>>>(rays, n, stride, reinterpret_cast<float*>(planes), planes + 24 * stride);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

// every float bit pattern: normalize_factor against the library sequence it replaces
__global__ void __launch_bounds__(256) check_normalize_factor_kernel(unsigned long long* mismatches, unsigned* first_bad) {
    unsigned long long bad = 0;
    for (unsigned long long b = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b < (1ull << 32);
         b += (unsigned long long)gridDim.x * blockDim.x) {
        const float x = __uint_as_float((unsigned)b);
        float len = xsqrt(x);
        if (len != 0) len = xrcp(len);
        const float got = normalize_factor(x);
        if (__float_as_uint(got) != __float_as_uint(len) && !(got != got && len != len)) {
            if (!bad) atomicMin(first_bad, (unsigned)b);
            ++bad;
        }
    }
    if (bad) atomicAdd(mismatches, bad);
}
cudaError_t check_normalize_factor(unsigned long long* mismatches, unsigned* first_bad, int* launches) {
    unsigned long long* d = nullptr;
    cudaError_t e = cudaMalloc(&d, 16);
    if (e != cudaSuccess) return e;
    const unsigned long long init[2] = {0ull, 0xFFFFFFFFull};
    e = cudaMemcpy(d, init, 16, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        check_normalize_factor_kernel<<<sm_count() * 8, 256>>>(d, reinterpret_cast<unsigned*>(d + 1));
        if (launches) *launches += 1;
        e = cudaGetLastError();
    }
    unsigned long long out[2] = {0, 0};
    if (e == cudaSuccess) e = cudaMemcpy(out, d, 16, cudaMemcpyDeviceToHost);
    cudaFree(d);
    *mismatches = out[0];
    *first_bad = (unsigned)out[1];
    return e;
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
