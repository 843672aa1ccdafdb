// kolb_pool.cu -- the GUARDED kernel of the raytraced lens.
#include "kernel_common.cuh"

namespace zoicb {

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
#ifndef ZOICB_POOL_WARPS
#define ZOICB_POOL_WARPS 7
#endif
constexpr int kWarpsPerCta = ZOICB_POOL_WARPS;
#ifndef ZOICB_POOL_CTAS
#define ZOICB_POOL_CTAS 4
#endif

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

template <int kN, bool kImage, bool kLut, bool kInner>
__global__ void __launch_bounds__(kWarpsPerCta * 32, ZOICB_POOL_CTAS)
kolb_pool_kernel(const __grid_constant__ CameraState cam, const float4* __restrict__ samples, uint32_t n,
                 uint64_t first_index, uint64_t seed, RayRecord* __restrict__ rays,
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
            store_ray(rays, idx, make_float4(-ox, -oy, -oz, weight), make_float4(-ux, -uy, -uz, (float)tries));
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
                    store_ray(rays, idx, o4, d4);
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
            __syncwarp();   // pops are complete before this pass pushes onto the same stack positions
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
            __syncwarp();   // pops are complete before this pass pushes onto the same stack positions
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
                if (!kInner || __popc(__ballot_sync(0xffffffffu, todo)) < 16) break;
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
// launcher
// ------------------------------------------------------------------------------------------------
template <bool kImage, bool kLut>
static cudaError_t launch_pool_variant(const CameraState& cam, const float4* samples, uint64_t n, uint64_t first_index,
                                       uint64_t seed, RayRecord* rays, DeviceStats* stats, cudaStream_t st,
                                       const Workspace& ws, size_t rows_smem, int* launches) {
    const unsigned grid = (unsigned)sm_count() * ZOICB_POOL_CTAS;  // persistent: ZOICB_POOL_CTAS CTAs of 8 warps per SM
    const int threads = kWarpsPerCta * 32;
    const size_t pool_smem = ((rows_smem + 15) & ~(size_t)15) + kWarpsPerCta * sizeof(WarpPool);
    // 32-bit sample offsets inside a launch, so very large batches go in slices
    const uint64_t slice = 1ull << 31;
    for (uint64_t b = 0; b < n; b += slice) {
        const uint32_t m = (uint32_t)((n - b < slice) ? n - b : slice);
        if (b) {
            cudaError_t e = cudaMemsetAsync(ws.counters, 0, sizeof(unsigned long long), st);  // chunk cursor only
            if (e != cudaSuccess) return e;
        }
#define ZP(N, INNER)                                                                                                      \
    do {                                                                                                                 \
        cudaFuncSetAttribute(kolb_pool_kernel<N, kImage, kLut, INNER>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                             (int)pool_smem);                                                                            \
        kolb_pool_kernel<N, kImage, kLut, INNER><<<grid, threads, pool_smem, st>>>(                                      \
            cam, samples + b, m, first_index + b, seed, rays + b, stats, ws.counters, ws.queue,       \
            ws.counters + 1, ws.capacity, b);                                                                            \
    } while (0)
#define ZPN(N) do { if (cam.lens.inner_retry) ZP(N, true); else ZP(N, false); } while (0)
        switch (cam.lens.count) {  // unrolled instantiations for the element counts of the shipped lens tables
            case 7: ZPN(7); break;
            case 8: ZPN(8); break;
            case 9: ZPN(9); break;
            case 11: ZPN(11); break;
            case 12: ZPN(12); break;
            default: ZPN(0); break;
        }
#undef ZPN
#undef ZP
        if (launches) *launches += 1;
    }
    return cudaGetLastError();
}

cudaError_t launch_kolb_pool(const CameraState& cam, const float4* samples, uint64_t n, uint64_t first_index, uint64_t seed,
                             RayRecord* rays, DeviceStats* stats, cudaStream_t st, const Workspace& ws, size_t rows_smem,
                             int* launches) {
    const bool image = cam.use_image != 0, lut = cam.lens.use_lut != 0;
#define ZL(I, U) launch_pool_variant<I, U>(cam, samples, n, first_index, seed, rays, stats, st, ws, rows_smem, launches)
    if (image) return lut ? ZL(true, true) : ZL(true, false);
    return lut ? ZL(false, true) : ZL(false, false);
#undef ZL
}

}  // namespace zoicb
