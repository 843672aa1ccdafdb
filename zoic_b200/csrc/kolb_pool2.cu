// kolb_pool2.cu -- the GUARDED kernel of the raytraced lens, packed-fp32 edition.
//
// Schedule: per-warp slot pool in shared memory, two-stage march, ballot/popc stacks (the scalar one-ray-per-lane
// kernel of the same design was retired in round 2; its A/B numbers are in profiles/r01_ab_pool2.txt).  Here
// every lane carries TWO rays through each pass and all fp32 arithmetic is issued as sm_100 packed instructions
// (fma/add/mul.rn.f32x2 -> SASS FFMA2/FADD2/FMUL2): one issue slot does the work for both rays, which frees
// issue slots for the MUFU / compare / select / integer instructions the march also needs (the scalar kernel is
// issue-bound: tools/ubench_ffma2.cu measures FFMA2 + an equal number of ALU instructions at 95 % of the FFMA
// peak, against 50 % for scalar FFMA + ALU).  The arithmetic per ray is the scalar kernel's, operation for
// operation (fma.rn.f32x2 rounds each half like fma.rn.f32), so the decision margins carry over unchanged.
#include <stdlib.h>

#include "kernel_common.cuh"

namespace zoicb {

// CTA shape and pool size per kernel flavour (A/B on the GPU, profiles/r01_ab_pool2.txt).  Slots per warp: a pass
// takes 64 slots from stack A or B, or 32 free slots for new samples; with 160 slots (nA <= 63 and nB <= 63 leave
// at least 34 free) one of the three is always possible (pigeonhole); 128 slots buy more resident warps per SM
// (7 CTAs x 3 warps) at the price of an occasional partial pass.  The in-pass re-sampling flavour needs more
// registers (5 CTAs x 3 warps, 160 slots).
#ifndef ZOICB_POOL2_WARPS
#define ZOICB_POOL2_WARPS 3
#endif
#ifndef ZOICB_POOL2_CTAS
#define ZOICB_POOL2_CTAS 7
#endif
#ifndef ZOICB_POOL2_SLOTS
#define ZOICB_POOL2_SLOTS 128
#endif
#ifndef ZOICB_POOL2_CTAS_INNER
#define ZOICB_POOL2_CTAS_INNER 5
#endif
#ifndef ZOICB_POOL2_SLOTS_INNER
#define ZOICB_POOL2_SLOTS_INNER 160
#endif
#ifndef ZOICB_POOL2_ROLLED
#define ZOICB_POOL2_ROLLED 4
#endif
// where the next set-up pass's samples are requested: 0 = at the end of a set-up pass, 1 = at its start (behind the
// arrival of this pass's samples), 2 = as 0 plus an L2 prefetch two passes ahead
#ifndef ZOICB_POOL2_PREFETCH
#define ZOICB_POOL2_PREFETCH 1   // +2.8 % (headline) / +3.5 % (fisheye) over 0, profiles/r01b_ab.txt
#endif
constexpr int kRollUnroll = ZOICB_POOL2_ROLLED > 0 ? ZOICB_POOL2_ROLLED : 1;
constexpr int kWarps2 = ZOICB_POOL2_WARPS;
// the rim pre-test loop goes round again while at least this many of the 64 rays of the pass still owe an attempt
// (lower: fewer trips through the pool; higher: fewer idle lanes in the loop)
#ifndef ZOICB_POOL2_PRE_KEEP
#define ZOICB_POOL2_PRE_KEEP 48
#endif
constexpr int kPreKeepGoing = ZOICB_POOL2_PRE_KEEP;
// kRoomy: the flavours that need more registers than 7 CTAs x 3 warps leave (97): in-pass re-sampling, image-based
// aperture sampling (two table searches), run-time stage boundary
template <bool kRoomy> struct PoolShape {
    static constexpr int kSlots = kRoomy ? ZOICB_POOL2_SLOTS_INNER : ZOICB_POOL2_SLOTS;
    static constexpr int kCtas = kRoomy ? ZOICB_POOL2_CTAS_INNER : ZOICB_POOL2_CTAS;
    static_assert(kSlots % 32 == 0 && kSlots >= 96 && kSlots <= 256, "slot ids are bytes; whole warps of slots");
};

template <int kSlots2>
struct alignas(16) WarpPool2 {   // 72 bytes per slot
    float4 film[kSlots2];   // fx, fy, max_scale, translation
    float4 misc[kSlots2];   // sn, cs, sample index (bits), packed counters (bits)
    uint4 rng[kSlots2];     // per-sample xorshift128 state
    float4 ray0[kSlots2];   // stage A -> B: ox, oy, oz, ux
    float2 tail[kSlots2];   // stage A -> B: uy, uz; a fresh sample: its first lens point (ua, ub)
    unsigned char qa[kSlots2], qb[kSlots2], qf[kSlots2];
};
// packed counters: tries [0..7] | fresh [8] | tir [9..15] | surface visits [16..31]
__device__ __forceinline__ unsigned pk2_tries(unsigned p) { return p & 0xffu; }
__device__ __forceinline__ bool pk2_fresh(unsigned p) { return (p >> 8) & 1u; }
__device__ __forceinline__ unsigned pk2_tir(unsigned p) { return (p >> 9) & 0x7fu; }
__device__ __forceinline__ unsigned pk2_visits(unsigned p) { return p >> 16; }

// ---- packed fp32 helpers: low half = ray 0 of the lane, high half = ray 1.  The carrier is a 64-bit register
// (an aligned register pair by construction), not a float2: with float2 the compiler keeps the halves in
// unrelated registers and re-packs them with MOVs in front of every packed instruction.
typedef unsigned long long f2;
__device__ __forceinline__ f2 mk(float a, float b) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ f2 bc(float a) { return mk(a, a); }
__device__ __forceinline__ float lo(f2 p) { float a; asm("{ .reg .f32 t; mov.b64 {%0, t}, %1; }" : "=f"(a) : "l"(p)); return a; }
__device__ __forceinline__ float hi(f2 p) { float a; asm("{ .reg .f32 t; mov.b64 {t, %0}, %1; }" : "=f"(a) : "l"(p)); return a; }
__device__ __forceinline__ f2 neg2(f2 p) {
    f2 r;
    asm("{ .reg .f32 a, b; mov.b64 {a, b}, %1; neg.f32 a, a; neg.f32 b, b; mov.b64 %0, {a, b}; }" : "=l"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ f2 abs2(f2 p) {
    f2 r;
    asm("{ .reg .f32 a, b; mov.b64 {a, b}, %1; abs.f32 a, a; abs.f32 b, b; mov.b64 %0, {a, b}; }" : "=l"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float half_of(f2 v, int h) { return h ? hi(v) : lo(v); }

struct RayPair { f2 ox, oy, oz, ux, uy, uz; };

// First half of the surface step for the two rays of a lane: intersection with the sphere the reference
// intersects, hit point (returned in place of the origin), and the quantities of the rim / miss decision.
__device__ __forceinline__ void surface_hit(const float4 q0, const float4 q1, const float4 q2, const float4 q3, f2& px, f2& py,
                                            f2& pz, const f2 vx, const f2 vy, const f2 vz, f2* w, f2* margin, f2* guard,
                                            f2* disc_out) {
    const float e_center = q0.x, e_sgn = q0.w, e_rim2 = q1.x;
    const float e_rim2_guard = q2.y, e_dt_guard = q2.z, e_vertex = q2.w, e_r2_corr = q3.x, e_vertex_m2r = q3.z;
    const f2 dz = sub2(bc(e_vertex), pz);
    const f2 m2 = sub2(bc(e_vertex_m2r), pz);                            // dz - 2R
    const f2 Lz = sub2(bc(e_center), pz);
    const f2 tca = fma2(Lz, vz, neg2(fma2(px, vx, mul2(py, vy))));
    // C = |o - c|^2 - radius2 = dz (dz - 2R) + ox^2 + oy^2 + (R^2 - fl(R^2))
    const f2 C = fma2(dz, m2, fma2(px, px, fma2(py, py, bc(e_r2_corr))));
    const f2 disc = fma2(tca, tca, neg2(C));
    const f2 s = mul2(bc(e_sgn), mk(approx_sqrt(fmaxf(lo(disc), 0.0f)), approx_sqrt(fmaxf(hi(disc), 0.0f))));
    const f2 ts = mul2(tca, s);
    const f2 den = sub2(tca, s);
    const f2 sum = add2(tca, s);
    const f2 conj = mul2(C, mk(approx_rcp(lo(den)), approx_rcp(hi(den))));   // conjugate root when tca + s cancels
    const f2 t = mk(lo(ts) < 0.0f ? lo(conj) : lo(sum), hi(ts) < 0.0f ? hi(conj) : hi(sum));
    px = fma2(vx, t, px); py = fma2(vy, t, py); pz = fma2(vz, t, pz);   // the hit point is the next origin
    *w = fma2(px, vx, mul2(py, vy));
    *margin = fma2(px, px, fma2(py, py, bc(-e_rim2)));
    *guard = fma2(abs2(*w), bc(e_dt_guard), bc(e_rim2_guard));
    *disc_out = disc;
}

// Surfaces [from, to) for the two rays of a lane; from/to are warp-uniform.  A ray that is stopped (or was not
// alive on entry) gets NaN state: every ordered comparison of the decision tests is then false for it, so it
// marches on next to its live partner without any per-surface bookkeeping.  No lane leaves the loop early (the
// warp would execute the surface for its other lanes anyway), so the surface index stays warp-uniform and the
// element constants come through the uniform datapath.  rc/visited are only meaningful for rays that entered
// alive; a stopped ray's state is dead.
// kEarlyExit: the WARP leaves the loop once every ray in it is stopped (a vote per surface; worth it where most
// attempts die on the first surfaces of a long stage).  `elems` is the element table in shared memory.
template <int kN, bool kEarlyExit>
__device__ __forceinline__ void march_pair(const float4* __restrict__ elems, float gscale, int from, int to, RayPair& r,
                                           bool a0, bool a1, int* rc0, int* rc1, int* visited0, int* visited1) {
    const float tir_band = 1e-4f * gscale;
    const float qnan = __int_as_float(0x7fc00000);
    int last0 = to - 1, last1 = to - 1, c0 = kPass, c1r = kPass;
    int dead = (a0 ? 0 : 1) | (a1 ? 0 : 2);
    f2 px = mk(a0 ? lo(r.ox) : qnan, a1 ? hi(r.ox) : qnan), py = r.oy, pz = r.oz, vx = r.ux, vy = r.uy, vz = r.uz;
#if ZOICB_POOL2_ROLLED
    // rolled: one copy of the surface code (stays in the instruction cache)
#pragma unroll kRollUnroll
    for (int i = from; i < to; ++i) {
#else
#pragma unroll
    for (int i = 0; i < (kN > 0 ? kN : kMaxElements); ++i) {
        if (i < from) continue;   // warp-uniform
        if (i >= to) break;       // warp-uniform
#endif
        // the 16 constants of surface i: four 128-bit shared-memory loads, one address for the whole warp
        const float4* ep = elems + 4 * i;
        const float4 q0 = ep[0], q1 = ep[1], q2 = ep[2], q3 = ep[3];
        const float e_center = q0.x, e_eta = q1.y, e_eta2 = q1.z, e_inv_radius = q1.w, e_miss_guard = q3.y, e_one_m_eta2 = q3.w;
        f2 w, margin, guard, disc;
        surface_hit(q0, q1, q2, q3, px, py, pz, vx, vy, vz, &w, &margin, &guard, &disc);   // (px, py, pz) becomes the hit point
        const f2 nzr = sub2(bc(e_center), pz);
        const f2 c1 = mul2(fma2(neg2(vz), nzr, w), bc(e_inv_radius));
        const f2 rad = fma2(mul2(bc(e_eta2), c1), c1, bc(e_one_m_eta2));   // 1 - cs2; negative => TIR
        f2 kk = mul2(fma2(bc(e_eta), c1, neg2(mk(approx_sqrt(lo(rad)), approx_sqrt(hi(rad))))), bc(e_inv_radius));
        // One test per surface for both rays and both ways of stopping (rim / miss first, then total internal
        // reflection); everything a stopped ray computed past its stopping point is discarded here.
        const bool s0 = lo(margin) > -lo(guard) || lo(disc) < e_miss_guard;   // stopped at the rim, or too close to call
        const bool s1 = hi(margin) > -hi(guard) || hi(disc) < e_miss_guard;
        const bool t0 = lo(rad) < tir_band, t1 = hi(rad) < tir_band;
        if (s0 || s1 || t0 || t1) {   // rare past the first surfaces: one divergent region, selects inside
            const int k0 = ((lo(disc) < -e_miss_guard) || (lo(margin) > lo(guard))) ? kBlocked : kUndecided;
            const int k1 = ((hi(disc) < -e_miss_guard) || (hi(margin) > hi(guard))) ? kBlocked : kUndecided;
            const int j0 = lo(rad) < -tir_band ? kTir : kUndecided;
            const int j1 = hi(rad) < -tir_band ? kTir : kUndecided;
            const bool z0 = s0 || t0, z1 = s1 || t1;
            c0 = s0 ? k0 : (t0 ? j0 : c0); last0 = z0 ? i : last0;
            c1r = s1 ? k1 : (t1 ? j1 : c1r); last1 = z1 ? i : last1;
            kk = mk(z0 ? qnan : lo(kk), z1 ? qnan : hi(kk));   // poisons the new direction, and with it everything after
            if (kEarlyExit) dead |= (z0 ? 1 : 0) | (z1 ? 2 : 0);
        }
        if (kEarlyExit && __all_sync(0xffffffffu, dead == 3)) break;
        vx = fma2(kk, neg2(px), mul2(bc(e_eta), vx));
        vy = fma2(kk, neg2(py), mul2(bc(e_eta), vy));
        vz = fma2(kk, nzr, mul2(bc(e_eta), vz));
    }
    r.ox = px; r.oy = py; r.oz = pz; r.ux = vx; r.uy = vy; r.uz = vz;
    *rc0 = c0; *rc1 = c1r;
    *visited0 = last0 - from + 1;
    *visited1 = last1 - from + 1;
}

// fastSin parabola for both rays (same unfused operations as parabola_sin, lens_math.cuh)
__device__ __forceinline__ f2 parabola_sin2(f2 x) {
    const float B = 4.0f / ZOICB_PI_F;
    const float C = -4.0f / (ZOICB_PI_F * ZOICB_PI_F);
    const f2 y = add2(mul2(bc(B), x), mul2(mul2(bc(C), x), abs2(x)));
    return add2(mul2(bc(0.225f), sub2(mul2(y, abs2(y)), y)), y);
}

// concentric_disk_fast (kernel_common.cuh) for both rays
__device__ __forceinline__ void concentric_disk_fast2(f2 u, f2 v, f2* lx, f2* ly) {
    const f2 a = fma2(bc(2.0f), u, bc(-1.0f));
    const f2 b = fma2(bc(2.0f), v, bc(-1.0f));
    const f2 aa = mul2(a, a), bb = mul2(b, b);
    const bool f0 = lo(aa) > lo(bb), f1 = hi(aa) > hi(bb);
    const f2 num = mk(f0 ? lo(b) : lo(a), f1 ? hi(b) : hi(a)), den = mk(f0 ? lo(a) : lo(b), f1 ? hi(a) : hi(b));
    const f2 qt = mul2(num, mk(approx_rcp(lo(den)), approx_rcp(hi(den))));
    const f2 p1 = mul2(bc(0.78539816339f), qt);
    const f2 p2 = fma2(bc(-0.78539816339f), qt, bc(1.57079632679489661923f));
    const f2 phi = mk(f0 ? lo(p1) : lo(p2), f1 ? hi(p1) : hi(p2));
    // phi in [-pi/4, 3pi/4]: phi + pi < 2pi always; (phi + pi/2) + pi may pass 2pi once
    const float two_pi = ZOICB_PI_F * 2.0f;
    const f2 xs = sub2(add2(phi, bc(ZOICB_PI_F)), bc(ZOICB_PI_F));
    const f2 vc = add2(add2(phi, bc(ZOICB_PI_F * 0.5f)), bc(ZOICB_PI_F));
    const f2 vw = sub2(vc, bc(two_pi));
    const f2 xc = sub2(mk(lo(vc) >= two_pi ? lo(vw) : lo(vc), hi(vc) >= two_pi ? hi(vw) : hi(vc)), bc(ZOICB_PI_F));
    *lx = mul2(den, parabola_sin2(xc));
    *ly = mul2(den, parabola_sin2(xs));
}

// kSplit >= 0: the stage boundary is a compile-time constant (both stages become straight-line code without
// per-surface entry/exit tests); kSplit < 0: taken from the camera state at run time.
// kPre: the flavour for cameras whose attempts overwhelmingly die at the FIRST surface (rear rim): stage A is a tight
// re-sampling loop of (draw, aperture sample, aim, surface-0 intersection + rim test) without refraction; a ray that
// clears the rim leaves the loop and stage B marches it through the whole stack, surface 0 included.
template <int kN, int kSplit, bool kImage, bool kLut, bool kInner, bool kPre = false>
__global__ void __launch_bounds__(kWarps2 * 32, PoolShape<(kInner || kImage || kSplit < 0)>::kCtas)
kolb_pool2_kernel(const __grid_constant__ CameraState cam, const float4* __restrict__ samples, uint32_t n,
                  uint64_t first_index, uint64_t seed, RayRecord* __restrict__ rays,
                  DeviceStats* stats, unsigned long long* chunk_counter, QueueRecord* queue,
                  unsigned long long* queue_count, unsigned long long capacity, uint64_t queue_base) {
    // dynamic shared memory: [bokeh row tables (2h floats, 16-byte aligned)] [element table] [one WarpPool2 per warp]
    BokehView bk;
    if (kImage) bk = stage_bokeh(cam);
    const unsigned rows_bytes = kImage ? bokeh_smem_bytes(cam.bokeh.h) : 0u;
    const LensState& L = cam.lens;
    float4* elems = reinterpret_cast<float4*>(reinterpret_cast<char*>(s_rows) + rows_bytes);
    for (int i = threadIdx.x; i < 4 * kMaxElements; i += blockDim.x) elems[i] = reinterpret_cast<const float4*>(L.e)[i];
    __syncthreads();
    constexpr int kSlots2 = PoolShape<(kInner || kImage || kSplit < 0)>::kSlots;
    WarpPool2<kSlots2>& P = reinterpret_cast<WarpPool2<kSlots2>*>(elems + 4 * kMaxElements)[threadIdx.x >> 5];
    const int count = kN > 0 ? kN : L.count;
    const int split = kPre ? 0 : (kSplit >= 0 ? kSplit : L.split);   // stage B starts at surface `split`
    const unsigned lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    LocalStats ls = {0, 0, 0, 0, 0, 0, 0};
    for (int i = lane; i < kSlots2; i += 32) P.qf[i] = (unsigned char)i;
    __syncwarp();
    int nA = 0, nB = 0, nF = kSlots2;   // warp-uniform stack heights
    uint32_t cur = 0, end = 0;
    bool exhausted = false;
    float4 pre = make_float4(0, 0, 0, 0);   // samples[cur + lane], loaded one set-up pass ahead
    bool pre_ok = false;

    // push the slots of both rays of every lane (where p0 / p1 is set) onto a stack; returns the new height
    auto push2 = [&](unsigned char* stack, int height, bool p0, int slot0, bool p1, int slot1) {
        const unsigned m0 = __ballot_sync(0xffffffffu, p0);
        const unsigned m1 = __ballot_sync(0xffffffffu, p1);
        const int n0 = __popc(m0);
        if (p0) stack[height + __popc(m0 & lt_mask)] = (unsigned char)slot0;
        if (p1) stack[height + n0 + __popc(m1 & lt_mask)] = (unsigned char)slot1;
        return height + n0 + __popc(m1);
    };
    // a finished sample: counters and the output record.  A sample that ran out of retries gets weight 0 and --
    // its half-traced state being meaningless in the reference too -- the film point as
    // origin and the optical axis as direction.
    auto emit = [&](int slot, uint32_t idx, unsigned packed, float ox, float oy, float oz, float ux, float uy, float uz) {
        const unsigned tries = pk2_tries(packed);
        float weight = 1.0f;
        if (tries > (unsigned)kMaxTries) {   // the marched state of a stopped ray is dead (NaN): film point, optical axis
            const float4 f = P.film[slot];
            weight = 0.0f; ls.vignetted++;
            ox = f.x; oy = f.y; oz = L.origin_shift; ux = 0.0f; uy = 0.0f; uz = 1.0f;
        }
        else ls.success++;
        weight *= cam.weight_scale;
        store_ray(rays, idx, make_float4(-ox, -oy, -oz, weight), make_float4(-ux, -uy, -uz, (float)tries));
        ls.rays++;
        ls.attempts += tries + 1;
        ls.visits += pk2_visits(packed);
        ls.tir += pk2_tir(packed);
    };
    // undecided samples go to the exact re-run queue (one atomic per warp for both rays of every lane), with the packed
    // counters of the attempts BEFORE the undecided one (tries = its number): the exact kernel resumes there instead of
    // repeating the decided attempts.  The counters are read back from the slot (misc.w) inside the rare store, minus
    // `sub` (stage B: the attempt's stage-A visits), so nothing extra stays live across the march.
    auto enqueue2 = [&](bool u0, uint32_t idx0, int slot0, bool u1, uint32_t idx1, int slot1, unsigned sub) {
        const unsigned m0 = __ballot_sync(0xffffffffu, u0);
        const unsigned m1 = __ballot_sync(0xffffffffu, u1);
        if ((m0 | m1) == 0u) return;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(queue_count, (unsigned long long)(__popc(m0) + __popc(m1)));
        base = __shfl_sync(0xffffffffu, base, 0);
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            const bool u = h ? u1 : u0;
            if (!u) continue;
            const uint32_t idx = h ? idx1 : idx0;
            const unsigned long long pos = base + (h ? __popc(m0) + __popc(m1 & lt_mask) : __popc(m0 & lt_mask));
            if (pos < capacity) {
                const unsigned before = __float_as_uint(P.misc[h ? slot1 : slot0].w) - sub;
                queue[pos] = queue_pack(queue_base + idx, pk2_tries(before), pk2_tir(before), pk2_visits(before));
            } else {  // queue full: settle it here, exactly
                float4 o4, d4;
                kolb_exact_sample<kImage, kLut>(cam, bk, samples[idx], first_index + idx, seed, &o4, &d4, ls);
                store_ray(rays, idx, o4, d4);
                ls.reruns++;
            }
        }
    };

    for (;;) {
        // ---------------- pick the next pass: a full pass from one of the stacks whenever there is one
        const bool more = !exhausted || cur < end;
        int mode, m = 64;   // mode 0: stage B, 1: stage A, 2: take new samples
        if (nB >= 64) mode = 0;
        else if (nA >= 64) mode = 1;
        else if (more && nF >= 32) mode = 2;
        else if (nB > 0 && nB >= nA) { mode = 0; m = nB; }      // partial passes: the tail of the launch, or a small pool
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
                pre_ok = false;
            }
            int take = nF < 32 ? nF : 32;
            if (take > (int)(end - cur)) take = (int)(end - cur);
            float4 s = pre;
            if (!pre_ok && (int)lane < take) s = __ldcs(samples + cur + lane);
#if ZOICB_POOL2_PREFETCH == 1
            // The next set-up pass's samples: the load is issued HERE, right after this pass's samples have arrived (its
            // address carries a run-time zero made from s.x, so it cannot be scheduled any earlier and never shares a
            // scoreboard wait with the load above), and the ~270 instructions of this pass cover its latency.  Issued at
            // the end of the pass instead, the loop-top branch waited for it: 9 % of all warp time (ncu: stall_long_sb
            // on that one instruction).
            {
                // capacity <= 2^27 (get_workspace), so this is zero -- but only at run time
                const unsigned zero = __float_as_uint(s.x) & (unsigned)(capacity >> 32);
                pre_ok = take == 32 && cur + 64 <= end;
                if (pre_ok) pre = __ldcs(samples + cur + 32 + lane + zero);
            }
#elif ZOICB_POOL2_PREFETCH == 2
            // L2 prefetch two passes ahead (no destination register, no scoreboard)
            if (cur + 96 <= end) asm volatile("prefetch.global.L2 [%0];" ::"l"(samples + cur + 64 + lane));
#endif
            if ((int)lane < take) {
                const int slot = P.qf[nF - 1 - lane];
                const uint32_t idx = cur + lane;
                const KolbSampleState k = kolb_sample_setup<kLut, false>(L, s.x, s.y);
                const Xor128 g = sample_stream(seed, first_index + idx);
                P.film[slot] = make_float4(k.fx, k.fy, k.max_scale, k.translation);
                P.misc[slot] = make_float4(k.sn, k.cs, __uint_as_float(idx), __uint_as_float(1u << 8));  // fresh, tries 0
                P.rng[slot] = make_uint4(g.x, g.y, g.z, g.w);
                P.tail[slot] = make_float2(s.z, s.w);
                P.qa[nA + lane] = (unsigned char)slot;
            }
            nF -= take;
            nA += take;
            cur += take;
            __syncwarp();
            // the next set-up pass's samples travel while the march passes in between compute (issued last, so that
            // no scoreboard wait of this pass covers it)
#if ZOICB_POOL2_PREFETCH != 1
            pre_ok = take == 32 && cur + 32 <= end;
            if (pre_ok) pre = __ldcs(samples + cur + lane);
#endif
        } else if (mode == 0) {
            // ---------------- stage B: surfaces [split, count) for survivors of stage A
            const bool act0 = (int)lane < m, act1 = (int)lane + 32 < m;
            const int slot0 = act0 ? P.qb[nB - 1 - lane] : 0;
            const int slot1 = act1 ? P.qb[nB - 33 - lane] : 0;
            nB -= m;
            __syncwarp();   // pops are complete before this pass pushes onto the same stack positions
            float4 a0 = make_float4(0, 0, 0, 0), b0 = a0;
            float2 a1 = make_float2(0, 1), b1 = a1, ai = make_float2(0, 0), bi = ai;
            if (act0) { a0 = P.ray0[slot0]; a1 = P.tail[slot0]; ai = *reinterpret_cast<const float2*>(&P.misc[slot0].z); }
            if (act1) { b0 = P.ray0[slot1]; b1 = P.tail[slot1]; bi = *reinterpret_cast<const float2*>(&P.misc[slot1].z); }
            RayPair r = {mk(a0.x, b0.x), mk(a0.y, b0.y), mk(a0.z, b0.z), mk(a0.w, b0.w), mk(a1.x, b1.x), mk(a1.y, b1.y)};
            const uint32_t idx0 = __float_as_uint(ai.x), idx1 = __float_as_uint(bi.x);
            unsigned packed0 = __float_as_uint(ai.y), packed1 = __float_as_uint(bi.y);
            int rc0, rc1, v0, v1;
            march_pair<kN, false>(elems, cam.guard_scale, split, count, r, act0, act1, &rc0, &rc1, &v0, &v1);
            bool again[2], done[2], und[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const bool act = h ? act1 : act0;
                const int rc = h ? rc1 : rc0;
                unsigned& packed = h ? packed1 : packed0;
                if (act) {
                    packed += (unsigned)(h ? v1 : v0) << 16;
                    if (rc == kTir) packed += 1u << 9;
                }
                const bool failed = act && (rc == kBlocked || rc == kTir);
                again[h] = failed && pk2_tries(packed) <= (unsigned)kMaxTries;
                und[h] = act && rc == kUndecided;
                done[h] = act && (rc == kPass || (failed && !again[h]));
                if (again[h]) P.misc[h ? slot1 : slot0].w = __uint_as_float(packed);
                if (done[h])
                    emit(h ? slot1 : slot0, h ? idx1 : idx0, packed, half_of(r.ox, h), half_of(r.oy, h), half_of(r.oz, h), half_of(r.ux, h),
                         half_of(r.uy, h), half_of(r.uz, h));
            }
            // an undecided attempt is repeated from its start by the exact kernel: the counters before it are the ones the
            // slot still holds (stage-B entry) minus the attempt's stage-A share (all `split` surfaces)
            enqueue2(und[0], idx0, slot0, und[1], idx1, slot1, (unsigned)split << 16);
            nA = push2(P.qa, nA, again[0], slot0, again[1], slot1);
            nF = push2(P.qf, nF, done[0] || und[0], slot0, done[1] || und[1], slot1);
            __syncwarp();
        } else {
            // ---------------- stage A: lens point, aim, surfaces [0, split)
            const bool act0 = (int)lane < m, act1 = (int)lane + 32 < m;
            const int slot0 = act0 ? P.qa[nA - 1 - lane] : 0;
            const int slot1 = act1 ? P.qa[nA - 33 - lane] : 0;
            nA -= m;
            __syncwarp();   // pops are complete before this pass pushes onto the same stack positions
            float4 fa = make_float4(0, 0, 1, 0), ra = make_float4(0, 1, 0, 0), fb = fa, rb = ra;
            float2 ta = make_float2(0.5f, 0.25f), tb = ta;
            uint4 ga = make_uint4(1, 2, 3, 4), gb = ga;
            if (act0) { fa = P.film[slot0]; ra = P.misc[slot0]; ga = P.rng[slot0]; ta = P.tail[slot0]; }
            if (act1) { fb = P.film[slot1]; rb = P.misc[slot1]; gb = P.rng[slot1]; tb = P.tail[slot1]; }
            const uint32_t idx0 = __float_as_uint(ra.z), idx1 = __float_as_uint(rb.z);
            unsigned packed0 = __float_as_uint(ra.w), packed1 = __float_as_uint(rb.w);
            bool fresh0 = pk2_fresh(packed0), fresh1 = pk2_fresh(packed1);
            packed0 &= ~(1u << 8);
            packed1 &= ~(1u << 8);
            const f2 fx = mk(fa.x, fb.x), fy = mk(fa.y, fb.y), scale = mk(fa.z, fb.z), trans = mk(fa.w, fb.w);
            const f2 sn = mk(ra.x, rb.x), cs = mk(ra.y, rb.y);
            float ua0 = ta.x, ub0 = ta.y, ua1 = tb.x, ub1 = tb.y;   // the first lens point of a fresh sample
            Xor128 g0 = {ga.x, ga.y, ga.z, ga.w}, g1 = {gb.x, gb.y, gb.z, gb.w};
            RayPair r = {fx, fy, bc(L.origin_shift), bc(0.0f), bc(0.0f), bc(1.0f)};
            int rc0 = kPass, rc1 = kPass;
            int lastv0 = 0, lastv1 = 0;        // surfaces the ray's latest attempt of this pass visited
            bool todo0 = act0, todo1 = act1;   // rays that still owe an attempt in this pass
            // While at least half the rays of the pass were stopped inside stage A, those rays re-sample right here
            // instead of going round through the stacks (the cheap path for cameras whose attempts mostly die at the
            // rear rim).
            for (;;) {
                if (todo0 && !fresh0) { draw_pair(g0, &ua0, &ub0); packed0 += 1u; }   // ++tries
                if (todo1 && !fresh1) { draw_pair(g1, &ua1, &ub1); packed1 += 1u; }
                f2 lx, ly;
                if (kImage) {
                    float x0, y0, x1, y1;
                    bokeh_sample(bk, ua0, ub0, &x0, &y0);
                    bokeh_sample(bk, ua1, ub1, &x1, &y1);
                    lx = mk(x0, x1); ly = mk(y0, y1);
                } else {
                    concentric_disk_fast2(mk(ua0, ua1), mk(ub0, ub1), &lx, &ly);
                }
                // kolb_aim (lens_math.cuh): the retry arithmetic adds the translation to BOTH components (:1933 vs :1914)
                f2 dx, dy;
                if (kLut) {
                    const f2 px = add2(mul2(lx, scale), trans);
                    const f2 py0 = mul2(ly, scale);
                    const f2 py1 = add2(py0, trans);
                    const f2 py = mk(fresh0 ? lo(py0) : lo(py1), fresh1 ? hi(py0) : hi(py1));
                    const f2 rx = sub2(mul2(px, cs), mul2(py, sn));
                    const f2 ry = add2(mul2(px, sn), mul2(py, cs));
                    dx = sub2(rx, fx); dy = sub2(ry, fy);
                } else {
                    dx = sub2(mul2(lx, scale), fx); dy = sub2(mul2(ly, scale), fy);
                }
                const float dzc = L.neg_first_thickness;
                const f2 q = fma2(dx, dx, fma2(dy, dy, bc(dzc * dzc)));
                f2 y = mk(approx_rsqrt(lo(q)), approx_rsqrt(hi(q)));
                y = mul2(y, fma2(mul2(mul2(bc(-0.5f), q), y), y, bc(1.5f)));
                RayPair a = {fx, fy, bc(L.origin_shift), mul2(dx, y), mul2(dy, y), mul2(bc(dzc), y)};
                int nrc0, nrc1, v0, v1;
                if (kPre) {
                    // surface 0: intersection and rim / miss test only, the arithmetic of the march's first surface.  A ray
                    // that clears the rim goes on to stage B, which marches the whole stack from the film point (and
                    // counts the visit to surface 0); a blocked one has visited one surface.
                    const float4 q0 = elems[0], q1 = elems[1], q2 = elems[2], q3 = elems[3];
                    f2 hx = a.ox, hy = a.oy, hz = a.oz, w, margin, guard, disc;
                    surface_hit(q0, q1, q2, q3, hx, hy, hz, a.ux, a.uy, a.uz, &w, &margin, &guard, &disc);
                    const float miss = q3.y;
                    const bool s0 = lo(margin) > -lo(guard) || lo(disc) < miss, s1 = hi(margin) > -hi(guard) || hi(disc) < miss;
                    const bool b0 = lo(disc) < -miss || lo(margin) > lo(guard), b1 = hi(disc) < -miss || hi(margin) > hi(guard);
                    nrc0 = !s0 ? kPass : (b0 ? kBlocked : kUndecided);
                    nrc1 = !s1 ? kPass : (b1 ? kBlocked : kUndecided);
                    v0 = nrc0 == kBlocked ? 1 : 0;
                    v1 = nrc1 == kBlocked ? 1 : 0;
                } else {
                    march_pair<kN, kInner>(elems, cam.guard_scale, 0, split, a, todo0, todo1, &nrc0, &nrc1, &v0, &v1);
                }
                if (todo0) { rc0 = nrc0; lastv0 = v0; packed0 += (unsigned)v0 << 16; if (nrc0 == kTir) packed0 += 1u << 9; fresh0 = false; }
                if (todo1) { rc1 = nrc1; lastv1 = v1; packed1 += (unsigned)v1 << 16; if (nrc1 == kTir) packed1 += 1u << 9; fresh1 = false; }
                if (kPre) {   // the origin stays the film point; only the direction of a re-sampled ray changes
                    r.ux = mk(todo0 ? lo(a.ux) : lo(r.ux), todo1 ? hi(a.ux) : hi(r.ux));
                    r.uy = mk(todo0 ? lo(a.uy) : lo(r.uy), todo1 ? hi(a.uy) : hi(r.uy));
                    r.uz = mk(todo0 ? lo(a.uz) : lo(r.uz), todo1 ? hi(a.uz) : hi(r.uz));
                } else if (kInner) {   // rays that were not re-sampled in this round keep their state
                    r.ox = mk(todo0 ? lo(a.ox) : lo(r.ox), todo1 ? hi(a.ox) : hi(r.ox));
                    r.oy = mk(todo0 ? lo(a.oy) : lo(r.oy), todo1 ? hi(a.oy) : hi(r.oy));
                    r.oz = mk(todo0 ? lo(a.oz) : lo(r.oz), todo1 ? hi(a.oz) : hi(r.oz));
                    r.ux = mk(todo0 ? lo(a.ux) : lo(r.ux), todo1 ? hi(a.ux) : hi(r.ux));
                    r.uy = mk(todo0 ? lo(a.uy) : lo(r.uy), todo1 ? hi(a.uy) : hi(r.uy));
                    r.uz = mk(todo0 ? lo(a.uz) : lo(r.uz), todo1 ? hi(a.uz) : hi(r.uz));
                } else {
                    r = a;
                }
                todo0 = todo0 && (rc0 == kBlocked || rc0 == kTir) && pk2_tries(packed0) <= (unsigned)kMaxTries;
                todo1 = todo1 && (rc1 == kBlocked || rc1 == kTir) && pk2_tries(packed1) <= (unsigned)kMaxTries;
                if (!kInner) break;
                if (__popc(__ballot_sync(0xffffffffu, todo0)) + __popc(__ballot_sync(0xffffffffu, todo1)) < (kPre ? kPreKeepGoing : 32)) break;
            }
            bool again[2], onward[2], done[2], und[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const bool act = h ? act1 : act0;
                const int rc = h ? rc1 : rc0;
                const unsigned packed = h ? packed1 : packed0;
                const int slot = h ? slot1 : slot0;
                const uint32_t idx = h ? idx1 : idx0;
                const Xor128& g = h ? g1 : g0;
                const bool failed = act && (rc == kBlocked || rc == kTir);
                again[h] = failed && pk2_tries(packed) <= (unsigned)kMaxTries;
                onward[h] = act && rc == kPass;
                done[h] = failed && !again[h];
                und[h] = act && rc == kUndecided;
                if (again[h] || onward[h]) {
                    P.rng[slot] = make_uint4(g.x, g.y, g.z, g.w);
                    if (onward[h]) {
                        P.ray0[slot] = make_float4(half_of(r.ox, h), half_of(r.oy, h), half_of(r.oz, h), half_of(r.ux, h));
                        P.tail[slot] = make_float2(half_of(r.uy, h), half_of(r.uz, h));
                    }
                    P.misc[slot].w = __uint_as_float(packed);
                }
                // undecided: leave the counters of the attempts before this one in the slot for enqueue2
                if (und[h]) P.misc[slot].w = __uint_as_float(packed - ((unsigned)(h ? lastv1 : lastv0) << 16));
                if (done[h])
                    emit(slot, idx, packed, half_of(r.ox, h), half_of(r.oy, h), half_of(r.oz, h), half_of(r.ux, h), half_of(r.uy, h),
                         half_of(r.uz, h));
            }
            enqueue2(und[0], idx0, slot0, und[1], idx1, slot1, 0u);   // reads back the lane's own slot words written above
            nA = push2(P.qa, nA, again[0], slot0, again[1], slot1);
            nB = push2(P.qb, nB, onward[0], slot0, onward[1], slot1);
            nF = push2(P.qf, nF, done[0] || und[0], slot0, done[1] || und[1], slot1);
            __syncwarp();
        }
    }
    flush_stats(ls, stats);
}

// ------------------------------------------------------------------------------------------------
// launcher
// ------------------------------------------------------------------------------------------------
template <bool kImage, bool kLut>
static cudaError_t launch_pool2_variant(const CameraState& cam, const float4* samples, uint64_t n, uint64_t first_index,
                                        uint64_t seed, RayRecord* rays, DeviceStats* stats, cudaStream_t st,
                                        const Workspace& ws, size_t rows_smem, int* launches) {
    const int threads = kWarps2 * 32;
    const size_t fixed_smem = ((rows_smem + 15) & ~(size_t)15) + kMaxElements * sizeof(Element);
    // 32-bit sample offsets inside a launch, so very large batches go in slices
    const uint64_t slice = 1ull << 31;
    for (uint64_t b = 0; b < n; b += slice) {
        const uint32_t m = (uint32_t)((n - b < slice) ? n - b : slice);
        if (b) {
            cudaError_t e = cudaMemsetAsync(ws.counters, 0, sizeof(unsigned long long), st);  // chunk cursor only
            if (e != cudaSuccess) return e;
        }
#define ZPX(N, S, INNER, PRE)                                                                                              \
    do {                                                                                                                 \
        typedef PoolShape<((INNER) || kImage || (S) < 0)> Shape;                                                         \
        const unsigned grid = (unsigned)sm_count() * Shape::kCtas;   /* persistent */                                    \
        const size_t pool_smem = fixed_smem + kWarps2 * sizeof(WarpPool2<Shape::kSlots>);                                \
        cudaFuncSetAttribute(kolb_pool2_kernel<N, S, kImage, kLut, INNER, PRE>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                             (int)pool_smem);                                                                            \
        kolb_pool2_kernel<N, S, kImage, kLut, INNER, PRE><<<grid, threads, pool_smem, st>>>(                             \
            cam, samples + b, m, first_index + b, seed, rays + b, stats, ws.counters, ws.queue,                          \
            ws.counters + 1, ws.capacity, b);                                                                            \
    } while (0)
#define ZP(N, S, INNER) ZPX(N, S, INNER, false)
#define ZPI(N, S) do { if (cam.lens.inner_retry) ZP(N, S, true); else ZP(N, S, false); } while (0)
// straight-line instantiation when the calibrated stage boundary is the usual one (LUT sampling only)
#define ZPN(N, S)                                                          \
    do {                                                                   \
        bool fixed = false;                                                \
        if constexpr (kLut) {                                              \
            if (cam.lens.split == S) { ZPI(N, S); fixed = true; }          \
        }                                                                  \
        if (!fixed) ZPI(N, -1);                                            \
    } while (0)
        if (cam.lens.pretest) {   // first-surface-dominated cameras: element count and stage boundary do not matter to stage A
            ZPX(0, 0, true, true);
            if (launches) *launches += 1;
            continue;
        }
#ifdef ZOICB_POOL2_TUNE   // tuning builds (tools/build_variants.py): only the two benchmark cameras, quick to compile
        if constexpr (kLut && !kImage) {
            static const int force_inner = [] { const char* v = getenv("ZOICB_INNER"); return v ? atoi(v) : -1; }();
            const bool inner = force_inner >= 0 ? force_inner != 0 : cam.lens.inner_retry != 0;
            if (cam.lens.count == 11 && cam.lens.split == 1 && !inner) ZP(11, 1, false);
            else if (cam.lens.count == 11 && cam.lens.split == 1 && inner) ZP(11, 1, true);
            else if (cam.lens.count == 12 && cam.lens.split == 6 && cam.lens.inner_retry) ZP(12, 6, true);
            else return cudaErrorNotSupported;
        } else {
            return cudaErrorNotSupported;
        }
#else
        switch (cam.lens.count) {  // unrolled instantiations for the element counts of the shipped lens tables
            case 7: ZPN(7, 1); break;
            case 8: ZPN(8, 1); break;
            case 9: ZPN(9, 1); break;
            case 11: ZPN(11, 1); break;
            case 12: {
                bool six = false;
                if constexpr (kLut) {
                    if (cam.lens.split == 6) { ZPI(12, 6); six = true; }
                }
                if (!six) ZPN(12, 1);
                break;
            }
            default: ZPI(0, -1); break;
        }
#endif
#undef ZPN
#undef ZPI
#undef ZP
#undef ZPX
        if (launches) *launches += 1;
    }
    return cudaGetLastError();
}

cudaError_t launch_kolb_pool2(const CameraState& cam, const float4* samples, uint64_t n, uint64_t first_index, uint64_t seed,
                              RayRecord* rays, DeviceStats* stats, cudaStream_t st, const Workspace& ws, size_t rows_smem,
                              int* launches) {
    const bool image = cam.use_image != 0, lut = cam.lens.use_lut != 0;
#define ZL(I, U) launch_pool2_variant<I, U>(cam, samples, n, first_index, seed, rays, stats, st, ws, rows_smem, launches)
    if (image) return lut ? ZL(true, true) : ZL(true, false);
    return lut ? ZL(false, true) : ZL(false, false);
#undef ZL
}

}  // namespace zoicb
