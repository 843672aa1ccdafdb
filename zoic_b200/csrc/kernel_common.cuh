// kernel_common.cuh -- device helpers shared by the kernel translation units: image-based aperture sampling,
// per-block counter reduction, the EXACT per-sample functions, fast-math primitives.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"
#include "lens_math.cuh"

#ifndef ZOICB_BOKEH_COUNT
#define ZOICB_BOKEH_COUNT 8   // brackets of up to this many entries (warp maximum) are resolved by counting; 0: always halve
#endif

namespace zoicb {

#ifndef ZOICB_CHUNK
#define ZOICB_CHUNK 2048
#endif
constexpr int kChunk = ZOICB_CHUNK;   // samples handed to a warp per grab of the global cursor (multiple of 64)

// One output ray = one 32-byte record (zoicb_ray): origin.xyz, weight, dir.xyz, tries.  A ray is written by a
// single 256-bit store (STG.E.256), i.e. one full 32-byte sector, whatever order rays finish in.
__device__ __forceinline__ void store_ray(RayRecord* __restrict__ rays, uint64_t idx, float4 o, float4 d) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(rays + idx), "f"(o.x), "f"(o.y), "f"(o.z),
                 "f"(o.w), "f"(d.x), "f"(d.y), "f"(d.z), "f"(d.w)
                 : "memory");
}

// ------------------------------------------------------------------------------------------------
// image-based aperture sampling (reference imageData::bokehSample, src/zoic.cpp:420-485)
// ------------------------------------------------------------------------------------------------
// std::upper_bound over a[0..n): first index whose value is greater than u.  The guide table (camera_state.h) has
// G = 2^shift EXACT cells: k = floor(u G) is exact in fp32, so for 0 <= u < final the answer lies in
//   [guide[k], guide[k + 1]]           (k clamped to G; guide[G + 1] = T, the start of the CDF's flat tail)
// -- zero to two entries for an aperture-shaped image at G >= n -- and u >= final is answered directly (= n).  The guide
// entries are clamped to T (the zero-probability pixels around the aperture shape all carry the final value), so no
// bracket ever spans the flat tail.
// Brackets of up to ZOICB_BOKEH_COUNT entries (warp maximum) are resolved by COUNTING the entries that are not greater
// than u: independent, predicated loads (a lane only loads the entries of its own bracket) instead of a chain of
// dependent probes.  Longer brackets (photograph-like images whose CDF is nearly flat towards 1), NaN and negative u
// (whole range) take the libstdc++ first/len halving with a warp-uniform number of rounds and predicated updates: no
// lane leaves the loop early.  (With a data-dependent trip count ptxas 12.9 let the early lanes run ahead and re-use the
// uniform registers that hold the table pointers while the late lanes were still reading them.)
// `last()` = the table's final entry, load(n - 1), for callers that have it somewhere cheaper
template <typename Load, typename Guide, typename Last>
__device__ __forceinline__ int upper_bound_guided(int n, int shift, float u, Load load, Guide guide, Last last) {
    int first = 0, len = n;
    bool past = false;
    if (u >= 0.0f) {
        const int G = 1 << shift;
        const float f = u * (float)G;   // exact: G is a power of two
        const int k = f >= (float)G ? G : (int)f;
        first = guide(k);
        len = guide(k + 1) - first;
        past = u >= last();   // false for a NaN table (black image): its guides all hold n
    }
    const int maxlen = (int)__reduce_max_sync(__activemask(), (unsigned)len);
    constexpr int kCount = ZOICB_BOKEH_COUNT;
    if (kCount > 0 && maxlen <= kCount) {
        int cnt = 0;
#pragma unroll
        for (int j = 0; j < kCount; ++j) {
            if (j >= maxlen) break;   // warp-uniform
            if (j < len) cnt += (u < load(first + j)) ? 0 : 1;
        }
        return past ? n : first + cnt;
    }
    const int rounds = 32 - __clz(maxlen);
    for (int it = 0; it < rounds; ++it) {
        const int half = len >> 1;
        const int mid = first + half;
        const float v = load(mid < n ? mid : n - 1);
        const bool live = len > 0;
        const bool left = u < v;
        first = (live && !left) ? mid + 1 : first;
        len = live ? (left ? half : len - half - 1) : 0;
    }
    return past ? n : first;
}

template <typename Load, typename Guide>
__device__ __forceinline__ int upper_bound_guided(int n, int shift, float u, Load load, Guide guide) {
    return upper_bound_guided(n, shift, u, load, guide, [&]() { return load(n - 1); });
}

// The row tables (cdfRow, rowIndices: 8 bytes per image row) are always staged in dynamic shared memory --
// s_rows[0..h) holds the CDF, s_rows[h..2h) the row indices -- and addressed as shared memory (no generic
// pointers); the per-row column tables stay in global memory (L1/L2 resident).
extern __shared__ float s_rows[];

#ifndef ZOICB_THIN_SPECULATE
#define ZOICB_THIN_SPECULATE 1
#endif

__host__ __device__ inline unsigned bokeh_smem_bytes(int h) { return ((unsigned)h * 8u + 15u) & ~15u; }

struct BokehView {
    const float* cdf_col;     // global
    const uint16_t* rel_col;
    const uint16_t* row_guide;
    const uint16_t* col_guide;
    const float* dx_of_col;
    const float* dy_of_row;
    int w, h;
    int row_shift, col_shift;
    const uint8_t* col_guide8;   // camera_state.h: BokehCompact (kCompact callers only)
    const uint8_t* rel_col8;
};

// kCompact: the byte-wide column tables (BokehCompact) and the rows' final CDF values from shared memory,
// s_rows[h + actual row] (staged by the caller, stage_bokeh_compact): the same values from smaller / nearer places.
template <bool kCompact = false>
__device__ __forceinline__ void bokeh_sample(const BokehView& b, float u_row, float u_col, float* dx, float* dy) {
    // (a straight-line form of the row search for brackets of at most two entries, like the column search below, is no
    // faster: the row tables are in shared memory, there is no chain of long loads to shorten; profiles/r02_ab.txt call 26)
    int r = upper_bound_guided(b.h, b.row_shift, u_row, [&](int i) { return s_rows[i]; }, [&](int k) { return (int)__ldg(b.row_guide + k); });
    if (r >= b.h) r = b.h - 1;
    // kCompact: rows as [CDF: h floats][final column-CDF value by actual row: h floats][row indices: h x 16 bit] (stage_bokeh_compact)
    const int row = kCompact ? (int)reinterpret_cast<const uint16_t*>(s_rows + 2 * b.h)[r] : __float_as_int(s_rows[b.h + r]);
    const int start = row * b.w;
    const float* __restrict__ col = b.cdf_col + start;
    const int goff = row * ((1 << b.col_shift) + 2);
    int c, rel;
    if (kCompact) {
        const uint8_t* __restrict__ cg = b.col_guide8 + goff;
        const uint8_t* __restrict__ rl = b.rel_col8 + start;
#if ZOICB_THIN_SPECULATE
        // The chain guide -> CDF entries -> pixel index -> lens coordinate is four dependent loads, and the kernel waits
        // on them (long scoreboard is its first stall).  With brackets of at most two entries (the usual case at
        // G >= 2 w) the answer is one of first, first + 1, first + 2: their pixel indices are loaded NEXT TO the CDF
        // entries instead of after them, and the count picks one.  Same search result (upper_bound_guided's counting
        // branch, spelled out); every other case -- longer brackets, u < 0 or NaN in any lane -- takes the general path.
        int first = 0, len = b.w;
        bool past = false;
        if (u_col >= 0.0f) {
            const int G = 1 << b.col_shift;
            const float f = u_col * (float)G;   // exact: G is a power of two
            const int k = f >= (float)G ? G : (int)f;
            first = (int)__ldg(cg + k);
            len = (int)__ldg(cg + k + 1) - first;
            past = u_col >= s_rows[b.h + row];
        }
        const int maxlen = (int)__reduce_max_sync(__activemask(), (unsigned)len);
        if (maxlen <= 2) {
            const int w1 = b.w - 1;
            const int r0 = (int)__ldg(rl + min(first, w1)), r1 = (int)__ldg(rl + min(first + 1, w1));
            int cnt = 0;
            if (0 < len) cnt += (u_col < __ldg(col + first)) ? 0 : 1;
            rel = cnt == 0 ? r0 : r1;
            if (maxlen == 2) {   // warp-uniform
                const int r2 = (int)__ldg(rl + min(first + 2, w1));
                if (1 < len) cnt += (u_col < __ldg(col + first + 1)) ? 0 : 1;
                rel = cnt == 2 ? r2 : rel;
            }
            if (past) rel = (int)__ldg(rl + w1);   // upper_bound = n, clamped to the last entry
        } else
#endif
        {
            c = upper_bound_guided(b.w, b.col_shift, u_col, [&](int i) { return __ldg(col + i); }, [&](int k) { return (int)__ldg(cg + k); },
                                   [&]() { return s_rows[b.h + row]; });
            if (c >= b.w) c = b.w - 1;
            rel = (int)__ldg(rl + c);
        }
    } else {
        const uint16_t* __restrict__ cg = b.col_guide + goff;
        c = upper_bound_guided(b.w, b.col_shift, u_col, [&](int i) { return __ldg(col + i); }, [&](int k) { return (int)__ldg(cg + k); });
        if (c >= b.w) c = b.w - 1;
        rel = (int)__ldg(b.rel_col + start + c);
    }
    // the reference centres the row with the WIDTH (:441) and the column with the HEIGHT (:466) and divides (:479-484):
    // both divisions depend on the column / the row only and are tabulated once per camera with the same operations
    *dx = __ldg(b.dx_of_col + rel);
    *dy = __ldg(b.dy_of_row + row);
}

template <bool kImage, bool kCompact = false>
__device__ __forceinline__ void lens_sample(const BokehView& b, float u, float v, float* lx, float* ly) {
    if (kImage) bokeh_sample<kCompact>(b, u, v, lx, ly);
    else concentric_disk(u, v, lx, ly);
}

// two draws of the per-sample stream; the FIRST draw feeds the SECOND parameter (g++ evaluates the
// reference's argument lists right to left; pinned in tests/test_oracle_golden.py::test_port_draw_order_matches_reference)
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
    b.row_guide = cam.bokeh.row_guide;
    b.col_guide = cam.bokeh.col_guide;
    b.dx_of_col = cam.bokeh.dx_of_col;
    b.dy_of_row = cam.bokeh.dy_of_row;
    b.row_shift = cam.bokeh.row_shift; b.col_shift = cam.bokeh.col_shift;
    b.col_guide8 = cam.compact.col_guide8;
    b.rel_col8 = cam.compact.rel_column8;
    for (int i = threadIdx.x; i < b.h; i += blockDim.x) {
        s_rows[i] = cam.bokeh.cdf_row[i];
        s_rows[b.h + i] = __int_as_float(cam.bokeh.row_indices[i]);
    }
    __syncthreads();
    return b;
}
// The row tables of the kCompact callers: [CDF: h floats][final value of every row's column CDF, by actual row: h floats]
// [row indices: h x 16 bit] -- 10 bytes per row (compact_smem_bytes), so that eight CTAs' tables of a 255-row image
// still fit the 32 KB carve-out.
__host__ __device__ inline unsigned compact_smem_bytes(int h) { return ((unsigned)h * 10u + 15u) & ~15u; }
__device__ __forceinline__ BokehView stage_bokeh_compact(const CameraState& cam) {
    BokehView b;
    b.w = cam.bokeh.w; b.h = cam.bokeh.h;
    b.cdf_col = cam.bokeh.cdf_column;
    b.rel_col = cam.bokeh.rel_column;
    b.row_guide = cam.bokeh.row_guide;
    b.col_guide = cam.bokeh.col_guide;
    b.dx_of_col = cam.bokeh.dx_of_col;
    b.dy_of_row = cam.bokeh.dy_of_row;
    b.row_shift = cam.bokeh.row_shift; b.col_shift = cam.bokeh.col_shift;
    b.col_guide8 = cam.compact.col_guide8;
    b.rel_col8 = cam.compact.rel_column8;
    uint16_t* idx16 = reinterpret_cast<uint16_t*>(s_rows + 2 * b.h);
    for (int i = threadIdx.x; i < b.h; i += blockDim.x) {
        s_rows[i] = cam.bokeh.cdf_row[i];
        s_rows[b.h + i] = cam.bokeh.cdf_column[(size_t)i * b.w + (b.w - 1)];
        idx16[i] = (uint16_t)cam.bokeh.row_indices[i];
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
                // sqrt(s) < ov_radius_true, decided on s itself (camera_state.h: ov_s_threshold)
                if (xadd(xmul(qx, qx), xmul(qy, qy)) < T.ov_s_threshold) break;
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
// fast-math primitives of the guarded path
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float approx_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float approx_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float approx_rsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

enum { kUndecided = 3 };

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


// ------------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------------
inline int sm_count() {   // of the calling thread's current device (cached per device)
    static int counts[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    int c = counts[dev];
    if (!c) {
        cudaDeviceGetAttribute(&c, cudaDevAttrMultiProcessorCount, dev);
        if (c <= 0) c = 148;
        counts[dev] = c;
    }
    return c;
}

// kolb_pool2.cu (two rays per lane, packed fp32)
cudaError_t launch_kolb_pool2(const CameraState& cam, const float4* samples, uint64_t n, uint64_t first_index, uint64_t seed,
                              RayRecord* rays, DeviceStats* stats, cudaStream_t st, const Workspace& ws, size_t rows_smem,
                              int* launches);

}  // namespace zoicb
