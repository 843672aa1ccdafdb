// differentials.cu -- ray differentials for a batch of generated rays (second half).
//
// The reference leaves AtCameraOutput's dOdx / dOdy / dDdx / dDdy unset (TODO at src/zoic.cpp:12-13) and papers over the
// visual problem with "if (tries > 0) { dOdy = origin; dDdy = dir; }" (:1971-1977).  This kernel computes what the TODO
// asks for: the derivative of the ray with respect to the screen position at a FIXED point of the aperture, as forward
// differences over one pixel.  Contract (stated for the CPU in oracle/zoic_port.cpp: zport_differentials; the GPU result
// equals it bit for bit):
//   1. weight 0 -> four zero vectors;
//   2. the aperture draw of the accepted attempt is (lensx, lensy) for tries == 0, else the tries-th pair of the sample's
//      retry stream, mapped to its aim point like camera_create_ray maps it (the sample's own exit-pupil LUT entry);
//   3. base ray (sx, sy), x-neighbour (fl(sx + dsx), sy) and y-neighbour (sx, fl(sy + dsy)) run from their film points
//      through that ONE aim point, with the exact arithmetic of lens_math.cuh;
//   4. dOdx = origin_x - origin_base, dDdx = dir_x - dir_base (fp32 subtractions), same for y; a neighbour stopped inside
//      the lens gives zero vectors for its axis.
// One thread per sample; 16 + 32 bytes read, 48 bytes written per ray.
#include <cuda_runtime.h>

#include "kernel_common.cuh"

namespace zoicb {

template <int kModel, bool kImage, bool kLut>
__global__ void __launch_bounds__(256)
differentials_kernel(const __grid_constant__ CameraState cam, const float4* __restrict__ samples, uint64_t n,
                     uint64_t first_index, uint64_t seed, float dsx, float dsy, const RayRecord* __restrict__ rays,
                     float4* __restrict__ out) {
    BokehView bk;
    if (kImage) bk = stage_bokeh(cam);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t rounds = (n + stride - 1) / stride;   // every thread runs every round: the table searches are warp-wide
    for (uint64_t it = 0; it < rounds; ++it) {
        const uint64_t i = it * stride + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        const bool live_thread = i < n;
        const uint64_t j = live_thread ? i : n - 1;
        const float4 s = samples[j];
        const float4 ow = rays[j].origin_w, dt = rays[j].dir_tries;
        const int tries = (int)dt.w;
        float u = s.z, v = s.w;
        if (tries > 0) {
            Xor128 rng = sample_stream(seed, first_index + j);
            for (int t = 0; t < tries; ++t) draw_pair(rng, &u, &v);
        }
        float lx, ly;
        lens_sample<kImage>(bk, u, v, &lx, &ly);
        Vec3 o[3], d[3];
        bool ok[3] = {true, true, true};
        const float fsx[3] = {s.x, xadd(s.x, dsx), s.x}, fsy[3] = {s.y, s.y, xadd(s.y, dsy)};
        if (kModel == 0) {
            const ThinState& T = cam.thin;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                d[k] = vnormalize(vmake(xmul(fsx[k], T.tan_fov), xmul(fsy[k], T.tan_fov), 1.0f));
                o[k] = vmake(0.0f, 0.0f, 0.0f);
                if (T.use_dof) {
                    o[k] = vmake(xmul(lx, T.aperture_radius), xmul(ly, T.aperture_radius), 0.0f);
                    const Vec3 focus = vscale(d[k], fabsf(xdiv(T.focal_distance, d[k].z)));
                    d[k] = vnormalize(vsub(focus, o[k]));
                }
                d[k].z = -d[k].z;
            }
        } else {
            const LensState& L = cam.lens;
            const KolbSampleState base = kolb_sample_setup<kLut, true>(L, s.x, s.y);
            float ax, ay;
            kolb_aim_point<kLut>(base, lx, ly, tries > 0, &ax, &ay);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                Ray r;
                r.o = vmake(xmul(fsx[k], L.half_sensor), xmul(fsy[k], L.half_sensor), L.origin_shift);
                r.d = vmake(xsub(ax, r.o.x), xsub(ay, r.o.y), L.neg_first_thickness);
                int visited;
                ok[k] = exact_march(L, r, &visited) == kPass;
                o[k] = vmake(-r.o.x, -r.o.y, -r.o.z);
                d[k] = vmake(-r.d.x, -r.d.y, -r.d.z);
            }
        }
        float4 q0 = make_float4(0, 0, 0, 0), q1 = q0, q2 = q0;   // dOdx.xyz dOdy.x | dOdy.yz dDdx.xy | dDdx.z dDdy.xyz
        if (ow.w != 0.0f && ok[0]) {
            if (ok[1]) {
                const Vec3 a = vsub(o[1], o[0]), b = vsub(d[1], d[0]);
                q0.x = a.x; q0.y = a.y; q0.z = a.z; q1.z = b.x; q1.w = b.y; q2.x = b.z;
            }
            if (ok[2]) {
                const Vec3 a = vsub(o[2], o[0]), b = vsub(d[2], d[0]);
                q0.w = a.x; q1.x = a.y; q1.y = a.z; q2.y = b.x; q2.z = b.y; q2.w = b.z;
            }
        }
        if (live_thread) {
            out[3 * i] = q0;
            out[3 * i + 1] = q1;
            out[3 * i + 2] = q2;
        }
    }
}

// camera -> world for the differentials: all four are DIFFERENCES of points or directions, so each is multiplied by the
// 3x3 part of the camera-to-world matrix, with the fma chain zoicb_transform_rays uses for directions:
//   v'_r = fma(m[r][0], vx, fma(m[r][1], vy, m[r][2] * vz))
struct DiffXform { float m[12]; };
__global__ void __launch_bounds__(256)
transform_diffs_kernel(const __grid_constant__ DiffXform X, const float4* __restrict__ in, uint64_t n, float4* __restrict__ out) {
    const float* m = X.m;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float4 a = in[3 * i], b = in[3 * i + 1], c = in[3 * i + 2];
        const float v[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
        float w[12];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float x = v[3 * k], y = v[3 * k + 1], z = v[3 * k + 2];
#pragma unroll
            for (int r = 0; r < 3; ++r)
                w[3 * k + r] = __fmaf_rn(m[4 * r], x, __fmaf_rn(m[4 * r + 1], y, __fmul_rn(m[4 * r + 2], z)));
        }
        out[3 * i] = make_float4(w[0], w[1], w[2], w[3]);
        out[3 * i + 1] = make_float4(w[4], w[5], w[6], w[7]);
        out[3 * i + 2] = make_float4(w[8], w[9], w[10], w[11]);
    }
}

cudaError_t launch_transform_diffs(const float* m3x4, const float4* in, uint64_t n, float4* out, cudaStream_t st, int* launches) {
    if (n == 0) return cudaSuccess;
    DiffXform X;
    for (int i = 0; i < 12; ++i) X.m[i] = m3x4[i];
    const uint64_t want = (n + 255) / 256, cap = (uint64_t)sm_count() * 8;
    transform_diffs_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(X, in, n, out);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_differentials(const CameraState& cam, const float4* samples, uint64_t n, uint64_t first_index, uint64_t seed,
                                 float dsx, float dsy, const RayRecord* rays, float4* out, cudaStream_t st, int* launches) {
    if (n == 0) return cudaSuccess;
    const bool image = cam.use_image != 0, lut = cam.lens.use_lut != 0;
    const size_t smem = image ? bokeh_smem_bytes(cam.bokeh.h) : 0;
    const uint64_t want = (n + 255) / 256, cap = (uint64_t)sm_count() * 4;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
#define ZD(M, I, U) differentials_kernel<M, I, U><<<grid, 256, smem, st>>>(cam, samples, n, first_index, seed, dsx, dsy, rays, out)
    if (cam.lens_model == 0) { if (image) ZD(0, true, false); else ZD(0, false, false); }
    else if (image) { if (lut) ZD(1, true, true); else ZD(1, true, false); }
    else { if (lut) ZD(1, false, true); else ZD(1, false, false); }
#undef ZD
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace zoicb
