// kernels.h -- launchers for the sm_100a kernels (kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "camera_state.h"

namespace zoicb {

struct alignas(32) RayRecord { float4 origin_w; float4 dir_tries; };   // = zoicb_ray (include/zoicb.h)

// One undecided sample of the guarded kernel, handed to the exact kernel: ONE 64-bit word (the pool kernel's stores and
// register allocation are those of a plain index queue).  The attempts before `tries` were decided (stopped for certain)
// by the fast path, so the exact kernel RESUMES at attempt `tries`: it advances the sample's retry stream by that many
// draw pairs and adds the counters of the earlier attempts instead of marching them again.
//   bits 0..39  sample index within the generate call (zoicb_generate takes at most 2^40 samples per call)
//   bits 40..44 tries (<= 26)   bits 45..49 total internal reflections (<= 26)   bits 50..63 element visits (<= 26 x 24)
typedef unsigned long long QueueRecord;
__host__ __device__ inline QueueRecord queue_pack(unsigned long long index, unsigned tries, unsigned tir, unsigned visits) {
    return index | ((unsigned long long)tries << 40) | ((unsigned long long)tir << 45) | ((unsigned long long)visits << 50);
}
__host__ __device__ inline unsigned long long queue_index(QueueRecord q) { return q & ((1ull << 40) - 1); }
__host__ __device__ inline unsigned queue_tries(QueueRecord q) { return (unsigned)(q >> 40) & 31u; }
__host__ __device__ inline unsigned queue_tir(QueueRecord q) { return (unsigned)(q >> 45) & 31u; }
__host__ __device__ inline unsigned queue_visits(QueueRecord q) { return (unsigned)(q >> 50); }

// Per-stream scratch: counters[0] = chunk cursor of the main kernel, counters[1] = number of queued (undecided)
// samples, counters[2] = work cursor of the exact persistent kernel, counters[3] spare;
// queue[capacity] = the undecided samples.
struct Workspace {
    unsigned long long* counters;
    QueueRecord* queue;
    unsigned long long capacity;
};

cudaError_t launch_generate(const CameraState& cam, int mode, const float4* samples, uint64_t n, uint64_t first_index,
                            uint64_t seed, RayRecord* rays, DeviceStats* stats, cudaStream_t st, const Workspace& ws,
                            int* launches);
cudaError_t launch_synth(uint32_t W, uint32_t H, uint32_t spp, uint64_t seed, uint64_t first_index, uint64_t n,
                         float4* out, cudaStream_t st, int* launches);
cudaError_t launch_lut_trace(const LensState& L, const float* d_film_x, int n_film, int per_film, const uint32_t* d_draws,
                             uint8_t* d_accept, cudaStream_t st, int* launches);
// bounding boxes of the accepted LUT candidates, folded in order with the reference's re-arm quirk (:1423):
// d_boxes[f] = (min.x, min.y, max.x, max.y)
cudaError_t launch_lut_bbox(const uint32_t* d_draws, const uint8_t* d_accept, int n_film, int per_film, float ap,
                            float4* d_boxes, cudaStream_t st, int* launches);
cudaError_t launch_draw_paths(const CameraState& cam, const float4* samples, uint32_t n, const unsigned long long* indices,
                              uint64_t first_index, uint64_t seed,
                              float4* quads, uint8_t* kinds, uint32_t* counts, uint32_t cap, cudaStream_t st, int* launches);
cudaError_t launch_transform(const float* m3x4, const RayRecord* in, uint64_t n, RayRecord* out, cudaStream_t st, int* launches);
// Image-based aperture tables built on the device (bokeh_build.cu).  d_work [w*h] floats and d_scratch_idx [w*h]
// int32 are scratch; d_total [2]; d_row_mass [h]; the remaining pointers are the tables of BokehTables (guide tables:
// 2^row_shift + 2 and h * (2^col_shift + 2) entries).
cudaError_t launch_bokeh_build(const float* d_rgb, int w, int h, int nch, float* d_work, int32_t* d_scratch_idx,
                               float* d_total, float* d_row_mass, float* d_cdf_row, int32_t* d_row_idx,
                               float* d_cdf_col, uint16_t* d_rel_col, int row_shift, int col_shift, uint16_t* d_row_guide,
                               uint16_t* d_col_guide, cudaStream_t st, int* launches);
// ray differentials (differentials.cu): out = n x 12 floats (dOdx, dOdy, dDdx, dDdy)
cudaError_t launch_differentials(const CameraState& cam, const float4* samples, uint64_t n, uint64_t first_index, uint64_t seed,
                                 float dsx, float dsy, const RayRecord* rays, float4* out, cudaStream_t st, int* launches);
cudaError_t launch_transform_diffs(const float* m3x4, const float4* in, uint64_t n, float4* out, cudaStream_t st, int* launches);
// records -> planes (zoicb_generate_host_planar): six float planes of `stride` entries each, then `stride` bytes of flags
cudaError_t launch_pack_planar(const RayRecord* rays, uint64_t n, uint64_t stride, uint8_t* planes, cudaStream_t st, int* launches);
cudaError_t check_normalize_factor(unsigned long long* mismatches, unsigned* first_bad, int* launches);
cudaError_t measure_fp32_peak(double* tflops, int* launches);

}  // namespace zoicb
