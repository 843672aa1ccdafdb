// bokeh_build.cu -- the image-based aperture tables built on the GPU (reference
// imageData::bokehProbability, src/zoic.cpp:222-417).
//
// The tables have to equal the reference's entry for entry, and the reference builds them with SEQUENTIAL fp32
// sums (one running total over all pixels, one per row, one running CDF per row) and with std::sort.  Neither may
// be re-associated, so the build is cut by what each step allows:
//   * per pixel, independent: luminance, pdf = lum * (1/total), conditional pdf = pdf / row mass    -> one thread per pixel
//   * per row, sequential along the row: row mass, column sort, running column CDF, guide table      -> one thread per row
//     (rows are staged through shared memory in 32 x 32 tiles so the global reads stay coalesced)
//   * sequential over the whole image: the grand total                                               -> one thread adds,
//     the other warps of the CTA stream the next tile into shared memory meanwhile (4 cycles per pixel, the
//     FADD latency: 0.13 ms for 255 x 255)
//   * over the rows: row sort, running row CDF, row guide table                                      -> one thread
// The sorts are gnu_sort.h (libstdc++'s introsort restated, ties included).  Everything stays on the device; the
// C ABI downloads a copy for zoicb_get_bokeh_tables and the host-side calibration.
#include <cuda_runtime.h>
#include <stdint.h>

#include "gnu_sort.h"
#include "kernels.h"
#include "lens_math.cuh"

namespace zoicb {
namespace {

constexpr int kTotalTile = 4096;   // floats per shared-memory tile of the grand-total kernel
constexpr int kTotalThreads = 288; // warp 0 adds, warps 1..8 load

// lum = r*0.3 + g*0.59 + b*0.11 with the reference's association (src/zoic.cpp:234-262)
__global__ void bokeh_lum_kernel(const float* __restrict__ rgb, int nch, size_t np, float* __restrict__ lum) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < np; i += (size_t)gridDim.x * blockDim.x) {
        const float* px = rgb + i * nch;   // nch >= 3 here: the reference ignores images with fewer channels (:135-137)
        lum[i] = xadd(xadd(xmul(px[0], 0.3f), xmul(px[1], 0.59f)), xmul(px[2], 0.11f));
    }
}

// total = ((lum[0] + lum[1]) + lum[2]) + ...   (src/zoic.cpp:262); out[0] = total, out[1] = 1/total (:270)
__global__ void __launch_bounds__(kTotalThreads, 1)
bokeh_total_kernel(const float* __restrict__ lum, size_t np, float* __restrict__ out) {
    __shared__ float tile[2][kTotalTile];
    const int t = threadIdx.x;
    const size_t ntiles = (np + kTotalTile - 1) / kTotalTile;
    auto load = [&](size_t k) {
        if (t < 32) return;
        const size_t base = k * kTotalTile;
        for (int j = t - 32; j < kTotalTile; j += kTotalThreads - 32) {
            const size_t i = base + j;
            tile[k & 1][j] = i < np ? lum[i] : 0.0f;
        }
    };
    float acc = 0.0f;
    load(0);
    __syncthreads();
    for (size_t k = 0; k < ntiles; ++k) {
        if (k + 1 < ntiles) load(k + 1);
        if (t == 0) {
            const size_t left = np - k * kTotalTile;
            const int m = left < (size_t)kTotalTile ? (int)left : kTotalTile;
            const float* p = tile[k & 1];
            int j = 0;
#pragma unroll 1
            for (; j + 8 <= m; j += 8) {
                const float4 a = *reinterpret_cast<const float4*>(p + j);
                const float4 b = *reinterpret_cast<const float4*>(p + j + 4);
                acc = xadd(acc, a.x); acc = xadd(acc, a.y); acc = xadd(acc, a.z); acc = xadd(acc, a.w);
                acc = xadd(acc, b.x); acc = xadd(acc, b.y); acc = xadd(acc, b.z); acc = xadd(acc, b.w);
            }
            for (; j < m; ++j) acc = xadd(acc, p[j]);
        }
        __syncthreads();
    }
    if (t == 0) { out[0] = acc; out[1] = xdiv(1.0f, acc); }
}

// pdf = lum * (1/total) in place (:273-276); row_mass[r] = running sum along the row (:279-292).
// One warp per 32 rows; 32 x 32 tiles through shared memory, lane k then owns row r0 + k.
__global__ void __launch_bounds__(32)
bokeh_pdf_rowmass_kernel(float* __restrict__ pdf, const float* __restrict__ total, int w, int h, float* __restrict__ row_mass) {
    __shared__ float tile[32][33];
    const int lane = threadIdx.x;
    const int r0 = blockIdx.x * 32;
    const float inv_total = total[1];
    float acc = 0.0f;
    for (int c0 = 0; c0 < w; c0 += 32) {
        const int cols = w - c0 < 32 ? w - c0 : 32;
        for (int k = 0; k < 32; ++k) {
            const int r = r0 + k;
            if (r < h && lane < cols) {
                const size_t i = (size_t)r * w + c0 + lane;
                const float v = xmul(pdf[i], inv_total);
                pdf[i] = v;
                tile[k][lane] = v;
            }
        }
        __syncwarp();
        if (r0 + lane < h)
            for (int c = 0; c < cols; ++c) acc = xadd(acc, tile[lane][c]);
        __syncwarp();
    }
    if (r0 + lane < h) row_mass[r0 + lane] = acc;
}

// rows by descending mass (:305-327), running row CDF (:330-341), row guide table
__global__ void bokeh_rows_kernel(const float* __restrict__ row_mass, int h, int32_t* __restrict__ row_idx,
                                  float* __restrict__ cdf_row, int row_shift, uint16_t* __restrict__ row_guide) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    for (int r = 0; r < h; ++r) row_idx[r] = r;
    gnusort::sort(row_idx, (long)h, gnusort::Before<int32_t>{row_mass});
    float run = 0.0f;
    for (int r = 0; r < h; ++r) { run = xadd(run, row_mass[row_idx[r]]); cdf_row[r] = run; }
    build_guide_table(cdf_row, h, row_shift, row_guide);
}

// conditional pdf of a pixel within its row (:344-362), in place over pdf
__global__ void bokeh_cond_kernel(float* __restrict__ pdf, const float* __restrict__ row_mass, int w, size_t np) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < np; i += (size_t)gridDim.x * blockDim.x) {
        const float m = row_mass[i / w];
        const float p = pdf[i];
        pdf[i] = (p != 0.0f && m != 0.0f) ? xdiv(p, m) : 0.0f;
    }
}

// per row: columns by descending conditional pdf (:365-391), running column CDF (:394-414), guide table.
// The reference sorts global pixel indices r*w + c with values looked up in the whole image; sorting the
// row-relative index c against the row's own values makes the same comparisons and the same moves.
__global__ void bokeh_columns_kernel(const float* __restrict__ cond, int w, int h, int32_t* __restrict__ scratch_idx,
                                     float* __restrict__ cdf_col, uint16_t* __restrict__ rel_col,
                                     int col_shift, uint16_t* __restrict__ col_guide) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= h) return;
    const float* v = cond + (size_t)r * w;
    int32_t* idx = scratch_idx + (size_t)r * w;
    for (int c = 0; c < w; ++c) idx[c] = c;
    gnusort::sort(idx, (long)w, gnusort::Before<int32_t>{v});
    float* cdf = cdf_col + (size_t)r * w;
    uint16_t* rel = rel_col + (size_t)r * w;
    float run = 0.0f;
    for (int c = 0; c < w; ++c) {
        const int32_t k = idx[c];
        run = xadd(run, v[k]);
        cdf[c] = run;
        rel[c] = (uint16_t)k;
    }
    build_guide_table(cdf, w, col_shift, col_guide + (size_t)r * (((size_t)1 << col_shift) + 2));
}

}  // namespace

cudaError_t launch_bokeh_build(const float* d_rgb, int w, int h, int nch, float* d_work, int32_t* d_scratch_idx,
                               float* d_total, float* d_row_mass, float* d_cdf_row, int32_t* d_row_idx,
                               float* d_cdf_col, uint16_t* d_rel_col, int row_shift, int col_shift, uint16_t* d_row_guide,
                               uint16_t* d_col_guide, cudaStream_t st, int* launches) {
    const size_t np = (size_t)w * h;
    if (np == 0) return cudaSuccess;
    const unsigned px_grid = (unsigned)((np + 255) / 256 < 148 * 8 ? (np + 255) / 256 : 148 * 8);
    bokeh_lum_kernel<<<px_grid, 256, 0, st>>>(d_rgb, nch, np, d_work);
    bokeh_total_kernel<<<1, kTotalThreads, 0, st>>>(d_work, np, d_total);
    bokeh_pdf_rowmass_kernel<<<(h + 31) / 32, 32, 0, st>>>(d_work, d_total, w, h, d_row_mass);
    bokeh_rows_kernel<<<1, 32, 0, st>>>(d_row_mass, h, d_row_idx, d_cdf_row, row_shift, d_row_guide);
    bokeh_cond_kernel<<<px_grid, 256, 0, st>>>(d_work, d_row_mass, w, np);
    bokeh_columns_kernel<<<(h + 31) / 32, 32, 0, st>>>(d_work, w, h, d_scratch_idx, d_cdf_col, d_rel_col, col_shift, d_col_guide);
    if (launches) *launches += 6;
    return cudaGetLastError();
}

}  // namespace zoicb
