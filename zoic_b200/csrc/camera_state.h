// camera_state.h -- plain-old-data camera state shared by the host setup code and the CUDA kernels.
//
// The host (host_setup.cpp) derives these constants bit-exactly the way the reference's node_update does
// (reference src/zoic.cpp:1575-1720) and the kernels receive the whole struct as a __grid_constant__
// kernel parameter (constant bank, broadcast to all lanes).
#pragma once
#include <stdint.h>

namespace zoicb {

constexpr int kMaxElements = 24;
constexpr int kLutSize = 32;
constexpr int kMaxBokehRows = 5120;  // row CDF + row indices are staged in 40 KB of shared memory
// Guide ("cutpoint") tables of the inverse-CDF searches have G + 2 entries with G = 2^shift cells (see BokehTables).
// Default resolution: the smallest power of two that is >= the table length, within [2^4, 2^16].
inline int default_guide_shift(int n) {
    int m = 4;
    while (m < 16 && (1 << m) < n) ++m;
    return m;
}
constexpr int kMaxTries = 25;  // reference src/zoic.cpp:1767

// One refracting surface, rear element (nearest the sensor) first.  Everything the march needs per
// element visit is precomputed once on the host with the reference's own rounding.
struct alignas(16) Element {
    float center;     // z of the sphere centre (src/zoic.cpp:963-969)
    float radius;     // signed radius of curvature R, cm (the stop is the R = 9999.9 "sphere")
    float radius2;    // fl(R*R)
    float sgn;        // R < 0 ? -1 : 1
    float rim2;       // accept iff hx^2+hy^2 <= rim2: largest float <= (aperture/2)^2 evaluated in double
                      // (src/zoic.cpp:1114), min'ed with fl(userApertureRadius^2) on the stop (:1115)
    float eta;        // ior_i / ior_next, or ior_i when the next medium is 1.0 (src/zoic.cpp:1013)
    float eta2;       // fl(eta*eta)
    float inv_radius; // fl(1/R)       (fast path only)
    int32_t tir_possible;  // ior_i > ior_next (src/zoic.cpp:1019)
    // guarded fast path: a rim test whose margin |h2 - rim2| is below
    //   rim2_guard + dt_guard * |hit.xy . dir.xy|   is "undecided" and the ray is re-run exactly
    float rim2_guard;      // relative part: accumulated fp32 drift of the ray state
    float dt_guard;        // 2 * (largest plausible error of the reference's own ray parameter t at this surface)
    float vertex;          // z where the sphere the REFERENCE intersects (centre fl(z - R), radius^2 fl(R*R))
                           // crosses the axis: fl(center + R)
    float r2_corr;         // R*R - fl(R*R) evaluated in double: |o-c|^2 - radius2 = dz*(dz-2R) + ox^2+oy^2 + r2_corr
    float miss_guard;      // 1e-5 * radius2: |discriminant| below this => the hit/miss test is undecided
    float vertex_m2r;      // vertex - 2R
    float one_m_eta2;      // fl(1 - eta^2) evaluated in double: 1 - cs2 = (1 - eta^2) + eta^2 c1^2 in one fma
};

struct LensState {
    int32_t count;
    int32_t aperture_element;
    int32_t use_lut;
    int32_t lut_size;
    float origin_shift;       // film plane z (negative), src/zoic.cpp:1675
    float half_sensor;        // fl(sensorWidth * 0.5)
    float first_aperture;     // lenses[0].aperture (a diameter used as a half extent -- kept)
    float neg_first_thickness;  // -lenses[0].thickness
    float user_aperture_radius;
    int32_t split;            // guarded kernel: surfaces [0, split) run in stage A, [split, count) in stage B
    int32_t inner_retry;      // guarded kernel: rays stopped in stage A re-sample inside the pass (most attempts die there)
    int32_t pretest;          // guarded kernel: most attempts die at the first surface -> rim pre-test flavour (kolb_pool2.cu)
    int32_t pad2;
    float lut_scale[kLutSize];  // boundingBox2d::getMaxScale() per LUT entry (src/zoic.cpp:503-517)
    float lut_cx[kLutSize];     // boundingBox2d::getCentroid().x per LUT entry
    Element e[kMaxElements];    // 16-byte aligned: the packed kernel reads an element as four 128-bit constant loads
};

struct ThinState {
    float tan_fov;          // src/zoic.cpp:1607
    float aperture_radius;  // src/zoic.cpp:1608
    float focal_distance;
    float ov_distance;      // opticalVignettingDistance
    float ov_radius_true;   // fl(apertureRadius * opticalVignettingRadius), src/zoic.cpp:1302
    int32_t use_dof;
    int32_t use_ov;         // opticalVignettingDistance > 0
    float ov_guard;         // guarded fast path: |hyp - ov_radius_true| below this => undecided
    // The reference's test is sqrt(s) < ov_radius_true with s = qx^2 + qy^2 (:1302-1304).  A correctly rounded square root is
    // monotone, so the s that pass are exactly those below one float: ov_s_threshold = the smallest s >= 0 whose rounded
    // root is >= ov_radius_true (0 when the radius is <= 0 or NaN, +inf when every finite s passes).  The kernels compare
    // s < ov_s_threshold -- the same decision for every s (NaN and +inf fail both forms) without the IEEE root.
    float ov_s_threshold;
};

// Image-based aperture sampling tables (device pointers), src/zoic.cpp:117-122
struct BokehTables {
    const float* cdf_row;        // [h]
    const int32_t* row_indices;  // [h]
    const float* cdf_column;     // [h*w], rows in ORIGINAL row order (indexed by actual row * w)
    const uint16_t* rel_column;  // [h*w], columnIndices[c] - row*w
    // Guide ("cutpoint") tables of the two inverse-CDF searches, with EXACT cells: G = 2^shift cells per table, and
    // the cell of u is k = floor(u * G) -- exact in fp32 because G is a power of two, so no rounding margins are
    // needed.  guide[k] = min(upper_bound(cdf, k / G), T) for k <= G and guide[G + 1] = T, where T is the start of
    // the CDF's flat tail (the first entry that carries the final value; the zero-probability pixels around the
    // aperture shape).  For 0 <= u < final the answer of std::upper_bound lies in [guide[k], guide[k + 1]] (k clamped
    // to G), usually zero to two entries; u >= final => n.  The search result stays std::upper_bound's.
    const uint16_t* row_guide;   // [2^row_shift + 2]
    const uint16_t* col_guide;   // [h * (2^col_shift + 2)], by actual row like cdf_column
    // the lens coordinates of a pixel (src/zoic.cpp:441,466,479-484), tabulated with the reference's arithmetic:
    // dx_of_col[c] = fl(fl(fl(c - (h-1)/2) / fl(w)) * 2), dy_of_row[r] = fl(fl(-fl(r - (w-1)/2) / fl(h)) * 2)
    const float* dx_of_col;      // [w]
    const float* dy_of_row;      // [h]
    int32_t w, h;
    int32_t row_shift, col_shift;   // log2 of the guide resolutions
};

// Byte-wide copies of the two big 16-bit column tables, for images of at most 255 columns (every entry is a column
// number or a column count, <= w): the thin-lens retry kernel is bound by the latency of its table loads, and the
// working set of config 3 (column CDF 167 KB + guide 262 KB + pixel indices 84 KB) is more than twice the L1 it runs in;
// the narrow copies take 197 KB off it.  Null when absent (wider images, the raytraced model).
struct BokehCompact {
    const uint8_t* col_guide8;    // [h * (2^col_shift + 2)]
    const uint8_t* rel_column8;   // [h*w]
};
constexpr int kCompactMaxWidth = 255, kCompactMaxRows = 512;   // rows: the kernel also stages a third row table (kernels.cu)

struct CameraState {
    int32_t lens_model;  // 0 thin lens, 1 raytraced
    int32_t use_image;
    float weight_scale;  // exposure: 1+e^2 (e>0), 1/(1+e^2) (e<0), 1 (src/zoic.cpp:1981-1987)
    float guard_scale;   // 1 = shipped decision margins (scales the fixed miss / TIR margins of the fast path)
    ThinState thin;
    BokehTables bokeh;
    LensState lens;
    BokehCompact compact;
};

// per-launch counters, accumulated with one atomicAdd per block per counter
struct DeviceStats {
    unsigned long long rays, success, vignetted, tir, attempts, element_visits, exact_reruns, pad;
};

}  // namespace zoicb
