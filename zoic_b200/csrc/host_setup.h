// host_setup.h -- host side of camera creation: lens-table parsing, the reference's setup pipeline
// (bit-exact), exit-pupil LUT and bokeh CDF tables.  Produces the CameraState the kernels consume.
#pragma once
#include <string>
#include <vector>

#include "../../include/zoicb.h"
#include "camera_state.h"

namespace zoicb {

struct LensRow { float curvature, thickness, ior, aperture, abbe, center; };

struct HostBokeh {
    int w = 0, h = 0;
    std::vector<float> cdf_row, cdf_column;
    std::vector<int32_t> row_indices, column_indices;
    std::vector<uint16_t> row_guide, col_guide;   // search accelerators for the device (camera_state.h); not part of parity
    int row_shift = 0, col_shift = 0;             // log2 of the guide resolutions; set by build_camera before the tables are built
    // The reference accepts an image with 1 or 2 channels but treats it as invalid (imageData::isValid wants >= 3,
    // src/zoic.cpp:135-137): no tables, and every bokehSample returns the lens centre (0, 0) (:420-425).  `degenerate`
    // marks such an image: w, h hold its size, the tables stay empty, and the device gets a 1 x 1 stand-in table whose
    // only answer is (0, 0).
    bool degenerate = false;
    bool valid() const { return w > 0 && h > 0 && !degenerate; }
};

struct HostCamera {
    zoicb_params params;           // string members are cleared; see lens_path
    std::string lens_path;
    CameraState state;             // bokeh pointers are filled in by the C-ABI after upload
    zoicb_constants constants;
    HostBokeh bokeh;
    std::vector<LensRow> rows;     // rear element first, after all rescaling
};

// A callback that classifies the exit-pupil LUT candidate rays (accept = passes the whole stack) and, if it can, also
// folds the accepted candidates of every film position into the reference's bounding box (src/zoic.cpp:1421-1440).
// Returns 0 on failure (the host threads take over), 1 when `accept` [n_film * per_film] is filled (the host replays the
// boxes), 2 when `boxes` [n_film][4] = (min.x, min.y, max.x, max.y) is filled.  The C-ABI passes the GPU implementation.
typedef int (*LutTraceFn)(void* user, const LensState& lens, const float* film_x, int n_film,
                          const uint32_t* draws, int samples_per_film, uint8_t* accept, float* boxes);

// the reference's in-order fold of one film position's accepted candidates, re-arm quirk (:1423) included
void lut_fold_boxes_host(const uint32_t* draws, const uint8_t* accept, int n_film, int per_film, float ap, float* boxes);

// the smallest float s >= 0 whose correctly rounded square root is >= r (0 for r <= 0 or NaN; +inf when r = +inf)
float sqrt_threshold(float r);

// A callback that builds the image-based aperture tables (all members of HostBokeh) from a validated image.
// The C-ABI passes the GPU build (bokeh_build.cu); without one the host statement runs
// (zoicb_setup_host_only, which has no device).
typedef bool (*BokehBuildFn)(void* user, const float* rgb, int w, int h, int nch, HostBokeh* out);

// What zoicb_create accepts as a bokeh image (>= 1 channel, <= 65535 columns, <= kMaxBokehRows rows).
zoicb_status check_bokeh_image(const float* rgb, int w, int h, int nch, std::string* err);
// idx = 0..n-1 ordered by the toolchain's std::sort with the reference's "greater by value" comparator shape
void std_sort_desc(const float* values, int n, int32_t* idx);

// Returns ZOICB_OK or an error status; `err` receives a message.
zoicb_status build_camera(const zoicb_params& p, const float* rgb, int w, int h, int nch, HostCamera* out,
                          std::string* err, LutTraceFn lut_fn = nullptr, void* lut_user = nullptr,
                          BokehBuildFn bokeh_fn = nullptr, void* bokeh_user = nullptr);

}  // namespace zoicb
