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
    bool valid() const { return w > 0 && h > 0; }
};

struct HostCamera {
    zoicb_params params;           // string members are cleared; see lens_path
    std::string lens_path;
    CameraState state;             // bokeh pointers are filled in by the C-ABI after upload
    zoicb_constants constants;
    HostBokeh bokeh;
    std::vector<LensRow> rows;     // rear element first, after all rescaling
};

// A callback that classifies LUT candidate rays (accept = passes the whole stack).  The C-ABI passes a
// GPU implementation; tests may pass nothing to use the host threads.
typedef bool (*LutTraceFn)(void* user, const LensState& lens, const float* film_x, int n_film,
                           const uint32_t* draws, int samples_per_film, uint8_t* accept);

// Returns ZOICB_OK or an error status; `err` receives a message.
zoicb_status build_camera(const zoicb_params& p, const float* rgb, int w, int h, int nch, HostCamera* out,
                          std::string* err, LutTraceFn lut_fn = nullptr, void* lut_user = nullptr);

}  // namespace zoicb
