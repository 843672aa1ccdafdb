// capi_internal.h -- what the translation units behind include/zoicb.h share: the context, error plumbing, the
// device guard.  Not part of the ABI.
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/zoicb.h"
#include "host_setup.h"
#include "kernels.h"

struct zoicb_ctx {
    int device = 0;
    int mode = ZOICB_MODE_GUARDED;
    zoicb::HostCamera host;
    // device tables (bokeh)
    float* d_cdf_row = nullptr;
    int32_t* d_row_idx = nullptr;
    float* d_cdf_col = nullptr;
    uint16_t* d_rel_col = nullptr;
    uint16_t* d_row_guide = nullptr;
    uint16_t* d_col_guide = nullptr;
    float* d_dxy = nullptr;   // dx_of_col[w] then dy_of_row[h]
    uint8_t* d_compact = nullptr;   // byte-wide column guide, then byte-wide pixel indices (camera_state.h: BokehCompact)
    zoicb::DeviceStats* d_stats = nullptr;
    // guarded-mode scratch, one per stream the caller uses (stream order serialises reuse).  gen_mu is held from the
    // workspace lookup to the last launch of a generate call, so two host threads driving the same context cannot
    // interleave their counter resets and kernels on one stream, and a queue is never freed under a caller.
    std::vector<float> base_guards;
    std::mutex gen_mu;
    std::map<cudaStream_t, zoicb::Workspace> workspaces;
    // host-buffer pipeline (zoicb_generate_host)
    static constexpr int kSlots = 3;
    uint64_t chunk = 0;
    cudaStream_t streams[kSlots] = {nullptr, nullptr, nullptr};
    float4* d_in[kSlots] = {nullptr, nullptr, nullptr};
    zoicb::RayRecord* d_r[kSlots] = {nullptr, nullptr, nullptr};
    float4* h_in[kSlots] = {nullptr, nullptr, nullptr};   // pinned staging, only for pageable callers
    zoicb::RayRecord* h_r[kSlots] = {nullptr, nullptr, nullptr};
    uint8_t* d_p[kSlots] = {nullptr, nullptr, nullptr};   // planar form of a chunk (zoicb_generate_host_planar): 25 bytes per ray
    uint8_t* h_p[kSlots] = {nullptr, nullptr, nullptr};   // pinned staging of the same, only for pageable callers
    std::mutex host_mu;
    // optional: events recorded around the device-side bokeh table build (zoicb_build_bokeh_tables)
    cudaEvent_t bokeh_ev0 = nullptr, bokeh_ev1 = nullptr;
    // buffers, streams and events of zoicb_run_job, kept between jobs (allocating and freeing tens of GB per job cost
    // more than a small job itself); one job at a time per context
    std::mutex job_mu;
    void* job_cache = nullptr;
    void (*job_cache_free)(void*) = nullptr;
    // wall time of the creation pipeline (zoicb_get_create_times)
    double create_ms = 0.0, lut_ms = 0.0, bokeh_ms = 0.0;
};

namespace zoicb {

// The calling thread's current CUDA device is saved, switched to `device` and restored on scope exit: no entry point
// leaves the process on another device than it found it (a host application may drive several GPUs from one thread).
class DeviceGuard {
public:
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev_) != cudaSuccess) { cudaGetLastError(); prev_ = -1; }
        status_ = (prev_ == device) ? cudaSuccess : cudaSetDevice(device);
        switched_ = prev_ != device && status_ == cudaSuccess;
    }
    ~DeviceGuard() { if (switched_ && prev_ >= 0) cudaSetDevice(prev_); }
    cudaError_t status() const { return status_; }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
private:
    int prev_ = -1;
    bool switched_ = false;
    cudaError_t status_ = cudaSuccess;
};

zoicb_status api_fail(zoicb_status code, const std::string& msg);
zoicb_status api_cuda_fail(cudaError_t e, const char* what);
void api_count_launches(int k);
// guarded-mode scratch of `st`, large enough for n samples; call with ctx->gen_mu held
cudaError_t api_get_workspace(zoicb_ctx* c, cudaStream_t st, uint64_t n, Workspace* out);

}  // namespace zoicb

#define ZCUDA(call, what)                                                \
    do {                                                                 \
        cudaError_t e__ = (call);                                        \
        if (e__ != cudaSuccess) return zoicb::api_cuda_fail(e__, what);  \
    } while (0)
#define ZGUARD(dev)                                                                                \
    zoicb::DeviceGuard guard__(dev);                                                               \
    if (guard__.status() != cudaSuccess) return zoicb::api_cuda_fail(guard__.status(), "cudaSetDevice")
