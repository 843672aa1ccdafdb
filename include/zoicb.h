/* zoicb.h -- C ABI of libzoicb: batched, B200-native (sm_100a) camera-ray generation with the
 * behaviour of the zoic Arnold camera's camera_create_ray.
 *
 * This is the drop-in boundary.  Plain C types only: no C++ objects, no torch
 * types, no exceptions across the boundary.  Every entry point returns a zoicb_status; the text of
 * the last error of the calling thread is available from zoicb_last_error().
 *
 * What each entry point replaces in the reference (paths relative to the reference tree):
 *   zoicb_create        node_initialize + node_update            src/zoic.cpp:1565-1572, 1575-1720
 *   zoicb_generate      camera_create_ray, one call per BATCH    src/zoic.cpp:1752-1990
 *   zoicb_generate_host the same, host buffers in / out          src/zoic.cpp:1752-1990
 *   zoicb_generate_host_planar   the same, 25-byte planar output  src/zoic.cpp:1752-1990
 *   zoicb_get_stats     the counters printed by node_finish      src/zoic.cpp:1729-1732
 *   zoicb_destroy       node_finish                              src/zoic.cpp:1723-1749
 *   zoicb_transform_rays  (the renderer's camera-to-world step after camera_create_ray; nothing in zoic)
 *   zoicb_differentials   the ray derivatives zoic leaves as a TODO / fakes                 src/zoic.cpp:12-13, 1971-1977
 *   zoicb_write_draw_file writeToFile + the DRAW_ONLY ray dumps  src/zoic.cpp:1240-1293, 1121-1128, 1146-1153
 *   zoicb_params        the 14 node parameters                   src/zoic.cpp:1547-1562
 *   zoicb_run_job       the renderer's loop over camera_create_ray for a whole W x H x spp frame (nothing in zoic: Arnold
 *                       calls src/zoic.cpp:1752-1990 once per sample), tile by tile through rotating device buffers
 *   zoicb_gather_*      the final gather of the ray buffer to one GPU over NVLink (BASELINE north_star; nothing in zoic)
 *   zoicb_census        GUARDED-vs-EXACT comparison of every record on the device (test / bench instrument)
 * The Arnold-shaped per-sample surface (NodeLoader + the six node callbacks, src/zoic.cpp:1999-2007)
 * is exported by the same library from zoic_b200/csrc/arnold_adapter.cpp on top of these calls.
 *
 * There is no CPU fallback: every generate call runs CUDA kernels and fails with
 * ZOICB_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef ZOICB_H
#define ZOICB_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define ZOICB_API __attribute__((visibility("default")))
#else
#define ZOICB_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef enum zoicb_status {
    ZOICB_OK = 0,
    ZOICB_ERR_INVALID_ARGUMENT = 1,
    ZOICB_ERR_LENS_FILE = 2,      /* cannot open / parse the tabular lens file, bad column count   */
    ZOICB_ERR_LENS_DATA = 3,      /* more than one aperture stop, too many elements                */
    ZOICB_ERR_BOKEH_IMAGE = 4,    /* useImage set but no usable pixels                             */
    ZOICB_ERR_CUDA = 5,           /* no device, launch failure, allocation failure                 */
    ZOICB_ERR_UNSUPPORTED = 6
} zoicb_status;

/* lensModel values (reference enum LensModel, src/zoic.cpp:84-88) */
enum { ZOICB_THINLENS = 0, ZOICB_RAYTRACED = 1 };

/* Arithmetic mode of the ray kernels.
 * EXACT   : bit-for-bit the reference's fp32/double arithmetic (no FMA contraction, double where the
 *           reference promotes).  Identical accept/reject decisions, identical bits.
 * GUARDED : FMA / reciprocal fast path whose every accept/reject decision carries an error margin; a ray
 *           that comes within the margin of any threshold is re-run through the EXACT path, so the
 *           decision sequence (tries, zero weight) equals EXACT and values agree to ~1e-6.  Default. */
enum { ZOICB_MODE_EXACT = 0, ZOICB_MODE_GUARDED = 1 };

/* The 14 node parameters, same names, meaning, units (cm) and defaults as the reference
 * (src/zoic.cpp:1547-1562).  Use zoicb_default_params() to get the defaults. */
typedef struct zoicb_params {
    float sensorWidth;               /* 3.6   */
    float sensorHeight;              /* 2.4   */
    float focalLength;               /* 2.0   */
    float fStop;                     /* 4.0   */
    float focalDistance;             /* 100.0 */
    int32_t useImage;                /* false */
    int32_t lensModel;               /* ZOICB_RAYTRACED */
    int32_t kolbSamplingLUT;         /* true  */
    int32_t useDof;                  /* true  */
    float opticalVignettingDistance; /* 0.0   */
    float opticalVignettingRadius;   /* 1.0   */
    float exposureControl;           /* 0.0   */
    const char* lensDataPath;        /* ""  : tabular lens description (lenses_tabular/ *.dat grammar) */
    const char* bokehPath;           /* ""  : label only; pixels are handed to zoicb_create directly  */
} zoicb_params;

/* Counters (64-bit, exact).  success / vignetted / total_internal_reflection are the reference's
 * succesRays / vignettedRays / totalInternalReflection (src/zoic.cpp:533-534); attempts and
 * element_visits feed the roofline flop count; exact_reruns counts rays the GUARDED mode re-ran. */
typedef struct zoicb_stats {
    uint64_t rays;
    uint64_t success;
    uint64_t vignetted;
    uint64_t total_internal_reflection;
    uint64_t attempts;
    uint64_t element_visits;
    uint64_t exact_reruns;
} zoicb_stats;

/* Derived camera state, for parity tests against the reference's setup (src/zoic.cpp:1575-1720). */
#define ZOICB_MAX_ELEMENTS 24
#define ZOICB_LUT_SIZE 32
typedef struct zoicb_constants {
    int32_t lensCount;
    int32_t apertureElement;
    int32_t lutSize;
    int32_t bokehWidth, bokehHeight;
    float fov, tan_fov, apertureRadius;                    /* thin lens, :1606-1608 */
    float userApertureRadius, originShift, apertureDistance, focalLengthRatio;
    float tracedFocalLength[2], principalPlane[2], focalPoint[2];
    float curvature[ZOICB_MAX_ELEMENTS], thickness[ZOICB_MAX_ELEMENTS], ior[ZOICB_MAX_ELEMENTS];
    float aperture[ZOICB_MAX_ELEMENTS], center[ZOICB_MAX_ELEMENTS];
    float lutKey[ZOICB_LUT_SIZE];
    float lutMinX[ZOICB_LUT_SIZE], lutMinY[ZOICB_LUT_SIZE], lutMaxX[ZOICB_LUT_SIZE], lutMaxY[ZOICB_LUT_SIZE];
    int32_t guardedSplit;   /* surfaces [0, split) / [split, lensCount): the two stages of the guarded kernel */
    int32_t guardedInnerRetry; /* 1: rays stopped in stage A re-sample inside the pass (high-rejection cameras) */
} zoicb_constants;

/* One generated camera ray: a 32-byte record (the fields of AtCameraOutput that zoic writes).
 *   origin, dir : camera space, looking down -Z, cm
 *   weight      : 0 when every retry was vignetted, else the exposure scale (reference :1951-1953, :1981-1987)
 *   tries       : number of re-samples, as a float; tries > 0 <=> the reference also sets dOdy = origin,
 *                 dDdy = dir (:1974-1977)
 * Arrays of rays must be 32-byte aligned (each record is written with one 256-bit store). */
typedef struct zoicb_ray {
    float origin[3];
    float weight;
    float dir[3];
    float tries;
} zoicb_ray;

typedef struct zoicb_ctx zoicb_ctx;

ZOICB_API void zoicb_default_params(zoicb_params* p);

/* Build a camera on CUDA device `device`: parse params->lensDataPath, run the reference's setup pipeline
 * bit-exactly on the host (exit-pupil LUT traced on the GPU), build the bokeh row/column CDF tables from
 * `rgb` (row-major, channel-interleaved, height x width x nch floats; may be NULL unless useImage), and
 * upload everything.  An image with 1 or 2 channels is accepted the way the reference accepts it: as an INVALID image
 * (imageData::isValid, src/zoic.cpp:135-137) whose every aperture sample is the lens centre (0, 0) (:420-425).
 * The camera state is immutable afterwards and generate calls may come from several host threads: calls on one context
 * serialise their (short) enqueue phase on an internal lock; kernels of different streams still overlap. */
ZOICB_API zoicb_status zoicb_create(const zoicb_params* params, const float* rgb, int width, int height, int nch,
                          int device, zoicb_ctx** out);
ZOICB_API void zoicb_destroy(zoicb_ctx* ctx);

ZOICB_API zoicb_status zoicb_set_mode(zoicb_ctx* ctx, int mode);
ZOICB_API int zoicb_get_mode(const zoicb_ctx* ctx);
/* Validation hook: multiply every decision margin of the GUARDED mode by `scale` (1 = shipped margins,
 * 0 = no margins: plain fast arithmetic whose path flips tests/ and tools/validate_guarded.py count to show
 * how much head-room the shipped margins have).  Not meant to be called while generate calls are in flight. */
ZOICB_API zoicb_status zoicb_set_guard_scale(zoicb_ctx* ctx, float scale);

/* camera_create_ray for a flat batch.  Both pointers are DEVICE pointers owned by the caller:
 *   d_samples  n x float4 (sx, sy, lensx, lensy)   -- the AtCameraInput fields zoic reads
 *   d_rays     n x zoicb_ray, 32-byte aligned
 * Initial AtCameraOutput state is origin = 0, weight = 1.  Retried samples draw from a per-sample
 * xorshift128 stream seeded from (rng_seed, first_index + i), so results do not
 * depend on batch boundaries, launch order or GPU count.  Asynchronous on `stream` (a cudaStream_t).
 * In GUARDED mode a zero-weight ray carries the film point and the optical axis as origin / dir (the
 * reference leaves the half-traced state of its last failed attempt there; EXACT mode reproduces that). */
ZOICB_API zoicb_status zoicb_generate(zoicb_ctx* ctx, const void* d_samples, uint64_t n, uint64_t first_index,
                                      uint64_t rng_seed, zoicb_ray* d_rays, void* stream);

/* The same with HOST buffers: pipelines host->device copies, kernels and device->host copies through
 * pinned staging chunks.  Synchronous. */
ZOICB_API zoicb_status zoicb_generate_host(zoicb_ctx* ctx, const float* h_samples, uint64_t n, uint64_t first_index,
                                           uint64_t rng_seed, zoicb_ray* h_rays);

/* zoicb_generate_host with PLANAR output: the same rays in 25 bytes instead of 32.  The host link carries every result,
 * and at 32 bytes per ray the download is what bounds the end-to-end rate -- but two of the
 * record's eight floats carry almost nothing: weight is 0 or the camera's exposure scale, tries is an integer <= 27.
 *   origin[k][i], dir[k][i] : component k of ray i's origin / dir, bit for bit (n floats per plane)
 *   flags[i]                : bits 0-6 = tries, bit 7 set <=> weight == 0                (n bytes)
 * *live_weight (optional) receives the weight of the rays whose bit 7 is clear.  Planes may be pageable or pinned
 * (pinned: copied directly); lossless: tests/test_gpu_parity.py rebuilds the 32-byte records from the planes. */
typedef struct zoicb_ray_planes {
    float* origin[3];
    float* dir[3];
    uint8_t* flags;
} zoicb_ray_planes;
ZOICB_API zoicb_status zoicb_generate_host_planar(zoicb_ctx* ctx, const float* h_samples, uint64_t n, uint64_t first_index,
                                                  uint64_t rng_seed, const zoicb_ray_planes* out, float* live_weight);

/* camera_create_ray for ONE sample (the shape of Arnold's per-sample callback): sample = (sx, sy, lensx,
 * lensy), one zoicb_ray in host memory.  Uses per-thread pinned staging and a per-thread
 * stream, so it may be called concurrently from many render threads; it is a launch + two tiny copies per
 * call, i.e. a compatibility path -- throughput comes from the batched entry points above. */
ZOICB_API zoicb_status zoicb_generate_one(zoicb_ctx* ctx, const float* sample, uint64_t sample_index,
                                          uint64_t rng_seed, zoicb_ray* ray);

/* Camera -> world epilogue: what the renderer does with every ray right after
 * camera_create_ray (the reference's output is in camera space, src/zoic.cpp:1845, :1960-1961).  m3x4 is a
 * row-major 3x4 camera-to-world matrix in HOST memory; origin' = M (origin, 1), dir' = M3x3 dir, weight and
 * tries pass through.  d_rays / d_out are device pointers (32-byte aligned; d_out may equal d_rays).
 * Arithmetic: one fma chain per component, innermost term first (zoic_b200/csrc/kernels.cu). */
ZOICB_API zoicb_status zoicb_transform_rays(zoicb_ctx* ctx, const zoicb_ray* d_rays, uint64_t n, const float* m3x4,
                                            zoicb_ray* d_out, void* stream);

/* Ray differentials: the AtCameraOutput fields the reference leaves unset (TODO at src/zoic.cpp:12-13)
 * and fakes with "if (tries > 0) dOdy = origin, dDdy = dir" (:1971-1977).  For every sample, the derivative of its ray
 * with respect to the screen position at a FIXED aperture point, as forward differences over one pixel: dsx / dsy are
 * the screen-space extent of a pixel (AtCameraInput::dsx / dsy).  d_samples, first_index and rng_seed are those of the
 * zoicb_generate call that produced d_rays (its weight and tries select the accepted attempt).  Zero-weight rays and
 * neighbours that are stopped inside the lens get zero vectors.  Exact arithmetic; the full contract is stated in
 * zoic_b200/csrc/differentials.cu and, for the CPU, in oracle/zoic_port.cpp (zport_differentials). */
typedef struct zoicb_ray_diff {
    float dOdx[3], dOdy[3], dDdx[3], dDdy[3];
} zoicb_ray_diff;
ZOICB_API zoicb_status zoicb_differentials(zoicb_ctx* ctx, const void* d_samples, uint64_t n, uint64_t first_index,
                                           uint64_t rng_seed, float dsx, float dsy, const zoicb_ray* d_rays,
                                           zoicb_ray_diff* d_out, void* stream);

/* Camera -> world for the differentials: they are differences of points / of directions, so each of the four vectors is
 * multiplied by the 3x3 part of the row-major 3x4 matrix (host memory), with the fma chain of zoicb_transform_rays'
 * directions.  d_out may equal d_diffs. */
ZOICB_API zoicb_status zoicb_transform_differentials(zoicb_ctx* ctx, const zoicb_ray_diff* d_diffs, uint64_t n,
                                                     const float* m3x4, zoicb_ray_diff* d_out, void* stream);

/* draw.zoic writer: the file the reference's -D_DRAW build leaves for src/draw.py
 * (writeToFile, src/zoic.cpp:1240-1293, and the DRAW_ONLY blocks of traceThroughLensElements :1121-1128,
 * :1146-1153): "LENSMODEL{KOLB}", the lens cross-section header, then "RAYS{...}" with the (z, y) path of every
 * attempt of every given sample through the element stack, traced on the GPU with the exact arithmetic and the
 * draw build's conventions (film point x = 0, direction x = 0).  The reference draws one sample in 100 000; here
 * the caller picks them: h_samples is n x (sx, sy, lensx, lensy) in HOST memory; sample i has global index
 * h_indices[i], or first_index + i when h_indices is NULL (the index selects its retry stream).  Raytraced lens
 * model only; at most 65536 samples. */
ZOICB_API zoicb_status zoicb_write_draw_file(zoicb_ctx* ctx, const char* path, const float* h_samples, uint32_t n,
                                             const uint64_t* h_indices, uint64_t first_index, uint64_t rng_seed);

/* Wall time of zoicb_create's stages in milliseconds (each pointer may be NULL): the whole call, the exit-pupil LUT
 * (3.2 M candidate rays classified and folded into 32 bounding boxes on the GPU; reference src/zoic.cpp:1391-1452),
 * the image-based aperture tables (:222-417).  The reference spends 0.4-0.7 s in node_update on these. */
ZOICB_API zoicb_status zoicb_get_create_times(const zoicb_ctx* ctx, double* total_ms, double* lut_ms, double* bokeh_ms);

/* ---------------------------------------------------------------------------------------------------------------
 * Whole-frame jobs and the NVLink gather (section 8)
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct zoicb_gather zoicb_gather;

/* A job = the samples [first, first + count) of a synthetic W x H frame (zoicb_synth_samples with spp_per_pass samples
 * per pixel and pass; the pixel index wraps once per pass, so a frame of P passes has W*H*spp_per_pass*P samples),
 * generated tile by tile: synthesise on the device -> generate -> consume (a checksum kernel, the renderer's stand-in)
 * through rotating buffers, three streams deep.  Jobs larger than HBM run at full size.  One GPU: first = 0, count =
 * the frame.  G GPUs: rank r runs [r N/G, (r+1) N/G) -- whole passes, so every rank sees the whole film.
 *   census      1: every tile is generated a second time in EXACT mode and all records are compared on the device
 *   windows     records of n_windows index ranges [window_first[w], window_first[w] + window_count) (global sample
 *               indices; the parts inside [first, first+count) are copied) -> d_windows[w * window_count ...] (device)
 *   gather      NULL: every rank consumes its own tiles.  Otherwise the tiles of all ranks land in the consumer rank's
 *               round buffers and are consumed there; gather_counts[world] = every rank's `count` (same array on all
 *               ranks); `tile` is the gather's.  All ranks must call zoicb_run_job together.
 *   serial      1: one stream, no overlap of the stages (A/B and debugging) */
typedef struct zoicb_job {
    uint32_t W, H, spp_per_pass;
    int32_t census;
    uint64_t sample_seed, rng_seed;
    uint64_t first, count;
    uint64_t tile;                  /* samples per tile, 0 = 2^28 */
    float census_tol;               /* 0 = 1e-5 (north-star tolerance) */
    int32_t n_windows;
    const uint64_t* window_first;
    uint64_t window_count;
    zoicb_ray* d_windows;
    zoicb_gather* gather;
    const uint64_t* gather_counts;
    int32_t serial;
    int32_t reserved;
} zoicb_job;

typedef struct zoicb_job_result {
    uint64_t rays, tiles, launches;
    float device_ms;                /* CUDA-event time of the whole job (all stages, all streams joined)            */
    float generate_ms;              /* sum of the generate kernels' own event times on their stream                 */
    /* consumer totals (on the consumer rank of a gathered job: over the records of ALL ranks) */
    uint64_t checksum;              /* order-independent: sum over records of sum_j word_j * K_j mod 2^64           */
    uint64_t zero_weight, tries_sum, consumed;
    /* census (GUARDED against EXACT, every record) */
    uint64_t census_rays, census_flips, census_out_of_tol, census_live;
    float census_max_rel_origin, census_max_dir;
    zoicb_stats stats;              /* this job's counters (the context's running counters are not touched)         */
    zoicb_stats census_stats;       /* the EXACT pass's counters                                                     */
} zoicb_job_result;

ZOICB_API zoicb_status zoicb_run_job(zoicb_ctx* ctx, const zoicb_job* job, zoicb_job_result* result);

/* Compares n records of two device buffers (GUARDED output, EXACT output) like the census of zoicb_run_job and ADDS
 * the totals to result's census_* fields.  Synchronises `stream`. */
ZOICB_API zoicb_status zoicb_census(zoicb_ctx* ctx, const zoicb_ray* d_fast, const zoicb_ray* d_exact, uint64_t n, float tol,
                                    zoicb_job_result* result, void* stream);

/* Gather-to-consumer over NVLink, one process per GPU (zoic_b200/csrc/gather.cu).  Every rank contributes up to
 * tile_rays records per round; the consumer rank owns `slots` round buffers of world x tile_rays records.
 * Transports: FUSED -- the generate kernels store straight into the consumer's memory (CUDA IPC mapping; compute and
 * transfer are one kernel); PUSH -- local staging tile + copy-engine push on a second stream; NCCL -- local staging tile +
 * ncclSend / grouped ncclRecv (libnccl.so.2 is resolved at run time; none is needed for FUSED / PUSH).
 * Set-up: create on every rank -> export a blob -> exchange the blobs by any means (torch.distributed, MPI, files) ->
 * connect with all world blobs in rank order.  NCCL transport: instead (or as well) zoicb_gather_init_nccl with an id
 * made by zoicb_nccl_unique_id on one rank, or zoicb_gather_use_nccl_comm with a communicator the caller owns.
 * Failure: the ranks number their rounds without talking, so a gathered zoicb_run_job that returns an error on ANY rank
 * (a flag wait timed out after ZOICB_GATHER_TIMEOUT_S seconds, 10 by default; a CUDA error) leaves the gather out of
 * step -- destroy it on every rank and build a new one; the error flag is sticky so later jobs fail fast, not hang. */
enum { ZOICB_GATHER_FUSED = 1, ZOICB_GATHER_PUSH = 2, ZOICB_GATHER_NCCL = 3 };
#define ZOICB_GATHER_BLOB_BYTES 192
#define ZOICB_NCCL_ID_BYTES 128
ZOICB_API zoicb_status zoicb_gather_create(int device, int rank, int world, int consumer, uint64_t tile_rays, int slots,
                                           int transport, zoicb_gather** out);
ZOICB_API zoicb_status zoicb_gather_export(zoicb_gather* g, void* blob /* ZOICB_GATHER_BLOB_BYTES */);
ZOICB_API zoicb_status zoicb_gather_connect(zoicb_gather* g, const void* blobs /* world x ZOICB_GATHER_BLOB_BYTES */);
ZOICB_API zoicb_status zoicb_nccl_unique_id(void* id /* ZOICB_NCCL_ID_BYTES */);
ZOICB_API zoicb_status zoicb_gather_init_nccl(zoicb_gather* g, const void* id /* collective over all ranks */);
ZOICB_API zoicb_status zoicb_gather_use_nccl_comm(zoicb_gather* g, void* nccl_comm /* an ncclComm_t */);
/* Consumer rank only: copies n records of `rank`'s segment of round `round` of the LAST job (offset records in) to host
 * memory -- valid while the round's slot has not been recycled (slots >= rounds keeps the whole job). */
ZOICB_API zoicb_status zoicb_gather_read(zoicb_gather* g, uint64_t round, int rank, uint64_t offset, uint64_t n, zoicb_ray* h_out);
ZOICB_API void zoicb_gather_destroy(zoicb_gather* g);

/* Synthetic camera samples for benchmarks and parity tests: sample index i is
 * pixel-major / spp-minor over a W x H image, four 24-bit uniforms from a counter hash of (seed, i). */
ZOICB_API zoicb_status zoicb_synth_samples(zoicb_ctx* ctx, uint32_t W, uint32_t H, uint32_t spp, uint64_t seed,
                                 uint64_t first_index, uint64_t n, void* d_samples, void* stream);

/* Counters accumulate over generate calls; get synchronises the context's device. */
ZOICB_API zoicb_status zoicb_get_stats(zoicb_ctx* ctx, zoicb_stats* out);
ZOICB_API zoicb_status zoicb_reset_stats(zoicb_ctx* ctx);

ZOICB_API zoicb_status zoicb_get_constants(const zoicb_ctx* ctx, zoicb_constants* out);
/* Bokeh tables (each may be NULL): cdfRow[h], rowIndices[h], cdfColumn[w*h], columnIndices[w*h]. */
ZOICB_API zoicb_status zoicb_get_bokeh_tables(const zoicb_ctx* ctx, float* cdfRow, int32_t* rowIndices,
                                    float* cdfColumn, int32_t* columnIndices);

/* Host-only setup: runs the whole creation pipeline WITHOUT a device (exit-pupil LUT candidates are
 * classified by host threads instead of the GPU kernel) and returns the derived constants and, optionally,
 * the bokeh tables.  For parity tests of the host logic on machines without a GPU; it cannot generate rays. */
ZOICB_API zoicb_status zoicb_setup_host_only(const zoicb_params* params, const float* rgb, int width, int height, int nch,
                                   zoicb_constants* out, float* cdfRow, int32_t* rowIndices, float* cdfColumn,
                                   int32_t* columnIndices);

/* Builds the image-based aperture tables of an image on the GPU (the kernels zoicb_create runs for a camera with
 * useImage; reference imageData::bokehProbability, src/zoic.cpp:222-417) and returns them: cdfRow[h],
 * rowIndices[h], cdfColumn[w*h], columnIndices[w*h] (each may be NULL).  Entry for entry the tables the reference
 * builds, tie order of its std::sort calls included.  `ms` (may be NULL) receives the device time of the build. */
ZOICB_API zoicb_status zoicb_build_bokeh_tables(int device, const float* rgb, int width, int height, int nch,
                                      float* cdfRow, int32_t* rowIndices, float* cdfColumn, int32_t* columnIndices,
                                      float* ms);

/* Test hook, host only: orders idx = 0..n-1 by descending values[idx] twice -- with the restatement of libstdc++'s
 * std::sort the device code uses (csrc/gnu_sort.h) into `restated`, and with the toolchain's own std::sort and the
 * reference's comparator shape (src/zoic.cpp:317) into `library` -- so a test can show the two agree, ties included. */
ZOICB_API zoicb_status zoicb_debug_sort_orders(const float* values, int32_t n, int32_t* restated, int32_t* library);

/* Test hook: folds candidate points into the exit-pupil LUT's bounding boxes the way the reference does (in order, with
 * the re-arm quirk of src/zoic.cpp:1423) -- on the GPU (lut_bbox_kernel, what zoicb_create runs) into boxes_device and
 * with the host statement into boxes_host (either may be NULL).  draws: n_film x per_film pairs of xor128 outputs,
 * accept: n_film x per_film flags, boxes: n_film x (min.x, min.y, max.x, max.y). */
ZOICB_API zoicb_status zoicb_debug_lut_boxes(int device, const uint32_t* draws, const uint8_t* accept, int32_t n_film,
                                             int32_t per_film, float first_aperture, float* boxes_device, float* boxes_host);

/* Test hook: the float the thin-lens kernels compare qx^2 + qy^2 with instead of taking the root (the smallest s >= 0 whose
 * correctly rounded square root is >= radius; zoic_b200/csrc/camera_state.h: ov_s_threshold).  Host only. */
ZOICB_API float zoicb_debug_sqrt_threshold(float radius);

/* Test hook: the thin-lens kernel's merged 1 / sqrt(x) (one range check around the fast paths of the IEEE root and the
 * IEEE reciprocal, zoic_b200/csrc/lens_math.cuh: normalize_factor) against the two library operations, for EVERY float
 * bit pattern x, on the device.  *mismatches = number of x whose results differ in any bit (NaN = NaN), *first_bad = the
 * smallest such pattern (0xFFFFFFFF when there is none). */
ZOICB_API zoicb_status zoicb_debug_check_normalize_factor(int device, uint64_t* mismatches, uint32_t* first_bad);

/* Measured fp32 FMA throughput of the device (dependent-chain-free FFMA kernel), in TFLOP/s: the
 * denominator of the fp32 roofline that bench.py reports. */
ZOICB_API zoicb_status zoicb_measure_fp32_peak(int device, double* tflops);

/* Number of kernels this library has launched in the calling process (all contexts). */
ZOICB_API uint64_t zoicb_kernel_launches(void);

ZOICB_API const char* zoicb_last_error(void);
ZOICB_API const char* zoicb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ZOICB_H */
