// host_setup.cpp -- camera creation on the host.
//
// Re-derives, bit for bit, every constant the reference's node_update computes (reference
// src/zoic.cpp:1575-1720) and packs it into the CameraState the kernels consume.  These constants must
// never come from a fused/fast-math path: the setup traces go through the R = 9999.9 cm aperture
// "sphere", where two ~5000 cm numbers cancel, so one ulp there moves every ray origin by more than the
// parity tolerance.  Build flags: -ffp-contract=off, no -march, no fast-math.
#include "host_setup.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <thread>

#include "lens_math.cuh"

namespace zoicb {
namespace {

// ------------------------------------------------------------------ lens table (src/zoic.cpp:708-914)
// Grammar: lines that are empty or start with '#' are skipped; fields are separated by any of
// tab , ; : space; the column count is int(total fields / data lines) and must be 4 (R, thickness, ior,
// aperture) or 5 (R, thickness, ior, V-number, aperture).  The reference assigns fields through a
// column counter that (a) advances on every delimiter, including the empty field between two adjacent
// delimiters, and (b) is not reset at line ends; both are reproduced.
const char kDelims[] = "\t,;: ";

zoicb_status read_lens_table(const std::string& path, std::vector<LensRow>* rows, std::string* err) {
    std::ifstream in(path);
    if (!in.good()) { *err = "cannot open lens file '" + path + "'"; return ZOICB_ERR_LENS_FILE; }
    std::vector<std::string> data_lines;
    std::string line;
    while (std::getline(in, line)) {
        if (line.empty() || line[0] == '#') continue;
        data_lines.push_back(line);
    }
    if (data_lines.empty()) { *err = "lens file '" + path + "' holds no data lines"; return ZOICB_ERR_LENS_FILE; }
    long fields = 0;
    for (const std::string& l : data_lines) {
        size_t start = 0, pos;
        while ((pos = l.find_first_of(kDelims, start)) != std::string::npos) {
            if (pos > start) ++fields;
            start = pos + 1;
        }
        if (start < l.size()) ++fields;
    }
    const int ncol = (int)((float)fields / (float)data_lines.size());
    if (ncol < 4) { *err = "lens file has fewer than 4 columns"; return ZOICB_ERR_LENS_FILE; }
    if (ncol > 5) { *err = "lens file has more than 5 columns"; return ZOICB_ERR_LENS_FILE; }

    LensRow cur;
    std::memset(&cur, 0, sizeof cur);
    int col = 0;
    bool bad_number = false;
    auto store = [&](const std::string& field) {
        if (col < 0 || col >= ncol) return;
        float v = 0.0f;
        try { v = std::stof(field); } catch (...) { bad_number = true; return; }
        // column -> member; with 5 columns the 4th is the (unused) V-number
        switch (col) {
            case 0: cur.curvature = v; break;
            case 1: cur.thickness = v; break;
            case 2: cur.ior = v; break;
            case 3: if (ncol == 4) cur.aperture = v; else cur.abbe = v; break;
            case 4: cur.aperture = v; break;
        }
        if (col == ncol - 1) col = -1;
    };
    for (const std::string& l : data_lines) {
        size_t start = 0, pos;
        while ((pos = l.find_first_of(kDelims, start)) != std::string::npos) {
            if (pos > start) store(l.substr(start, pos - start));
            start = pos + 1;
            ++col;
        }
        if (start < l.size()) { store(l.substr(start)); ++col; }
        rows->push_back(cur);
    }
    if (bad_number) { *err = "lens file holds a field that is not a number"; return ZOICB_ERR_LENS_FILE; }
    std::reverse(rows->begin(), rows->end());  // rear element first (src/zoic.cpp:913)
    return ZOICB_OK;
}

// ------------------------------------------------------------------ setup-time ray tools
// (the hot path uses exact_surface(); the setup traces need the `reverse` root, an independent sign for
// the normal and "virtual" intersections that ignore misses: src/zoic.cpp:973-1025)
Vec3 sphere_hit(Vec3 o, Vec3 d, float center_z, float R, bool reverse) {
    Vec3 u = vnormalize(d);
    Vec3 L = vmake(xsub(0.0f, o.x), xsub(0.0f, o.y), xsub(center_z, o.z));
    float tca = vdot(L, u);
    float r2 = xmul(R, R);
    float d2 = xsub(vdot(L, L), xmul(tca, tca));
    float thc = xsqrt(fabsf(xsub(r2, d2)));
    float sgn = R < 0.0f ? -1.0f : 1.0f;
    float t = reverse ? xsub(tca, xmul(thc, sgn)) : xadd(tca, xmul(thc, sgn));
    return vadd(o, vscale(u, t));
}
Vec3 sphere_normal(Vec3 hit, float center_z, float R_for_sign) {
    float sgn = R_for_sign < 0.0f ? -1.0f : 1.0f;
    return vscale(vnormalize(vmake(xsub(0.0f, hit.x), xsub(0.0f, hit.y), xsub(center_z, hit.z))), sgn);
}
// returns false (direction untouched) on total internal reflection when `real`
bool refract(Vec3* dir, float ior1, float ior2, Vec3 normal, bool real) {
    Vec3 i = vnormalize(*dir);
    Vec3 n = vnormalize(normal);
    float eta = (ior2 == 1.0) ? ior1 : xdiv(ior1, ior2);
    float c1 = -vdot(i, n);
    float cs2 = (float)((double)xmul(eta, eta) * (1.0 - (double)xmul(c1, c1)));
    if (real && ior1 > ior2 && cs2 > 1.0) return false;
    float k = (float)((double)xmul(eta, c1) - std::sqrt(std::fabs(1.0 - (double)cs2)));
    *dir = vadd(vscale(i, eta), vscale(n, k));
    return true;
}
// intersection of a ray with the plane y = 0, written the way the reference does (src/zoic.cpp:1043-1049):
// the "point on the plane" is normalize((100, 0, 100)) and the quotient is a multiply by the reciprocal
Vec3 hit_plane_y0(Vec3 o, Vec3 d) {
    Vec3 coord = vnormalize(vmake(100.0f, 0.0f, 100.0f));
    Vec3 nrm = vmake(0.0f, 1.0f, 0.0f);
    Vec3 u = vnormalize(d);
    float num = xsub(vdot(coord, nrm), vdot(nrm, o));
    float inv = xrcp(vdot(nrm, u));
    Vec3 s = vscale(u, num);
    return vadd(o, vmake(xmul(s.x, inv), xmul(s.y, inv), xmul(s.z, inv)));
}
// z of the intersection of two lines in the (z, y) plane, each given by two points (src/zoic.cpp:1029-1039)
float line_line_z(Vec3 p1, Vec3 p2, Vec3 q1, Vec3 q2) {
    float A1 = xsub(p2.y, p1.y), B1 = xsub(p1.z, p2.z);
    float C1 = xadd(xmul(A1, p1.z), xmul(B1, p1.y));
    float A2 = xsub(q2.y, q1.y), B2 = xsub(q1.z, q2.z);
    float C2 = xadd(xmul(A2, q1.z), xmul(B2, q1.y));
    float delta = xsub(xmul(A1, B2), xmul(A2, B1));
    return xdiv(xsub(xmul(B2, C1), xmul(B1, C2)), delta);
}

// ------------------------------------------------------------------ clean-up (src/zoic.cpp:917-959)
zoicb_status clean_rows(std::vector<LensRow>& rows, int* stop_index, std::string* err) {
    int stops = 0;
    *stop_index = 0;  // ruling: a table without a zero-radius row leaves the stop at element 0
    for (size_t i = 0; i < rows.size(); ++i) {
        if (rows[i].curvature == 0.0f) {
            *stop_index = (int)i;
            if (++stops > 1) { *err = "multiple aperture stops in lens file"; return ZOICB_ERR_LENS_DATA; }
            rows[i].curvature = 99999.0f;
        }
        if (rows[i].ior == 0.0f) rows[i].ior = 1.0f;
    }
    for (LensRow& r : rows) {  // mm -> cm: a double multiply by 0.1, narrowed
        r.curvature = (float)((double)r.curvature * 0.1);
        r.thickness = (float)((double)r.thickness * 0.1);
        r.aperture = (float)((double)r.aperture * 0.1);
    }
    float total = 0.0f;
    for (const LensRow& r : rows) total = xadd(total, r.thickness);
    rows[0].thickness = xsub(rows[0].thickness, total);  // front vertex ends up at z = 0
    return ZOICB_OK;
}

// ------------------------------------------------------------------ paraxial trace (src/zoic.cpp:1161-1228)
struct Paraxial { float principal_plane, focal_point, focal_length; };
Paraxial trace_focal_length(const std::vector<LensRow>& rows) {
    const int n = (int)rows.size();
    const float height = (float)((double)rows[0].aperture * 0.1);
    Vec3 o = vmake(0.0f, height, 0.0f);
    Vec3 d = vmake(0.0f, 0.0f, 99999.0f);
    Vec3 hit = vmake(0, 0, 0);
    float z = 0.0f;
    Paraxial out = {0, 0, 0};
    for (int i = 0; i < n; ++i) {
        z = (i == 0) ? rows[0].thickness : xadd(z, rows[i].thickness);
        float cz = xsub(z, rows[i].curvature);
        hit = sphere_hit(o, d, cz, rows[i].curvature, false);
        Vec3 nrm = sphere_normal(hit, cz, rows[i].curvature);
        refract(&d, rows[i].ior, (i != n - 1) ? rows[i + 1].ior : 1.0f, nrm, true);
        if (i == n - 1) {
            // both constructions start from the PREVIOUS hit (`o` is updated after this block), as the
            // reference does (:1191-1204)
            Vec3 a1 = vmake(0.0f, height, 0.0f), a2 = vmake(0.0f, height, 999999.0f);
            Vec3 b2 = vmake(0.0f, (float)((double)o.y + ((double)d.y * 100000.0)),
                            (float)((double)o.z + ((double)d.z * 100000.0)));
            out.principal_plane = line_line_z(a1, a2, o, b2);
            out.focal_point = hit_plane_y0(o, d).z;
        }
        o = hit;
    }
    out.focal_length = xsub(out.focal_point, out.principal_plane);
    return out;
}

// ------------------------------------------------------------------ focus (src/zoic.cpp:1054-1095)
float image_distance(const std::vector<LensRow>& rows, float object_distance) {
    const int n = (int)rows.size();
    Vec3 o = vmake(0.0f, 0.0f, object_distance);
    Vec3 d = vmake(0.0f, xmul(xdiv(rows[n - 1].aperture, 2.0f), 0.05f), -object_distance);
    float z = 0.0f;
    for (int k = 0; k < n; ++k) z = xadd(z, rows[k].thickness);
    float result = 0.0f;
    for (int i = 0; i < n; ++i) {  // front element first
        const int k = n - 1 - i;
        if (i != 0) z = xsub(z, rows[n - i].thickness);
        float cz = xsub(z, rows[k].curvature);
        Vec3 hit = sphere_hit(o, d, cz, rows[k].curvature, true);
        Vec3 nrm = sphere_normal(hit, cz, -rows[k].curvature);
        refract(&d, (i == 0) ? 1.0f : rows[n - i].ior, rows[k].ior, nrm, false);
        if (i == n - 1) result = hit_plane_y0(hit, d).z;
        o = hit;
    }
    return result;
}

// ------------------------------------------------------------------ kernel-facing element constants
void pack_elements(const std::vector<LensRow>& rows, int stop, float user_radius, float origin_shift, LensState* L) {
    const int n = (int)rows.size();
    L->count = n;
    L->aperture_element = stop;
    float z = 0.0f;
    for (int i = 0; i < n; ++i) {
        z = (i == 0) ? rows[0].thickness : xadd(z, rows[i].thickness);  // src/zoic.cpp:963-969
        Element& e = L->e[i];
        e.center = xsub(z, rows[i].curvature);
        e.radius = rows[i].curvature;
        e.radius2 = xmul(e.radius, e.radius);
        e.sgn = e.radius < 0.0f ? -1.0f : 1.0f;
        // h2 > (double)(aperture*0.5)^2  <=>  h2 > T with T the largest float <= that double
        double half = (double)rows[i].aperture * 0.5;
        double lim = half * half;
        float T = (float)lim;
        if ((double)T > lim) T = nextafterf(T, -INFINITY);
        if (i == stop) {
            float u2 = xmul(user_radius, user_radius);
            if (u2 < T) T = u2;
        }
        e.rim2 = T;
        const float next_ior = (i != n - 1) ? rows[i + 1].ior : 1.0f;
        e.eta = (next_ior == 1.0f) ? rows[i].ior : xdiv(rows[i].ior, next_ior);
        e.eta2 = xmul(e.eta, e.eta);
        e.inv_radius = 1.0f / e.radius;
        e.tir_possible = rows[i].ior > next_ior ? 1 : 0;
        // Fast-path constants (kernels.cu fast_surface()).  The fast path intersects the very sphere the
        // reference does -- centre fl(z - R), squared radius fl(R*R) -- but through a cancellation-free form.
        e.vertex = (float)((double)e.center + (double)e.radius);
        e.r2_corr = (float)((double)e.radius * (double)e.radius - (double)e.radius2);
        // Decision margins.  (1) The reference forms t = tca + thc from numbers of magnitude |o - c| <= |R| + len;
        // a rounding analysis of its operation sequence bounds the error of t by ~4.9 ulp of that magnitude
        // (2.4e-3 cm at the R = 4967 cm stop), all of it ALONG the ray, which moves hx^2+hy^2 by 2*(h.u_xy)*dt.
        // (2) Direction noise of ~2e-7 rad per surface displaces later hits by a few 1e-6 cm.
        const float len = fabsf(origin_shift) + 1.0f;
        e.dt_guard = 2.0f * (8.0f * 5.9604645e-8f * (fabsf(e.radius) + len));
        e.rim2_guard = 2.0f * sqrtf(T) * (4e-6f * fmaxf(len, 2.0f));
        e.miss_guard = 1e-5f * e.radius2;
        e.vertex_m2r = (float)((double)e.center - (double)e.radius);
        e.one_m_eta2 = (float)(1.0 - (double)e.eta2);
    }
}

// ------------------------------------------------------------------ exit-pupil LUT (src/zoic.cpp:1391-1452)
// 32 film positions x 100000 candidate rays drawn from ONE sequential xorshift128 stream with the
// reference's seed constants (ruling: every camera creation starts a fresh stream).  The draws are
// generated sequentially (cheap), the candidates classified in parallel (host threads here, or the GPU
// through `lut_fn`), and the order-dependent bounding-box update of the reference is replayed serially.
bool lut_trace_host(const LensState& lens, const float* film_x, int n_film, const uint32_t* draws, int per_film,
                    uint8_t* accept) {
    unsigned hw = std::thread::hardware_concurrency();
    int nthreads = (int)std::min<unsigned>(hw ? hw : 1u, (unsigned)n_film);
    auto work = [&](int t) {
        for (int f = t; f < n_film; f += nthreads) {
            for (int s = 0; s < per_film; ++s) {
                size_t idx = (size_t)f * per_film + s;
                float U = xsub(xmul(u32_to_unit(draws[2 * idx]), 2.0f), 1.0f);
                float V = xsub(xmul(u32_to_unit(draws[2 * idx + 1]), 2.0f), 1.0f);
                Ray r;
                r.o = vmake(film_x[f], 0.0f, lens.origin_shift);
                r.d = vmake(xsub(xmul(U, lens.first_aperture), r.o.x), xsub(xmul(V, lens.first_aperture), r.o.y),
                            lens.neg_first_thickness);
                int visited;
                accept[idx] = exact_march(lens, r, &visited) == kPass ? 1 : 0;
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    return true;
}

void build_lut(LensState* L, zoicb_constants* C, LutTraceFn fn, void* user) {
    const int n_film = kLutSize, per_film = 100000;
    const float spacing = 4.0f / (float)n_film;
    std::vector<float> film_x(n_film);
    for (int i = 0; i < n_film; ++i) film_x[i] = xmul(spacing, (float)i);
    std::vector<uint32_t> draws((size_t)2 * n_film * per_film);
    Xor128 rng = {123456789u, 362436069u, 521288629u, 88675123u};
    for (uint32_t& d : draws) d = xor128_next(rng);
    std::vector<uint8_t> accept((size_t)n_film * per_film);
    std::vector<float> boxes((size_t)n_film * 4);
    const int how = fn ? fn(user, *L, film_x.data(), n_film, draws.data(), per_film, accept.data(), boxes.data()) : 0;
    if (how == 0) lut_trace_host(*L, film_x.data(), n_film, draws.data(), per_film, accept.data());
    if (how != 2) lut_fold_boxes_host(draws.data(), accept.data(), n_film, per_film, L->first_aperture, boxes.data());
    for (int f = 0; f < n_film; ++f) {
        const float minx = boxes[4 * f], miny = boxes[4 * f + 1], maxx = boxes[4 * f + 2], maxy = boxes[4 * f + 3];
        C->lutKey[f] = film_x[f];
        C->lutMinX[f] = minx; C->lutMinY[f] = miny; C->lutMaxX[f] = maxx; C->lutMaxY[f] = maxy;
        // boundingBox2d::getCentroid / getMaxScale (src/zoic.cpp:495-517)
        float cx = xmul(xadd(minx, maxx), 0.5f), cy = xmul(xadd(miny, maxy), 0.5f);
        float ex = xsub(maxx, cx), ey = xsub(maxy, cy);
        float sx = xsqrt(xmul(ex, ex)), sy = xsqrt(xmul(ey, ey));
        L->lut_scale[f] = (sx >= sy) ? sx : sy;
        L->lut_cx[f] = cx;
    }
    L->lut_size = n_film;
    C->lutSize = n_film;
}

// ------------------------------------------------------------------ stage boundary of the guarded kernel
// The guarded kernel marches the surfaces in two stages with a warp-level compaction in between, so that rays
// stopped early (rear rim, stop) do not idle through the rest of the stack.  Where to cut depends on where
// this camera's attempts die: trace a few thousand attempts on the host (exact arithmetic, samples spread over
// the sensor), histogram the stopping surface and minimise a simple issue-slot model.
int choose_split(const LensState& L, float sensor_w, float sensor_h, int* inner_retry, int* pretest) {
    const int n = L.count;
    *inner_retry = 0;
    *pretest = 0;
    if (n < 2) return 1;
    std::vector<double> stop_at(n, 0.0);
    double pass = 0.0, total = 0.0;
    Xor128 rng = {0x9E3779B9u, 0x243F6A88u, 0xB7E15162u, 0x8AED2A6Bu};
    const float aspect = sensor_w > 0.0f ? sensor_h / sensor_w : 1.0f;
    for (int s = 0; s < 6000; ++s) {
        const float sx = 2.0f * u32_to_unit(xor128_next(rng)) - 1.0f;
        const float sy = (2.0f * u32_to_unit(xor128_next(rng)) - 1.0f) * aspect;
        const KolbSampleState k = L.use_lut ? kolb_sample_setup<true, true>(L, sx, sy) : kolb_sample_setup<false, true>(L, sx, sy);
        for (int a = 0; a < 4; ++a) {  // a few attempts per film point
            float lx, ly;
            concentric_disk(u32_to_unit(xor128_next(rng)), u32_to_unit(xor128_next(rng)), &lx, &ly);
            Ray r;
            r.o = vmake(k.fx, k.fy, L.origin_shift);
            r.d = L.use_lut ? kolb_aim<true>(L, k, lx, ly, a > 0) : kolb_aim<false>(L, k, lx, ly, a > 0);
            int visited = 0;
            const int rc = exact_march(L, r, &visited);
            total += 1.0;
            if (rc == kPass) pass += 1.0; else stop_at[visited - 1] += 1.0;
        }
    }
    if (std::getenv("ZOICB_DEBUG_SPLIT")) {   // calibration histogram: share of attempts stopped at each surface
        std::fprintf(stderr, "[zoicb] attempts %.0f pass %.4f stop_at:", total, pass / total);
        for (int i = 0; i < n; ++i) std::fprintf(stderr, " %d:%.4f", i, stop_at[i] / total);
        std::fprintf(stderr, "\n");
    }
    if (pass < 1.0) pass = 1.0;
    const double setup = 130.0, per_surface = 66.0;
    int best = 1;
    double best_cost = 1e300;
    for (int k = 1; k < n; ++k) {
        double dead_a = 0.0, dead_b = 0.0;
        for (int i = 0; i < n; ++i) (i < k ? dead_a : dead_b) += stop_at[i];
        const double survive_a = (total - dead_a) / total;
        const double util_b = survive_a > 0.0 ? 1.0 - 0.5 * (dead_b / total) / survive_a : 1.0;
        const double cost = (setup + k * per_surface) + survive_a * (n - k) * per_surface / util_b + 40.0;
        if (cost < best_cost) { best_cost = cost; best = k; }
    }
    // in-pass re-sampling pays when, on average, at least half a warp is stopped inside stage A
    double dead_a = 0.0;
    for (int i = 0; i < best; ++i) dead_a += stop_at[i];
    *inner_retry = dead_a / total > 0.5 ? 1 : 0;
    // rim pre-test flavour of the packed kernel: wins on every camera whose attempts mostly die inside stage A
    // (+5 % on the fisheye, whose attempts die on surfaces 0-5, to +12 % on the narrow-field lenses, where the first
    // surface alone stops ~95 % of them; profiles/r01_ab_pool2.txt)
    *pretest = *inner_retry;
    return best;
}

// ------------------------------------------------------------------ bokeh tables (src/zoic.cpp:222-417)
// Luminance -> normalised PDF -> rows sorted by descending mass -> per-row columns sorted by descending
// conditional probability -> running sums.  Sums are sequential fp32 and the sorts are std::sort with a
// "greater by value" index comparator: tie order among equal values (the zero pixels outside the aperture
// shape) is whatever libstdc++'s introsort yields for this exact call shape, as for the reference.
struct ByValueDesc {
    const float* v;
    bool operator()(int a, int b) const { return v[a] > v[b]; }
};

zoicb_status check_bokeh_image_impl(const float* rgb, int w, int h, int nch, std::string* err) {
    if (!rgb || w <= 0 || h <= 0 || nch < 1 || (long long)w * h > (1ll << 26)) {
        *err = "useImage is set but there are no usable pixels";   // the reference: "Couldn't open bokeh image!" + abort (:1587-1591)
        return ZOICB_ERR_BOKEH_IMAGE;
    }
    if (w > 65535) { *err = "bokeh image wider than 65535 pixels"; return ZOICB_ERR_UNSUPPORTED; }
    if (h > kMaxBokehRows) { *err = "bokeh image has more rows than the kernels stage in shared memory (5120)"; return ZOICB_ERR_UNSUPPORTED; }
    return ZOICB_OK;
}

// Host statement of the table build: what zoicb_setup_host_only runs (no device), and the yardstick the tables
// built on the GPU (bokeh_build.cu) are tested against.
zoicb_status build_bokeh(const float* rgb, int w, int h, int nch, HostBokeh* out, std::string* err) {
    zoicb_status ok = check_bokeh_image_impl(rgb, w, h, nch, err);
    if (ok != ZOICB_OK) return ok;
    const int row_shift = out->row_shift, col_shift = out->col_shift;
    const int np = w * h;
    std::vector<float> lum(np), pdf(np), row_mass(h), cond(np);
    float total = 0.0f;
    for (int i = 0; i < np; ++i) {
        const float* px = rgb + (size_t)i * nch;
        lum[i] = xadd(xadd(xmul(px[0], 0.3f), xmul(px[1], 0.59f)), xmul(px[2], 0.11f));
        total = xadd(total, lum[i]);
    }
    const float inv_total = xdiv(1.0f, total);
    for (int i = 0; i < np; ++i) pdf[i] = xmul(lum[i], inv_total);
    for (int r = 0; r < h; ++r) {
        float acc = 0.0f;
        for (int c = 0; c < w; ++c) acc = xadd(acc, pdf[r * w + c]);
        row_mass[r] = acc;
    }
    out->w = w; out->h = h;
    out->row_indices.resize(h);
    for (int r = 0; r < h; ++r) out->row_indices[r] = r;
    std::sort(out->row_indices.data(), out->row_indices.data() + h, ByValueDesc{row_mass.data()});
    out->cdf_row.resize(h);
    float run = 0.0f;
    for (int r = 0; r < h; ++r) { run = xadd(run, row_mass[out->row_indices[r]]); out->cdf_row[r] = run; }
    for (int r = 0; r < h; ++r)
        for (int c = 0; c < w; ++c) {
            int i = r * w + c;
            cond[i] = (pdf[i] != 0 && row_mass[r] != 0) ? xdiv(pdf[i], row_mass[r]) : 0.0f;
        }
    out->column_indices.resize(np);
    for (int i = 0; i < np; ++i) out->column_indices[i] = i;
    for (int r = 0; r < h; ++r)
        std::sort(out->column_indices.data() + r * w, out->column_indices.data() + r * w + w, ByValueDesc{cond.data()});
    out->cdf_column.resize(np);
    for (int r = 0; r < h; ++r) {
        run = 0.0f;
        for (int c = 0; c < w; ++c) {
            int i = r * w + c;
            run = xadd(run, cond[out->column_indices[i]]);
            out->cdf_column[i] = run;
        }
    }
    // guide tables for the device-side searches (camera_state.h): they only narrow the range the search visits,
    // the result stays std::upper_bound's
    out->row_shift = row_shift > 0 ? row_shift : default_guide_shift(h);
    out->col_shift = col_shift > 0 ? col_shift : default_guide_shift(w);
    const size_t gr = ((size_t)1 << out->row_shift) + 2, gc = ((size_t)1 << out->col_shift) + 2;
    out->row_guide.resize(gr);
    build_guide_table(out->cdf_row.data(), h, out->row_shift, out->row_guide.data());
    out->col_guide.resize((size_t)h * gc);
    for (int r = 0; r < h; ++r) build_guide_table(out->cdf_column.data() + (size_t)r * w, w, out->col_shift, out->col_guide.data() + (size_t)r * gc);
    return ZOICB_OK;
}

}  // namespace

// The smallest float s >= 0 with fl(sqrt(s)) >= r (camera_state.h: ThinState::ov_s_threshold): bisection over the bit patterns
// of the non-negative floats, which are ordered like their values; the host's sqrtf is the correctly rounded root the
// device's __fsqrt_rn computes.
float sqrt_threshold(float r) {
    if (!(r > 0.0f)) return 0.0f;   // r <= 0 or NaN: no s has a root below r
    uint32_t lo = 0u, hi = 0x7F800000u;   // invariant: root(lo) < r (root(0) = 0), root(hi) >= r (root(+inf) = +inf)
    while (hi - lo > 1u) {
        const uint32_t mid = lo + (hi - lo) / 2u;
        float m; std::memcpy(&m, &mid, 4);
        if (sqrtf(m) >= r) hi = mid; else lo = mid;
    }
    float out; std::memcpy(&out, &hi, 4);
    return out;
}

void lut_fold_boxes_host(const uint32_t* draws, const uint8_t* accept, int n_film, int per_film, float ap, float* boxes) {
    for (int f = 0; f < n_film; ++f) {
        float minx = 0, miny = 0, maxx = 0, maxy = 0;
        for (int s = 0; s < per_film; ++s) {
            size_t idx = (size_t)f * per_film + s;
            if (!accept[idx]) continue;
            float px = xmul(xsub(xmul(u32_to_unit(draws[2 * idx]), 2.0f), 1.0f), ap);
            float py = xmul(xsub(xmul(u32_to_unit(draws[2 * idx + 1]), 2.0f), 1.0f), ap);
            if (xadd(minx, miny) == 0.0f) { minx = maxx = px; miny = maxy = py; }  // :1423, re-arms on an exact 0 sum
            if (px > maxx) maxx = px;
            if (py > maxy) maxy = py;
            if (px < minx) minx = px;
            if (py < miny) miny = py;
        }
        boxes[4 * f] = minx; boxes[4 * f + 1] = miny; boxes[4 * f + 2] = maxx; boxes[4 * f + 3] = maxy;
    }
}

zoicb_status check_bokeh_image(const float* rgb, int w, int h, int nch, std::string* err) {
    return check_bokeh_image_impl(rgb, w, h, nch, err);
}

void std_sort_desc(const float* values, int n, int32_t* idx) {
    for (int i = 0; i < n; ++i) idx[i] = i;
    std::sort(idx, idx + n, ByValueDesc{values});
}

// ------------------------------------------------------------------ node_update (src/zoic.cpp:1575-1720)
zoicb_status build_camera(const zoicb_params& p, const float* rgb, int w, int h, int nch, HostCamera* out,
                          std::string* err, LutTraceFn lut_fn, void* lut_user, BokehBuildFn bokeh_fn, void* bokeh_user) {
    std::memset(&out->state, 0, sizeof out->state);
    std::memset(&out->constants, 0, sizeof out->constants);
    out->params = p;
    out->lens_path = p.lensDataPath ? p.lensDataPath : "";
    out->params.lensDataPath = nullptr;
    out->params.bokehPath = nullptr;
    CameraState& S = out->state;
    zoicb_constants& C = out->constants;
    S.lens_model = p.lensModel;
    S.compact.col_guide8 = nullptr;   // filled in by the C-ABI for narrow images (camera_state.h: BokehCompact)
    S.compact.rel_column8 = nullptr;
    S.use_image = p.useImage ? 1 : 0;
    const float e2 = xmul(p.exposureControl, p.exposureControl);  // :1981-1987
    S.weight_scale = 1.0f;
    S.guard_scale = 1.0f;
    if (p.exposureControl > 0.0f) S.weight_scale = xadd(1.0f, e2);
    else if (p.exposureControl < 0.0f) S.weight_scale = xdiv(1.0f, xadd(1.0f, e2));

    if (p.useImage) {
        zoicb_status rc = check_bokeh_image(rgb, w, h, nch, err);
        if (rc != ZOICB_OK) return rc;
        // guide resolutions (camera_state.h); ZOICB_GUIDE_ROW_LOG2 / ZOICB_GUIDE_COL_LOG2 override them for A/B runs
        auto shift_from_env = [](const char* name, int n) {
            const char* v = std::getenv(name);
            const int m = v ? std::atoi(v) : 0;
            return (m >= 1 && m <= 16) ? m : default_guide_shift(n);
        };
        out->bokeh.row_shift = shift_from_env("ZOICB_GUIDE_ROW_LOG2", h);
        out->bokeh.col_shift = shift_from_env("ZOICB_GUIDE_COL_LOG2", 2 * w);   // two cells per column: +1 % over one (profiles/r02_ab.txt)
        if (nch < 3) {
            // imageData::isValid() is false for fewer than 3 channels (src/zoic.cpp:135-137): the reference builds no
            // tables and every bokehSample answers (0, 0) (:420-425).  Kept: the device gets a 1 x 1 stand-in (capi.cu).
            out->bokeh.w = w; out->bokeh.h = h; out->bokeh.degenerate = true;
            C.bokehWidth = w; C.bokehHeight = h;
        } else if (bokeh_fn) {
            if (!bokeh_fn(bokeh_user, rgb, w, h, nch, &out->bokeh)) {
                *err = "building the bokeh tables on the device failed";
                rc = ZOICB_ERR_CUDA;
            }
        } else {
            rc = build_bokeh(rgb, w, h, nch, &out->bokeh, err);
        }
        if (rc != ZOICB_OK) return rc;
        C.bokehWidth = w; C.bokehHeight = h;
    }

    if (p.lensModel == ZOICB_THINLENS) {  // :1598-1610
        ThinState& T = S.thin;
        C.fov = (float)(2.0f * atan((double)xdiv(p.sensorWidth, xmul(2.0f, p.focalLength))));
        C.tan_fov = tanf(xdiv(C.fov, 2.0f));
        C.apertureRadius = xdiv(p.focalLength, xmul(2.0f, p.fStop));
        T.tan_fov = C.tan_fov;
        T.aperture_radius = C.apertureRadius;
        T.focal_distance = p.focalDistance;
        T.ov_distance = p.opticalVignettingDistance;
        T.ov_radius_true = xmul(C.apertureRadius, p.opticalVignettingRadius);
        T.use_dof = p.useDof ? 1 : 0;
        T.use_ov = p.opticalVignettingDistance > 0.0f ? 1 : 0;
        T.ov_guard = 2e-5f * T.ov_radius_true;
        T.ov_s_threshold = sqrt_threshold(T.ov_radius_true);
        return ZOICB_OK;
    }
    if (p.lensModel != ZOICB_RAYTRACED) { *err = "lensModel must be THINLENS (0) or RAYTRACED (1)"; return ZOICB_ERR_INVALID_ARGUMENT; }
    if (out->lens_path.empty()) { *err = "lensDataPath is empty"; return ZOICB_ERR_LENS_FILE; }

    std::vector<LensRow>& rows = out->rows;
    rows.clear();
    zoicb_status rc = read_lens_table(out->lens_path, &rows, err);
    if (rc != ZOICB_OK) return rc;
    if ((int)rows.size() > kMaxElements) { *err = "lens file has more elements than ZOICB_MAX_ELEMENTS"; return ZOICB_ERR_LENS_DATA; }
    int stop = 0;
    rc = clean_rows(rows, &stop, err);
    if (rc != ZOICB_OK) return rc;

    Paraxial first = trace_focal_length(rows);                 // :1651
    const float ratio = xdiv(p.focalLength, first.focal_length);  // :1654
    for (LensRow& r : rows) {                                   // :1231-1237
        r.curvature = xmul(r.curvature, ratio);
        r.thickness = xmul(r.thickness, ratio);
        r.aperture = xmul(r.aperture, ratio);
    }
    Paraxial second = trace_focal_length(rows);                // :1661
    float user_radius = (float)((double)second.focal_length / (2.0 * (double)p.fStop));  // :1664
    if (user_radius > rows[stop].aperture) user_radius = rows[stop].aperture;            // :1668-1672 (diameter vs radius, kept)
    const float shift = image_distance(rows, p.focalDistance);  // :1675
    float stop_z = 0.0f;                                        // :1678-1685
    for (int i = 0; i <= stop; ++i) stop_z = xadd(stop_z, rows[i].thickness);

    LensState& L = S.lens;
    pack_elements(rows, stop, user_radius, shift, &L);
    L.origin_shift = shift;
    L.half_sensor = (float)((double)p.sensorWidth * 0.5);
    L.first_aperture = rows[0].aperture;
    L.neg_first_thickness = -rows[0].thickness;
    L.user_aperture_radius = user_radius;
    L.use_lut = p.kolbSamplingLUT ? 1 : 0;
    for (size_t i = 0; i < rows.size(); ++i) rows[i].center = L.e[i].center;

    C.lensCount = (int)rows.size();
    C.apertureElement = stop;
    C.userApertureRadius = user_radius;
    C.originShift = shift;
    C.apertureDistance = stop_z;
    C.focalLengthRatio = ratio;
    C.tracedFocalLength[0] = first.focal_length; C.tracedFocalLength[1] = second.focal_length;
    C.principalPlane[0] = first.principal_plane; C.principalPlane[1] = second.principal_plane;
    C.focalPoint[0] = first.focal_point; C.focalPoint[1] = second.focal_point;
    for (size_t i = 0; i < rows.size(); ++i) {
        C.curvature[i] = rows[i].curvature; C.thickness[i] = rows[i].thickness; C.ior[i] = rows[i].ior;
        C.aperture[i] = rows[i].aperture; C.center[i] = rows[i].center;
    }
    if (p.kolbSamplingLUT) build_lut(&L, &C, lut_fn, lut_user);  // :1691-1692
    L.split = choose_split(L, p.sensorWidth, p.sensorHeight, &L.inner_retry, &L.pretest);
    C.guardedSplit = L.split;
    C.guardedInnerRetry = L.inner_retry;
    return ZOICB_OK;
}

}  // namespace zoicb
