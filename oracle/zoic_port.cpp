// zoic_port.cpp -- TEST INFRASTRUCTURE (parity oracle), not part of the shipped product.
//
// A scalar CPU restatement of the reference camera's setup pipeline and camera_create_ray, written
// from the reference's behaviour (every function cites the reference file:line it follows; paths are
// relative to /root/reference).  It exists so that
//   * parity tests have an oracle on machines where the reference tree is absent (the GPU box),
//   * the exact number of attempts / lens-element visits per batch is known (roofline flop counts),
//   * the CPU baseline can be timed on all host cores (the reference itself is not thread-safe).
// PARITY PINNING: this restatement is checked bit-for-bit against the compiled, unmodified reference
// (oracle/_ref, built by oracle/Makefile) in tests/test_oracle_vs_reference.py, and against the one
// externally authored known-answer test the reference holds (src/draw.zoic:1-10) in
// tests/test_oracle_golden.py.  Golden vectors generated from the compiled reference are committed
// under tests/golden/ so the pin also holds where /root/reference does not exist.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// compile, load or call this file.  The product (zoic_b200/) never does.
//
// Arithmetic rules: strict IEEE fp32, no FMA contraction (build with -ffp-contract=off, no -march),
// double precision exactly where the reference's C++ promotes to double.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <thread>
#include <vector>

namespace {

// ------------------------------------------------------------------------------------------------
// vector helpers -- the arithmetic of include/arnold_shim/ai.h (Arnold inline math)
// ------------------------------------------------------------------------------------------------
struct V3 { float x, y, z; };
inline V3 v3(float x, float y, float z) { V3 r = {x, y, z}; return r; }
inline V3 add(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 sub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 mul(V3 a, float f) { return v3(a.x * f, a.y * f, a.z * f); }
inline V3 divs(V3 a, float f) { float c = 1 / f; return v3(a.x * c, a.y * c, a.z * c); }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 normalize(V3 a) {
    float len = sqrtf(a.x * a.x + a.y * a.y + a.z * a.z);
    if (len != 0) len = 1 / len;
    return mul(a, len);
}

const float PI_F = 3.14159265358979323846f;      // AI_PI
const float PIOVER2_F = 1.57079632679489661923f;  // AI_PIOVER2

// ------------------------------------------------------------------------------------------------
// RNG: src/zoic.cpp:647-652 (Marsaglia xorshift128), with explicit state
// ------------------------------------------------------------------------------------------------
struct Xor128 { uint32_t x, y, z, w; };
const Xor128 kCanonicalSeed = {123456789u, 362436069u, 521288629u, 88675123u};  // src/zoic.cpp:648
inline uint32_t xor128(Xor128& s) {
    uint32_t t = s.x ^ (s.x << 11);
    s.x = s.y; s.y = s.z; s.z = s.w;
    return s.w = (s.w ^ (s.w >> 19) ^ t ^ (t >> 8));
}

// repo conventions: SplitMix64 finaliser, per-sample retry stream, synthetic samples
inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
const uint64_t kGolden = 0x9E3779B97F4A7C15ull;
inline Xor128 sample_stream(uint64_t seed, uint64_t index) {
    uint64_t h0 = mix64(seed + kGolden * (index + 1));
    uint64_t h1 = mix64(h0 + kGolden);
    Xor128 s;
    s.x = (uint32_t)h0; s.y = (uint32_t)(h0 >> 32);
    s.z = (uint32_t)h1; s.w = (uint32_t)(h1 >> 32) | 1u;
    return s;
}

// ------------------------------------------------------------------------------------------------
// src/zoic.cpp:655-657
inline float linearInterpolate(float perc, float a, float b) { return a + perc * (b - a); }

// src/zoic.cpp:661-668.  `fmod` resolves to the C double function; `- AI_PI` is a double subtraction
// narrowed on assignment.
inline float fastSin(float x) {
    x = (float)(fmod((double)(x + PI_F), (double)(PI_F * 2)) - (double)PI_F);
    const float B = 4.0f / PI_F;
    const float C = -4.0f / (PI_F * PI_F);
    float y = B * x + C * x * std::fabs(x);
    const float P = 0.225f;
    return P * (y * std::fabs(y) - y) + y;
}

// src/zoic.cpp:671-681.  `x += AI_PI * 0.5` is a double addition narrowed on assignment.
inline float fastCos(float x) {
    x = (float)((double)x + (double)PI_F * 0.5);
    x = (float)(fmod((double)(x + PI_F), (double)(PI_F * 2)) - (double)PI_F);
    const float B = 4.0f / PI_F;
    const float C = -4.0f / (PI_F * PI_F);
    float y = B * x + C * x * std::fabs(x);
    const float P = 0.225f;
    return P * (y * std::fabs(y) - y) + y;
}

// src/zoic.cpp:686-704 (Shirley/Cline concentric map; a = b = 0 gives 0/0 = NaN, kept)
inline void concentricDiskSample(float ox, float oy, float* lx, float* ly) {
    float phi, r;
    float a = (float)(2.0 * (double)ox - 1.0);
    float b = (float)(2.0 * (double)oy - 1.0);
    if ((a * a) > (b * b)) {
        r = a;
        phi = (0.78539816339f) * (b / a);
    } else {
        r = b;
        phi = (PIOVER2_F) - (0.78539816339f) * (a / b);
    }
    *lx = r * fastCos(phi);
    *ly = r * fastSin(phi);
}

// ------------------------------------------------------------------------------------------------
// bokeh image tables: src/zoic.cpp:115-486
// ------------------------------------------------------------------------------------------------
struct IndexGreater {  // src/zoic.cpp:106-112
    const float* values;
    bool operator()(int l, int r) const { return values[l] > values[r]; }
};

struct BokehImage {
    int x = 0, y = 0, nchannels = 0;
    std::vector<float> pixelData, cdfRow, cdfColumn;
    std::vector<int> rowIndices, columnIndices;

    bool isValid() const { return (x * y * nchannels > 0 && nchannels >= 3); }  // :135-137

    // src/zoic.cpp:222-417.  Sequential float sums and std::sort with the same comparator shape, so the
    // tables (including the order of tied entries) come out as libstdc++ produces them for the reference.
    void build() {
        if (!isValid()) return;
        int npixels = x * y;
        int o1 = (nchannels >= 2 ? 1 : 0);
        int o2 = (nchannels >= 3 ? 2 : o1);
        std::vector<float> pixelValues(npixels), normalized(npixels), summedRow(y), perRow(npixels);
        float total = 0.0f;
        for (int i = 0, j = 0; i < npixels; ++i, j += nchannels) {  // :243-249
            pixelValues[i] = pixelData[j] * 0.3f + pixelData[j + o1] * 0.59f + pixelData[j + o2] * 0.11f;
            total += pixelValues[i];
        }
        float invTotal = 1.0f / total;                                // :259
        for (int i = 0; i < npixels; ++i) normalized[i] = pixelValues[i] * invTotal;  // :262-268
        for (int i = 0, k = 0; i < y; ++i) {                          // :283-293
            summedRow[i] = 0.0f;
            for (int j = 0; j < x; ++j, ++k) summedRow[i] += normalized[k];
        }
        rowIndices.resize(y);
        for (int i = 0; i < y; ++i) rowIndices[i] = i;
        std::sort(rowIndices.data(), rowIndices.data() + y, IndexGreater{summedRow.data()});  // :317
        cdfRow.resize(y);
        float prev = 0.0f;
        for (int i = 0; i < y; ++i) { cdfRow[i] = prev + summedRow[rowIndices[i]]; prev = cdfRow[i]; }  // :333-339
        for (int r = 0, i = 0; r < y; ++r)                           // :352-364
            for (int c = 0; c < x; ++c, ++i)
                perRow[i] = ((normalized[i] != 0) && (summedRow[r] != 0)) ? normalized[i] / summedRow[r] : 0;
        columnIndices.resize(npixels);
        for (int i = 0; i < npixels; i++) columnIndices[i] = i;
        for (int i = 0; i < npixels; i += x)                         // :380-382
            std::sort(columnIndices.data() + i, columnIndices.data() + i + x, IndexGreater{perRow.data()});
        cdfColumn.resize(npixels);
        for (int r = 0, i = 0; r < y; ++r) {                         // :398-407
            prev = 0.0f;
            for (int c = 0; c < x; ++c, ++i) { cdfColumn[i] = prev + perRow[columnIndices[i]]; prev = cdfColumn[i]; }
        }
    }

    // src/zoic.cpp:420-485
    void bokehSample(float randomNumberRow, float randomNumberColumn, float* dx, float* dy) const {
        if (!isValid()) { *dx = 0.0f; *dy = 0.0f; return; }
        const float* ub = std::upper_bound(cdfRow.data(), cdfRow.data() + y, randomNumberRow);   // :432
        int r = (ub >= cdfRow.data() + y) ? y - 1 : (int)(ub - cdfRow.data());                   // :435
        int actualPixelRow = rowIndices[r];
        int recalculatedPixelRow = actualPixelRow - ((x - 1) / 2);                               // :441 (x, not y)
        int startPixel = actualPixelRow * x;
        const float* ubc = std::upper_bound(cdfColumn.data() + startPixel, cdfColumn.data() + startPixel + x,
                                            randomNumberColumn);                                 // :458
        int c = (ubc >= cdfColumn.data() + startPixel + x) ? startPixel + x - 1 : (int)(ubc - cdfColumn.data());
        int actualPixelColumn = columnIndices[c];
        int relativePixelColumn = actualPixelColumn - startPixel;
        int recalculatedPixelColumn = relativePixelColumn - ((y - 1) / 2);                       // :466 (y, not x)
        float flippedRow = (float)recalculatedPixelColumn;
        float flippedColumn = recalculatedPixelRow * -1.0f;
        *dx = (float)((double)(flippedRow / (float)x) * 2.0);                                    // :483
        *dy = (float)((double)(flippedColumn / (float)y) * 2.0);                                 // :484
    }
};

// ------------------------------------------------------------------------------------------------
// lens data: src/zoic.cpp:522-541
// ------------------------------------------------------------------------------------------------
struct LensElement { float curvature, thickness, ior, aperture, abbe, center; };

struct BBox2 {  // src/zoic.cpp:490-518
    float maxx, maxy, minx, miny;
    float centroidX() const { return (minx + maxx) * 0.5f; }
    float centroidY() const { return (miny + maxy) * 0.5f; }
    float maxScale() const {
        float x1 = maxx - centroidX();
        float y2 = maxy - centroidY();
        float scaleX = std::sqrt(x1 * x1);
        float scaleY = std::sqrt(y2 * y2);
        return (scaleX >= scaleY) ? scaleX : scaleY;
    }
};

struct Lensdata {
    std::vector<LensElement> lenses;
    int lensCount = 0;
    float userApertureRadius = 0;
    int apertureElement = 0;  // ruling: files without a stop leave this 0
    float apertureDistance = 0, focalLengthRatio = 0, originShift = 0, focalDistance = 0;
    // exit-pupil LUT, keys i * 0.125 (src/zoic.cpp:1391-1452); std::map replaced by a dense table
    std::vector<float> lutKey;
    std::vector<BBox2> lutBox;
    // setup log values (the reference prints them with AiMsgInfo)
    float tracedFocalLength[2] = {0, 0}, principalPlane[2] = {0, 0}, focalPoint[2] = {0, 0};
};

struct Stats {
    uint64_t success = 0, vignetted = 0, attempts = 0, elementVisits = 0, tir = 0;
};

// src/zoic.cpp:708-914.  Two passes: count columns, then assign tokens cyclically.  The column counter
// advances on EVERY delimiter (also on empty tokens) and carries across lines, exactly as written.
int readTabularLensData(const std::string& path, Lensdata* ld) {
    std::ifstream f(path);
    if (!f.good()) return 1;
    std::string line;
    int columns = 0, lines = 0;
    const char* delims = "\t,;: ";
    while (getline(f, line)) {
        if (line.empty() || line.front() == '#') continue;
        std::size_t prev = 0, pos;
        while ((pos = line.find_first_of(delims, prev)) != std::string::npos) {
            if (pos > prev) ++columns;
            prev = pos + 1;
        }
        if (prev < line.length()) ++columns;
        ++lines;
    }
    if (lines == 0) return 2;
    f.clear();
    f.seekg(0, std::ios::beg);
    int totalColumns = (int)((float)columns / (float)lines);  // :741
    if (totalColumns < 4 || totalColumns > 5) return 3;        // :745-754 (reference: error + abort + carries on)
    LensElement lens;
    memset(&lens, 0, sizeof lens);
    int counter = 0;
    const int last = totalColumns - 1;
    auto assign = [&](const std::string& tok) -> bool {
        float v;
        try { v = std::stof(tok); } catch (...) { return false; }
        if (totalColumns == 4) {
            if (counter == 0) lens.curvature = v;
            else if (counter == 1) lens.thickness = v;
            else if (counter == 2) lens.ior = v;
            else if (counter == 3) lens.aperture = v;
        } else {
            if (counter == 0) lens.curvature = v;
            else if (counter == 1) lens.thickness = v;
            else if (counter == 2) lens.ior = v;
            else if (counter == 3) lens.abbe = v;
            else if (counter == 4) lens.aperture = v;
        }
        if (counter == last) counter = -1;
        return true;
    };
    while (getline(f, line)) {
        if (line.empty() || line.front() == '#') continue;
        std::size_t prev = 0, pos;
        while ((pos = line.find_first_of(delims, prev)) != std::string::npos) {
            if (pos > prev) {
                if (counter >= 0 && counter <= last) { if (!assign(line.substr(prev, pos - prev))) return 4; }
            }
            prev = pos + 1;
            ++counter;
        }
        if (prev < line.length()) {
            if (counter >= 0 && counter <= last) { if (!assign(line.substr(prev))) return 4; }
            ++counter;
        }
        ld->lenses.push_back(lens);
    }
    ld->lensCount = (int)ld->lenses.size();
    std::reverse(ld->lenses.begin(), ld->lenses.end());  // :913
    return 0;
}

// src/zoic.cpp:917-959
int cleanupLensData(Lensdata* ld) {
    int apertureCount = 0;
    for (int i = 0; i < ld->lensCount; i++) {
        if (ld->lenses[i].curvature == 0.0) {
            ld->apertureElement = i;
            if (++apertureCount > 1) return 5;  // reference: error + abort, then carries on
            ld->lenses[i].curvature = 99999.0;
        }
        if (ld->lenses[i].ior == 0.0) ld->lenses[i].ior = 1.0;
    }
    for (int i = 0; i < ld->lensCount; i++) {  // mm -> cm in double, narrowed on store
        ld->lenses[i].curvature = (float)((double)ld->lenses[i].curvature * 0.1);
        ld->lenses[i].thickness = (float)((double)ld->lenses[i].thickness * 0.1);
        ld->lenses[i].aperture = (float)((double)ld->lenses[i].aperture * 0.1);
    }
    float summedThickness = 0.0;
    for (int i = 0; i < ld->lensCount; i++) summedThickness += ld->lenses[i].thickness;
    ld->lenses[0].thickness -= summedThickness;
    return 0;
}

// src/zoic.cpp:963-969
void computeLensCenters(Lensdata* ld) {
    float summedThickness = 0;
    for (int i = 0; i < ld->lensCount; i++) {
        if (i == 0) summedThickness = ld->lenses[0].thickness; else summedThickness += ld->lenses[i].thickness;
        ld->lenses[i].center = summedThickness - ld->lenses[i].curvature;
    }
}

// src/zoic.cpp:973-995
inline bool raySphereIntersection(V3* hit, V3 dir, V3 origin, V3 center, float radius, bool reverse, bool real) {
    dir = normalize(dir);
    V3 L = sub(center, origin);
    float tca = dot(L, dir);
    float radius2 = radius * radius;
    float d2 = dot(L, L) - (tca * tca);
    if (real && (d2 > radius2)) return false;
    float thc = std::sqrt(std::fabs(radius2 - d2));
    float sign = (radius < 0.0f ? -1.0f : 1.0f);
    if (reverse) *hit = add(origin, mul(dir, (tca - thc * sign)));
    else *hit = add(origin, mul(dir, (tca + thc * sign)));
    return true;
}

// src/zoic.cpp:999-1004
inline void intersectionNormal(V3 hit, V3 center, float radius, V3* n) {
    float sign = (radius < 0.0f ? -1.0f : 1.0f);
    *n = mul(normalize(sub(center, hit)), sign);
}

// src/zoic.cpp:1008-1025.  cs2 and the square root are evaluated in double and narrowed, as written.
inline bool calculateTransmissionVector(V3* out, float ior1, float ior2, V3 incident, V3 normal, bool real) {
    incident = normalize(incident);
    normal = normalize(normal);
    float eta;
    if (ior2 == 1.0) eta = ior1; else eta = ior1 / ior2;
    float c1 = -dot(incident, normal);
    float cs2 = (float)((double)(eta * eta) * (1.0 - (double)(c1 * c1)));
    if (real && (ior1 > ior2) && (cs2 > 1.0)) return false;
    float k = (float)((double)(eta * c1) - std::sqrt(std::fabs(1.0 - (double)cs2)));
    *out = add(mul(incident, eta), mul(normal, k));
    return true;
}

// src/zoic.cpp:1029-1039 (returns .x only: the principal-plane z)
inline float lineLineIntersectionX(V3 l1o, V3 l1d, V3 l2o, V3 l2d) {
    float A1 = l1d.y - l1o.y;
    float B1 = l1o.z - l1d.z;
    float C1 = A1 * l1o.z + B1 * l1o.y;
    float A2 = l2d.y - l2o.y;
    float B2 = l2o.z - l2d.z;
    float C2 = A2 * l2o.z + B2 * l2o.y;
    float delta = A1 * B2 - A2 * B1;
    return (B2 * C1 - B1 * C2) / delta;
}

// src/zoic.cpp:1043-1049
inline V3 linePlaneIntersection(V3 rayOrigin, V3 rayDirection) {
    V3 coord = v3(100.0, 0.0, 100.0);
    V3 planeNormal = v3(0.0, 1.0, 0.0);
    rayDirection = normalize(rayDirection);
    coord = normalize(coord);
    return add(rayOrigin, divs(mul(rayDirection, (dot(coord, planeNormal) - dot(planeNormal, rayOrigin))),
                               dot(planeNormal, rayDirection)));
}

// src/zoic.cpp:1054-1095
float calculateImageDistance(float objectDistance, Lensdata* ld) {
    const int n = ld->lensCount;
    V3 ray_origin = v3(0.0f, 0.0f, objectDistance);
    V3 ray_direction = v3(0.0f, (ld->lenses[n - 1].aperture / 2.0f) * 0.05f, -objectDistance);
    float summedThickness = 0.0, imageDistance = 0.0;
    V3 hit_point_normal, hit_point = v3(0, 0, 0);
    for (int k = 0; k < n; k++) summedThickness += ld->lenses[k].thickness;
    for (int i = 0; i < n; i++) {
        if (i != 0) summedThickness -= ld->lenses[n - i].thickness;
        V3 sphere_center = v3(0.0f, 0.0f, summedThickness - ld->lenses[n - 1 - i].curvature);
        raySphereIntersection(&hit_point, ray_direction, ray_origin, sphere_center, ld->lenses[n - 1 - i].curvature, true, false);
        intersectionNormal(hit_point, sphere_center, -ld->lenses[n - 1 - i].curvature, &hit_point_normal);
        if (i == 0) calculateTransmissionVector(&ray_direction, 1.0, ld->lenses[n - i - 1].ior, ray_direction, hit_point_normal, false);
        else calculateTransmissionVector(&ray_direction, ld->lenses[n - i].ior, ld->lenses[n - i - 1].ior, ray_direction, hit_point_normal, false);
        if (i == n - 1) imageDistance = linePlaneIntersection(hit_point, ray_direction).z;
        ray_origin = hit_point;
    }
    return imageDistance;
}

// src/zoic.cpp:1099-1158 (and its by-value twin :1309-1350)
inline bool traceThroughLensElements(V3* ray_origin, V3* ray_direction, const Lensdata* ld, Stats* st) {
    V3 hit_point, hit_point_normal, sphere_center;
    const int n = ld->lensCount;
    for (int i = 0; i < n; i++) {
        const LensElement& e = ld->lenses[i];
        sphere_center = v3(0.0f, 0.0f, e.center);
        if (st) ++st->elementVisits;
        if (!raySphereIntersection(&hit_point, *ray_direction, *ray_origin, sphere_center, e.curvature, false, true)) return false;
        float hitPoint2 = hit_point.x * hit_point.x + hit_point.y * hit_point.y;
        if (((double)hitPoint2 > ((double)e.aperture * 0.5) * ((double)e.aperture * 0.5))      // :1114
            || ((i == ld->apertureElement) && (hitPoint2 > (ld->userApertureRadius * ld->userApertureRadius)))) return false;
        intersectionNormal(hit_point, sphere_center, e.curvature, &hit_point_normal);
        *ray_origin = hit_point;
        float ior2 = (i != n - 1) ? ld->lenses[i + 1].ior : 1.0f;
        if (!calculateTransmissionVector(ray_direction, e.ior, ior2, *ray_direction, hit_point_normal, true)) {
            if (st) ++st->tir;
            return false;
        }
    }
    return true;
}

// src/zoic.cpp:1161-1228
float traceThroughLensElementsForFocalLength(Lensdata* ld, int pass) {
    float focalPointDistance = 0.0, principlePlaneDistance = 0.0, summedThickness = 0.0;
    float rayOriginHeight = (float)((double)ld->lenses[0].aperture * 0.1);
    V3 hit_point = v3(0, 0, 0), hit_point_normal;
    V3 ray_origin = v3(0.0, rayOriginHeight, 0.0);
    V3 ray_direction = v3(0.0, 0.0, 99999.0);
    const int n = ld->lensCount;
    for (int i = 0; i < n; i++) {
        if (i == 0) summedThickness = ld->lenses[0].thickness; else summedThickness += ld->lenses[i].thickness;
        V3 sphere_center = v3(0.0, 0.0, summedThickness - ld->lenses[i].curvature);
        raySphereIntersection(&hit_point, ray_direction, ray_origin, sphere_center, ld->lenses[i].curvature, false, false);
        intersectionNormal(hit_point, sphere_center, ld->lenses[i].curvature, &hit_point_normal);
        if (i != n - 1) {
            calculateTransmissionVector(&ray_direction, ld->lenses[i].ior, ld->lenses[i + 1].ior, ray_direction, hit_point_normal, true);
        } else {
            calculateTransmissionVector(&ray_direction, ld->lenses[i].ior, 1.0, ray_direction, hit_point_normal, true);
            V3 pp_line1start = v3(0.0, rayOriginHeight, 0.0);
            V3 pp_line1end = v3(0.0, rayOriginHeight, 999999.0);
            // NB: ray_origin is still the PREVIOUS hit here (:1191-1193, :1204), as in the reference
            V3 pp_line2end = v3(0.0, (float)((double)ray_origin.y + ((double)ray_direction.y * 100000.0)),
                                (float)((double)ray_origin.z + ((double)ray_direction.z * 100000.0)));
            principlePlaneDistance = lineLineIntersectionX(pp_line1start, pp_line1end, ray_origin, pp_line2end);
            focalPointDistance = linePlaneIntersection(ray_origin, ray_direction).z;
        }
        ray_origin = hit_point;
    }
    ld->principalPlane[pass] = principlePlaneDistance;
    ld->focalPoint[pass] = focalPointDistance;
    ld->tracedFocalLength[pass] = focalPointDistance - principlePlaneDistance;
    return ld->tracedFocalLength[pass];
}

// src/zoic.cpp:1231-1237
void adjustFocalLength(Lensdata* ld) {
    for (int i = 0; i < ld->lensCount; i++) {
        ld->lenses[i].curvature *= ld->focalLengthRatio;
        ld->lenses[i].thickness *= ld->focalLengthRatio;
        ld->lenses[i].aperture *= ld->focalLengthRatio;
    }
}

// src/zoic.cpp:1297-1305
inline bool empericalOpticalVignetting(V3 origin, V3 direction, float apertureRadius, float ovRadius, float ovDistance) {
    V3 p = sub(mul(direction, ovDistance), origin);
    float pointHypotenuse = std::sqrt((p.x * p.x) + (p.y * p.y));
    float virtualApertureTrueRadius = apertureRadius * ovRadius;
    return std::fabs(pointHypotenuse) < virtualApertureTrueRadius;
}

// src/zoic.cpp:1391-1452.  Draws come from `rng` (ruling: a fresh canonical-seed stream per setup).
void exitPupilLUT(Lensdata* ld, int filmSamplesX, int boundsSamples, Xor128& rng) {
    float filmWidth = 4.0;
    float filmSpacingX = filmWidth / (float)filmSamplesX;
    const float ap0 = ld->lenses[0].aperture;
    for (int i = 0; i < filmSamplesX; i++) {
        V3 sampleOrigin = v3((float)(filmSpacingX * (float)i), 0.0, ld->originShift);
        BBox2 b = {0, 0, 0, 0};
        for (int s = 0; s < boundsSamples; s++) {
            float lensU = ((xor128(rng) / 4294967296.0f) * 2.0f) - 1.0f;
            float lensV = ((xor128(rng) / 4294967296.0f) * 2.0f) - 1.0f;
            V3 dir = v3((lensU * ap0) - sampleOrigin.x, (lensV * ap0) - sampleOrigin.y, -ld->lenses[0].thickness);
            V3 o = sampleOrigin;
            if (traceThroughLensElements(&o, &dir, ld, nullptr)) {
                if ((b.minx + b.miny) == 0.0) {  // :1423 (re-initialises whenever the sum is exactly zero)
                    b.minx = lensU * ap0; b.miny = lensV * ap0; b.maxx = lensU * ap0; b.maxy = lensV * ap0;
                }
                if ((lensU * ap0) > b.maxx) b.maxx = lensU * ap0;
                if ((lensV * ap0) > b.maxy) b.maxy = lensV * ap0;
                if ((lensU * ap0) < b.minx) b.minx = lensU * ap0;
                if ((lensV * ap0) < b.miny) b.miny = lensV * ap0;
            }
        }
        ld->lutKey.push_back(sampleOrigin.x);
        ld->lutBox.push_back(b);
    }
}

// ------------------------------------------------------------------------------------------------
// camera: src/zoic.cpp:544-643 (parameters + derived data), :1575-1720 (node_update)
// ------------------------------------------------------------------------------------------------
enum { THINLENS = 0, RAYTRACED = 1 };

struct Params {
    float sensorWidth, sensorHeight, focalLength, fStop, focalDistance;
    int useImage, lensModel, kolbSamplingLUT, useDof;
    float opticalVignettingDistance, opticalVignettingRadius, exposureControl;
    const char* lensDataPath;
    const char* bokehPath;
};

struct Camera {
    Params params;
    std::string lensPath;
    float fov = 0, tan_fov = 0, apertureRadius = 0;
    BokehImage image;
    Lensdata lens;
};

int cameraUpdate(Camera* cam, const float* img, int w, int h, int nch) {
    const Params& p = cam->params;
    if (p.useImage) {  // :1587-1593
        if (!img || w <= 0 || h <= 0 || nch <= 0) return 10;
        cam->image.x = w; cam->image.y = h; cam->image.nchannels = nch;
        cam->image.pixelData.assign(img, img + (size_t)w * h * nch);
        cam->image.build();
    }
    if (p.lensModel == THINLENS) {  // :1598-1610
        cam->fov = (float)(2.0f * atan((double)(p.sensorWidth / (2.0f * p.focalLength))));
        cam->tan_fov = tanf(cam->fov / 2.0f);
        cam->apertureRadius = (p.focalLength) / (2.0f * p.fStop);
    } else if (p.lensModel == RAYTRACED) {  // :1612-1711
        Lensdata& ld = cam->lens;
        ld.focalDistance = p.focalDistance;
        if (cam->lensPath.empty()) return 11;
        int rc = readTabularLensData(cam->lensPath, &ld);
        if (rc) return rc;
        rc = cleanupLensData(&ld);
        if (rc) return rc;
        float kolbFocalLength = traceThroughLensElementsForFocalLength(&ld, 0);
        ld.focalLengthRatio = p.focalLength / kolbFocalLength;
        adjustFocalLength(&ld);
        kolbFocalLength = traceThroughLensElementsForFocalLength(&ld, 1);
        ld.userApertureRadius = (float)((double)kolbFocalLength / (2.0 * (double)p.fStop));  // :1664
        if (ld.userApertureRadius > ld.lenses[ld.apertureElement].aperture)                  // :1668-1672
            ld.userApertureRadius = ld.lenses[ld.apertureElement].aperture;
        ld.originShift = calculateImageDistance(p.focalDistance, &ld);
        ld.apertureDistance = 0.0;
        for (int i = 0; i < ld.lensCount; i++) {
            ld.apertureDistance += ld.lenses[i].thickness;
            if (i == ld.apertureElement) break;
        }
        computeLensCenters(&ld);
        if (p.kolbSamplingLUT) {
            Xor128 rng = kCanonicalSeed;
            exitPupilLUT(&ld, 32, 100000, rng);
        }
    } else {
        return 12;
    }
    return 0;
}

// src/zoic.cpp:1752-1990.  Outputs: origin (3) + weight, dir (3) + tries.  origin0 = 0, weight0 = 1.
// The retry draws follow the g++ evaluation order of the reference build: of the two xor128() calls in
// one argument list the FIRST draw feeds the SECOND parameter (pinned in tests against oracle/_ref).
inline void createRay(const Camera* cam, float sx, float sy, float lensx, float lensy, Xor128 rng,
                      float* o4, float* d4, Stats* st) {
    const Params& params = cam->params;
    const Lensdata& ld = cam->lens;
    int tries = 0;
    const int maxtries = 25;
    V3 origin = v3(0, 0, 0), dir = v3(0, 0, 0);
    float weight = 1.0f;
    auto draw_pair = [&](float* first_param, float* second_param) {
        uint32_t k1 = xor128(rng);  // first draw -> second parameter
        uint32_t k2 = xor128(rng);
        *second_param = (float)k1 / 4294967296.0f;
        *first_param = (float)k2 / 4294967296.0f;
    };
    auto sample_lens = [&](float u, float v, float* lx, float* ly) {
        if (!params.useImage) concentricDiskSample(u, v, lx, ly);
        else cam->image.bokehSample(u, v, lx, ly);
    };

    if (params.lensModel == THINLENS) {
        V3 p = v3(sx * cam->tan_fov, sy * cam->tan_fov, 1.0);
        dir = normalize(sub(p, origin));
        V3 originOriginal = origin;
        if (params.useDof) {
            float lx = 0, ly = 0;
            sample_lens(lensx, lensy, &lx, &ly);
            lx *= cam->apertureRadius; ly *= cam->apertureRadius;
            origin = v3(lx, ly, 0.0);
            float intersection = std::fabs(params.focalDistance / dir.z);
            V3 focusPoint = mul(dir, intersection);
            dir = normalize(sub(focusPoint, origin));
            if (params.opticalVignettingDistance > 0.0f) {
                while (!empericalOpticalVignetting(origin, dir, cam->apertureRadius, params.opticalVignettingRadius,
                                                   params.opticalVignettingDistance) && tries <= maxtries) {
                    float u, v;
                    draw_pair(&u, &v);
                    sample_lens(u, v, &lx, &ly);
                    lx *= cam->apertureRadius; ly *= cam->apertureRadius;
                    dir = normalize(sub(p, originOriginal));
                    origin = v3(lx, ly, 0.0);
                    float inter = std::fabs(params.focalDistance / dir.z);
                    V3 fp = mul(dir, inter);
                    dir = normalize(sub(fp, origin));
                    ++tries;
                    if (st) ++st->attempts;
                }
            }
            if (tries > maxtries) { weight = 0.0f; if (st) ++st->vignetted; }
            else if (st) ++st->success;
        }
        dir.z *= -1.0;
    } else if (params.lensModel == RAYTRACED) {
        origin = v3((float)((double)sx * ((double)params.sensorWidth * 0.5)),
                    (float)((double)sy * ((double)params.sensorWidth * 0.5)), ld.originShift);  // :1853-1855
        const V3 kolb_origin_original = origin;
        const float ap0 = ld.lenses[0].aperture, th0 = ld.lenses[0].thickness;
        float lx = 0, ly = 0;
        sample_lens(lensx, lensy, &lx, &ly);
        if (!params.kolbSamplingLUT) {  // :1873-1888
            dir = v3((lx * ap0) - origin.x, (ly * ap0) - origin.y, -th0);
            while (!traceThroughLensElements(&origin, &dir, &ld, st) && tries <= maxtries) {
                origin = kolb_origin_original;
                float u, v;
                draw_pair(&u, &v);
                sample_lens(u, v, &lx, &ly);
                dir = v3((lx * ap0) - origin.x, (ly * ap0) - origin.y, -th0);
                ++tries;
                if (st) ++st->attempts;
            }
        } else {  // :1889-1948
            float samplingErrorCorrection = 1.05;
            float distanceFromOrigin = std::fabs(std::sqrt(origin.x * origin.x + origin.y * origin.y));
            // std::map::lower_bound over keys i*0.125; rulings: r == 0 -> entry 0, pct 0;
            // r beyond the last key -> clamp to the last entry
            const int nk = (int)ld.lutKey.size();
            int low = (int)(std::lower_bound(ld.lutKey.begin(), ld.lutKey.end(), distanceFromOrigin) - ld.lutKey.begin());
            if (low >= nk) low = nk - 1;
            float lowerBound = ld.lutKey[low];
            float theta = (float)atan2((double)origin.y, (double)origin.x);  // :1899 (C double atan2, narrowed)
            float sin = fastSin(theta);
            float cos = fastCos(theta);
            float maxScale, translation;
            if (low == 0) {
                maxScale = ld.lutBox[0].maxScale() * samplingErrorCorrection;
                translation = ld.lutBox[0].centroidX();
            } else {
                int prv = low - 1;
                float prev = ld.lutKey[prv];
                float percentage = (distanceFromOrigin - lowerBound) / (prev - lowerBound);
                maxScale = linearInterpolate(percentage, ld.lutBox[low].maxScale(), ld.lutBox[prv].maxScale()) * samplingErrorCorrection;
                translation = linearInterpolate(percentage, ld.lutBox[low].centroidX(), ld.lutBox[prv].centroidX());
            }
            lx *= maxScale; ly *= maxScale;
            lx += translation;  // :1914 (x only)
            float rx = lx * cos - ly * sin;
            float ry = lx * sin + ly * cos;
            lx = rx; ly = ry;
            dir = v3(lx - origin.x, ly - origin.y, -th0);
            while (!traceThroughLensElements(&origin, &dir, &ld, st) && tries <= maxtries) {
                origin = kolb_origin_original;
                float u, v;
                draw_pair(&u, &v);
                sample_lens(u, v, &lx, &ly);
                lx *= maxScale; ly *= maxScale;
                lx += translation; ly += translation;  // :1933 (scalar += hits BOTH components)
                rx = lx * cos - ly * sin;
                ry = lx * sin + ly * cos;
                lx = rx; ly = ry;
                dir = v3(lx - origin.x, ly - origin.y, -th0);
                ++tries;
                if (st) ++st->attempts;
            }
        }
        if (tries > maxtries) { weight = 0.0f; if (st) ++st->vignetted; }
        else if (st) ++st->success;
        dir = mul(dir, -1.0);
        origin = mul(origin, -1.0);
    }
    if (st) ++st->attempts;  // the first attempt
    float e2 = params.exposureControl * params.exposureControl;  // :1981-1987
    if (params.exposureControl > 0.0f) weight *= 1.0f + e2;
    else if (params.exposureControl < 0.0f) weight *= 1.0f / (1.0f + e2);
    o4[0] = origin.x; o4[1] = origin.y; o4[2] = origin.z; o4[3] = weight;
    d4[0] = dir.x; d4[1] = dir.y; d4[2] = dir.z; d4[3] = (float)tries;
}

void generateRange(const Camera* cam, const float* samples, uint64_t n, uint64_t first_index, uint64_t seed,
                   float* origin_w, float* dir_tries, Stats* st) {
    for (uint64_t i = 0; i < n; ++i) {
        Xor128 rng = sample_stream(seed, first_index + i);
        createRay(cam, samples[4 * i], samples[4 * i + 1], samples[4 * i + 2], samples[4 * i + 3], rng,
                  origin_w + 4 * i, dir_tries + 4 * i, st);
    }
}

// synthetic camera samples (repo convention): pixel-major / spp-minor, 24-bit uniforms
inline void synthSample(uint32_t W, uint32_t H, uint32_t spp, uint64_t seed, uint64_t i, float* s4) {
    uint64_t pix = i / spp;
    uint32_t px = (uint32_t)(pix % W), py = (uint32_t)((pix / W) % H);
    uint64_t g0 = mix64((seed ^ 0xA5A5A5A55A5A5A5Aull) + kGolden * (i + 1));
    uint64_t g1 = mix64(g0 + kGolden);
    const float inv24 = 1.0f / 16777216.0f;
    float u0 = (float)(uint32_t)(g0 & 0xFFFFFF) * inv24;
    float u1 = (float)(uint32_t)((g0 >> 32) & 0xFFFFFF) * inv24;
    float u2 = (float)(uint32_t)(g1 & 0xFFFFFF) * inv24;
    float u3 = (float)(uint32_t)((g1 >> 32) & 0xFFFFFF) * inv24;
    float fx = (float)px + u0;
    float fy = (float)py + u1;
    s4[0] = (2.0f * fx) / (float)W - 1.0f;
    s4[1] = (1.0f - (2.0f * fy) / (float)H) * ((float)H / (float)W);
    s4[2] = u2;
    s4[3] = u3;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C API (ctypes-friendly)
// ------------------------------------------------------------------------------------------------
extern "C" {

typedef Params zport_params;

void* zport_create(const zport_params* p, const float* image, int w, int h, int nch, int* err) {
    Camera* cam = new Camera();
    cam->params = *p;
    cam->lensPath = p->lensDataPath ? p->lensDataPath : "";
    cam->params.lensDataPath = nullptr;
    cam->params.bokehPath = nullptr;
    int rc = cameraUpdate(cam, image, w, h, nch);
    if (err) *err = rc;
    if (rc) { delete cam; return nullptr; }
    return cam;
}

void zport_destroy(void* c) { delete (Camera*)c; }

// stats: [0] success [1] vignetted(zero weight) [2] attempts [3] element visits [4] total internal reflections
void zport_generate(void* c, const float* samples, uint64_t n, uint64_t first_index, uint64_t seed,
                    float* origin_w, float* dir_tries, uint64_t* stats, int nthreads) {
    const Camera* cam = (const Camera*)c;
    if (nthreads < 1) nthreads = 1;
    std::vector<Stats> st(nthreads);
    if (nthreads == 1 || n < 4096) {
        generateRange(cam, samples, n, first_index, seed, origin_w, dir_tries, &st[0]);
    } else {
        std::vector<std::thread> th;
        uint64_t chunk = (n + nthreads - 1) / nthreads;
        for (int t = 0; t < nthreads; ++t) {
            uint64_t b = std::min<uint64_t>(n, chunk * t), e = std::min<uint64_t>(n, b + chunk);
            th.emplace_back([=, &st]() {
                generateRange(cam, samples + 4 * b, e - b, first_index + b, seed, origin_w + 4 * b, dir_tries + 4 * b, &st[t]);
            });
        }
        for (auto& t : th) t.join();
    }
    if (stats) {
        memset(stats, 0, 5 * sizeof(uint64_t));
        for (auto& s : st) {
            stats[0] += s.success; stats[1] += s.vignetted; stats[2] += s.attempts;
            stats[3] += s.elementVisits; stats[4] += s.tir;
        }
    }
}

// One ray with an explicit retry-stream state (for pinning the draw order against oracle/_ref).
void zport_generate_one_with_state(void* c, const float* sample, const uint32_t state[4], float* origin_w, float* dir_tries) {
    Xor128 rng = {state[0], state[1], state[2], state[3]};
    createRay((const Camera*)c, sample[0], sample[1], sample[2], sample[3], rng, origin_w, dir_tries, nullptr);
}

void zport_synth_samples(uint32_t W, uint32_t H, uint32_t spp, uint64_t seed, uint64_t first_index, uint64_t n, float* out) {
    for (uint64_t i = 0; i < n; ++i) synthSample(W, H, spp, seed, first_index + i, out + 4 * i);
}

void zport_sample_stream(uint64_t seed, uint64_t index, uint32_t out[4]) {
    Xor128 s = sample_stream(seed, index);
    out[0] = s.x; out[1] = s.y; out[2] = s.z; out[3] = s.w;
}

// Camera -> world epilogue.  Nothing in the reference computes this (Arnold applies the camera
// matrix itself after camera_create_ray; the reference's output is camera space, src/zoic.cpp:1845, :1960-1961),
// so this is the CPU statement of the contract in include/zoicb.h: one fma chain per component, innermost term
// first.  rays: n x 8 floats (origin.xyz, weight, dir.xyz, tries); m: row-major 3x4.
void zport_transform_rays(const float* rays, uint64_t n, const float* m, float* out) {
    for (uint64_t i = 0; i < n; ++i) {
        const float* r = rays + 8 * i;
        float* o = out + 8 * i;
        const float ox = r[0], oy = r[1], oz = r[2], dx = r[4], dy = r[5], dz = r[6];
        for (int k = 0; k < 3; ++k) {
            const float* mr = m + 4 * k;
            o[k] = fmaf(mr[0], ox, fmaf(mr[1], oy, fmaf(mr[2], oz, mr[3])));
            o[4 + k] = fmaf(mr[0], dx, fmaf(mr[1], dy, mr[2] * dz));
        }
        o[3] = r[3];
        o[7] = r[7];
    }
}

// Ray differentials (the reference's open TODO, src/zoic.cpp:12-13, and its stand-in hack :1971-1977
// "if (tries > 0) dOdy = origin, dDdy = dir").  CPU statement of the contract of zoicb_differentials (include/zoicb.h):
// the derivative of the generated ray with respect to the screen position AT A FIXED POINT OF THE APERTURE, as forward
// differences over one pixel (dsx, dsy = Arnold's AtCameraInput::dsx / dsy):
//   1. a sample whose ray has weight 0 gets four zero vectors;
//   2. the aperture draw of the ACCEPTED attempt is (lensx, lensy) when tries == 0, otherwise the tries-th pair of the
//      sample's retry stream; it is mapped to its aim point exactly like createRay maps it (thin lens: lens point
//      (lx R, ly R, 0); raytraced: scaled / translated / rotated by the exit-pupil LUT of the sample's OWN film point,
//      with the retry arithmetic of :1933 when tries > 0);
//   3. the base ray (sx, sy), the x-neighbour (fl(sx + dsx), sy) and the y-neighbour (sx, fl(sy + dsy)) each run from
//      their film point through that one aim point: thin lens -- dir = normalize(focus(p') - lens point), origin = lens
//      point; raytraced -- (aim - film', -thickness0) marched through the stack with traceThroughLensElements, both
//      vectors negated like the reference's output;
//   4. dOdx = origin_x - origin_base, dDdx = dir_x - dir_base (component-wise fp32 subtractions), likewise for y; a
//      neighbour that is stopped inside the lens gives zero vectors for its axis.
// diffs: n x 12 floats (dOdx, dOdy, dDdx, dDdy).
void zport_differentials(void* c, const float* samples, uint64_t n, uint64_t first_index, uint64_t seed, float dsx, float dsy,
                         const float* tries_in, const float* weight_in, float* diffs) {
    const Camera* cam = (const Camera*)c;
    const Params& params = cam->params;
    const Lensdata& ld = cam->lens;
    for (uint64_t i = 0; i < n; ++i) {
        float* out = diffs + 12 * i;
        for (int k = 0; k < 12; ++k) out[k] = 0.0f;
        if (weight_in[i] == 0.0f) continue;
        const float sx = samples[4 * i], sy = samples[4 * i + 1];
        float u = samples[4 * i + 2], v = samples[4 * i + 3];
        const int tries = (int)tries_in[i];
        if (tries > 0) {
            Xor128 rng = sample_stream(seed, first_index + i);
            for (int t = 0; t < tries; ++t) {
                uint32_t k1 = xor128(rng), k2 = xor128(rng);  // first draw -> second parameter
                v = (float)k1 / 4294967296.0f;
                u = (float)k2 / 4294967296.0f;
            }
        }
        float lx = 0, ly = 0;
        if (!params.useImage) concentricDiskSample(u, v, &lx, &ly);
        else cam->image.bokehSample(u, v, &lx, &ly);
        V3 o[3], d[3];
        bool ok[3] = {true, true, true};
        const float fsx[3] = {sx, sx + dsx, sx}, fsy[3] = {sy, sy, sy + dsy};
        if (params.lensModel == THINLENS) {
            for (int k = 0; k < 3; ++k) {
                V3 p = v3(fsx[k] * cam->tan_fov, fsy[k] * cam->tan_fov, 1.0);
                d[k] = normalize(p);
                o[k] = v3(0, 0, 0);
                if (params.useDof) {
                    o[k] = v3(lx * cam->apertureRadius, ly * cam->apertureRadius, 0.0);
                    float intersection = std::fabs(params.focalDistance / d[k].z);
                    V3 focusPoint = mul(d[k], intersection);
                    d[k] = normalize(sub(focusPoint, o[k]));
                }
                d[k].z *= -1.0;
            }
        } else {
            const float half = (float)((double)params.sensorWidth * 0.5);
            const float ap0 = ld.lenses[0].aperture, th0 = ld.lenses[0].thickness;
            float ax, ay;
            {
                V3 origin = v3((float)((double)sx * ((double)params.sensorWidth * 0.5)),
                               (float)((double)sy * ((double)params.sensorWidth * 0.5)), ld.originShift);
                if (!params.kolbSamplingLUT) {
                    ax = lx * ap0; ay = ly * ap0;
                } else {  // the LUT arithmetic of createRay, :1889-1948
                    float samplingErrorCorrection = 1.05;
                    float distanceFromOrigin = std::fabs(std::sqrt(origin.x * origin.x + origin.y * origin.y));
                    const int nk = (int)ld.lutKey.size();
                    int low = (int)(std::lower_bound(ld.lutKey.begin(), ld.lutKey.end(), distanceFromOrigin) - ld.lutKey.begin());
                    if (low >= nk) low = nk - 1;
                    float lowerBound = ld.lutKey[low];
                    float theta = (float)atan2((double)origin.y, (double)origin.x);
                    float sin = fastSin(theta), cos = fastCos(theta);
                    float maxScale, translation;
                    if (low == 0) {
                        maxScale = ld.lutBox[0].maxScale() * samplingErrorCorrection;
                        translation = ld.lutBox[0].centroidX();
                    } else {
                        int prv = low - 1;
                        float prev = ld.lutKey[prv];
                        float percentage = (distanceFromOrigin - lowerBound) / (prev - lowerBound);
                        maxScale = linearInterpolate(percentage, ld.lutBox[low].maxScale(), ld.lutBox[prv].maxScale()) * samplingErrorCorrection;
                        translation = linearInterpolate(percentage, ld.lutBox[low].centroidX(), ld.lutBox[prv].centroidX());
                    }
                    lx *= maxScale; ly *= maxScale;
                    lx += translation;
                    if (tries > 0) ly += translation;   // :1933
                    ax = lx * cos - ly * sin;
                    ay = lx * sin + ly * cos;
                }
            }
            (void)half;
            for (int k = 0; k < 3; ++k) {
                o[k] = v3((float)((double)fsx[k] * ((double)params.sensorWidth * 0.5)),
                          (float)((double)fsy[k] * ((double)params.sensorWidth * 0.5)), ld.originShift);
                d[k] = v3(ax - o[k].x, ay - o[k].y, -th0);
                ok[k] = traceThroughLensElements(&o[k], &d[k], &ld, nullptr);
                d[k] = mul(d[k], -1.0);
                o[k] = mul(o[k], -1.0);
            }
        }
        if (!ok[0]) continue;   // cannot happen for a consistent (tries, weight): the accepted attempt passes
        if (ok[1]) {
            V3 a = sub(o[1], o[0]), b = sub(d[1], d[0]);
            out[0] = a.x; out[1] = a.y; out[2] = a.z; out[6] = b.x; out[7] = b.y; out[8] = b.z;
        }
        if (ok[2]) {
            V3 a = sub(o[2], o[0]), b = sub(d[2], d[0]);
            out[3] = a.x; out[4] = a.y; out[5] = a.z; out[9] = b.x; out[10] = b.y; out[11] = b.z;
        }
    }
}

// Camera -> world for the differentials (contract of zoicb_transform_differentials): every one of the four vectors times
// the 3x3 part of the row-major 3x4 matrix, v'_r = fma(m[r][0], vx, fma(m[r][1], vy, m[r][2] * vz)).
void zport_transform_differentials(const float* diffs, uint64_t n, const float* m, float* out) {
    for (uint64_t i = 0; i < n; ++i)
        for (int k = 0; k < 4; ++k) {
            const float* v = diffs + 12 * i + 3 * k;
            float* o = out + 12 * i + 3 * k;
            for (int r = 0; r < 3; ++r) o[r] = fmaf(m[4 * r], v[0], fmaf(m[4 * r + 1], v[1], m[4 * r + 2] * v[2]));
        }
}

// Derived camera state for host-side parity tests.
//   scalars[16]: fov, tan_fov, apertureRadius, userApertureRadius, originShift, apertureDistance,
//                focalLengthRatio, tracedFocalLength[0..1], principalPlane[0..1], focalPoint[0..1],
//                lensCount, apertureElement, lutSize
//   lenses: lensCount x 5 (curvature, thickness, ior, aperture, center); lut: lutSize x 6 (key, minx, miny, maxx, maxy, 0)
int zport_get_constants(void* c, float* scalars, float* lenses, int max_lenses, float* lut, int max_lut) {
    const Camera* cam = (const Camera*)c;
    const Lensdata& ld = cam->lens;
    float s[16] = {cam->fov, cam->tan_fov, cam->apertureRadius, ld.userApertureRadius, ld.originShift,
                   ld.apertureDistance, ld.focalLengthRatio, ld.tracedFocalLength[0], ld.tracedFocalLength[1],
                   ld.principalPlane[0], ld.principalPlane[1], ld.focalPoint[0], ld.focalPoint[1],
                   (float)ld.lensCount, (float)ld.apertureElement, (float)ld.lutKey.size()};
    memcpy(scalars, s, sizeof s);
    for (int i = 0; i < ld.lensCount && i < max_lenses; ++i) {
        const LensElement& e = ld.lenses[i];
        float row[5] = {e.curvature, e.thickness, e.ior, e.aperture, e.center};
        memcpy(lenses + 5 * i, row, sizeof row);
    }
    for (int i = 0; i < (int)ld.lutKey.size() && i < max_lut; ++i) {
        const BBox2& b = ld.lutBox[i];
        float row[6] = {ld.lutKey[i], b.minx, b.miny, b.maxx, b.maxy, 0.0f};
        memcpy(lut + 6 * i, row, sizeof row);
    }
    return 0;
}

// bokeh tables for host-side parity tests; returns x*y (0 when there is no image)
int zport_get_bokeh_tables(void* c, float* cdfRow, int* rowIndices, float* cdfColumn, int* columnIndices) {
    const Camera* cam = (const Camera*)c;
    const BokehImage& im = cam->image;
    if (!im.isValid()) return 0;
    if (cdfRow) memcpy(cdfRow, im.cdfRow.data(), im.y * sizeof(float));
    if (rowIndices) memcpy(rowIndices, im.rowIndices.data(), im.y * sizeof(int));
    if (cdfColumn) memcpy(cdfColumn, im.cdfColumn.data(), (size_t)im.x * im.y * sizeof(float));
    if (columnIndices) memcpy(columnIndices, im.columnIndices.data(), (size_t)im.x * im.y * sizeof(int));
    return im.x * im.y;
}

}  // extern "C"
