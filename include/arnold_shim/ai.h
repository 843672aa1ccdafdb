// ai.h -- minimal stand-in for the Arnold 5 SDK umbrella header.
//
// The Arnold SDK is closed source and is not available in this repository's
// build environments.  This header declares exactly the slice of the SDK that a
// "zoic"-style camera node touches (SURVEY.md section 8(b) lists it: 18
// functions, 4 parameter macros, 2 inline math helpers, 4 types, 6 constants),
// so that
//   * zoic_b200's Arnold-shaped adapter (zoic_b200/csrc/arnold_adapter.cpp), and
//   * the unmodified reference translation unit (built by oracle/Makefile into
//     oracle/_ref/ as the parity oracle)
// compile against the same definitions.  Where the real SDK is present, build
// with its include directory instead of this one.
//
// The inline math below is the de-facto arithmetic standard for parity
// ("third-party arithmetic"): length = sqrtf(x*x+y*y+z*z),
// normalise multiplies by 1/length when length != 0, dot sums left to right.
#pragma once

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>

// ---------------------------------------------------------------- constants
#define AI_PI      3.14159265358979323846f
#define AI_PIOVER2 1.57079632679489661923f
#define AI_VERSION "5.0.2.0"
#define AI_NODE_CAMERA 0x0002
#define AI_TYPE_NONE   0xFF
#define AI_MAXSTRING   64

// ---------------------------------------------------------------- vector types
struct AtVector2 {
    float x, y;
    AtVector2() {}
    AtVector2(float _x, float _y) : x(_x), y(_y) {}
    AtVector2& operator*=(float f) { x *= f; y *= f; return *this; }
    // scalar += adds to BOTH components (the reference's retry path relies on it)
    AtVector2& operator+=(float f) { x += f; y += f; return *this; }
    AtVector2& operator+=(const AtVector2& o) { x += o.x; y += o.y; return *this; }
};

struct AtVector {
    float x, y, z;
    AtVector() {}
    AtVector(float _x, float _y, float _z) : x(_x), y(_y), z(_z) {}
    AtVector operator+(const AtVector& o) const { return AtVector(x + o.x, y + o.y, z + o.z); }
    AtVector operator-(const AtVector& o) const { return AtVector(x - o.x, y - o.y, z - o.z); }
    AtVector operator-() const { return AtVector(-x, -y, -z); }
    AtVector operator*(float f) const { return AtVector(x * f, y * f, z * f); }
    AtVector operator/(float f) const { float c = 1 / f; return AtVector(x * c, y * c, z * c); }
    AtVector& operator*=(float f) { x *= f; y *= f; z *= f; return *this; }
    AtVector& operator+=(const AtVector& o) { x += o.x; y += o.y; z += o.z; return *this; }
};
inline AtVector operator*(float f, const AtVector& v) { return v * f; }

struct AtRGB {
    float r, g, b;
    AtRGB() {}
    AtRGB(float v) : r(v), g(v), b(v) {}
    AtRGB(float _r, float _g, float _b) : r(_r), g(_g), b(_b) {}
    AtRGB& operator=(float v) { r = g = b = v; return *this; }
    AtRGB& operator*=(float f) { r *= f; g *= f; b *= f; return *this; }
};

static const AtVector2 AI_P2_ZERO(0.0f, 0.0f);

inline float AiV3Dot(const AtVector& a, const AtVector& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float AiV3Length(const AtVector& a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
inline AtVector AiV3Normalize(const AtVector& a) {
    float tmp = AiV3Length(a);
    if (tmp != 0) tmp = 1 / tmp;
    return a * tmp;
}

// ---------------------------------------------------------------- strings
class AtString {
    const char* s_;
public:
    AtString() : s_("") {}
    explicit AtString(const char* s) : s_(s ? s : "") {}
    const char* c_str() const { return s_; }
    operator const char*() const { return s_; }
};

// ---------------------------------------------------------------- nodes
struct AtNode;
struct AtList;
struct AtNodeEntry;

struct AtCameraInput {
    float sx, sy;          // screen-space sample, sx in [-1,1], sy in [-1/aspect, 1/aspect]
    float dsx, dsy;        // derivatives of the screen sample (unused by zoic)
    float lensx, lensy;    // lens sample in [0,1)^2
    float relative_time;
};

struct AtCameraOutput {
    AtVector origin, dir;
    AtVector dOdx, dOdy, dDdx, dDdy;
    AtRGB weight;
};

struct AtCommonMethods {
    void (*Parameters)(AtList*, AtNodeEntry*);
    void (*Initialize)(AtNode*);
    void (*Update)(AtNode*);
    void (*Finish)(AtNode*);
};
struct AtCameraNodeMethods {
    void (*CreateRay)(const AtNode*, const AtCameraInput&, AtCameraOutput&, int tid);
    bool (*ReverseRay)(const AtNode*, const AtVector& Po, const AtVector& Ro, float relative_time, AtVector2& Ps);
};
struct AtNodeMethods {
    const AtCommonMethods* cmethods;
    const void* dmethods;
};

struct AtNodeLib {
    int node_type;
    uint8_t output_type;
    const char* name;
    const void* methods;
    char version[AI_MAXSTRING];
};

// Method-table plumbing.  AI_CAMERA_NODE_EXPORT_METHODS(tag) forward-declares the six
// callbacks and defines `const AtNodeMethods* tag`.
#define AI_CAMERA_NODE_EXPORT_METHODS(tag)                                                        \
    static void Parameters(AtList*, AtNodeEntry*);                                                \
    static void Initialize(AtNode*);                                                              \
    static void Update(AtNode*);                                                                  \
    static void Finish(AtNode*);                                                                  \
    static void CameraCreateRay(const AtNode*, const AtCameraInput&, AtCameraOutput&, int);       \
    static bool CameraReverseRay(const AtNode*, const AtVector&, const AtVector&, float, AtVector2&); \
    static AtCommonMethods ai_common_mtds = {Parameters, Initialize, Update, Finish};             \
    static AtCameraNodeMethods ai_cam_mtds = {CameraCreateRay, CameraReverseRay};                 \
    static AtNodeMethods ai_mtds = {&ai_common_mtds, &ai_cam_mtds};                               \
    const AtNodeMethods* tag = &ai_mtds;

#define node_parameters    static void Parameters(AtList* params, AtNodeEntry* nentry)
#define node_initialize    static void Initialize(AtNode* node)
#define node_update        static void Update(AtNode* node)
#define node_finish        static void Finish(AtNode* node)
#define camera_create_ray  static void CameraCreateRay(const AtNode* node, const AtCameraInput& input, AtCameraOutput& output, int tid)
#define camera_reverse_ray static bool CameraReverseRay(const AtNode* node, const AtVector& Po, const AtVector& Ro, float relative_time, AtVector2& Ps)
#define node_loader        extern "C" __attribute__((visibility("default"))) bool NodeLoader(int i, AtNodeLib* node)

// Parameter declaration: forwarded to the host so an embedding application can record names and
// defaults (the real SDK stores them in the node entry).
void AiShimDeclareFlt(AtList*, const char* name, float dflt);
void AiShimDeclareBool(AtList*, const char* name, bool dflt);
void AiShimDeclareStr(AtList*, const char* name, const char* dflt);
void AiShimDeclareEnum(AtList*, const char* name, int dflt, const char** names);
#define AiParameterFlt(n, d)      AiShimDeclareFlt(params, n, d);
#define AiParameterBool(n, d)     AiShimDeclareBool(params, n, d);
#define AiParameterStr(n, d)      AiShimDeclareStr(params, n, d);
#define AiParameterEnum(n, d, e)  AiShimDeclareEnum(params, n, d, e);

// ---------------------------------------------------------------- host API used by the node
float    AiNodeGetFlt(const AtNode*, const char* name);
bool     AiNodeGetBool(const AtNode*, const char* name);
int      AiNodeGetInt(const AtNode*, const char* name);
AtString AiNodeGetStr(const AtNode*, const char* name);
void     AiNodeSetLocalData(AtNode*, void* data);
void*    AiNodeGetLocalData(const AtNode*);

void AiCameraInitialize(AtNode*);
void AiCameraUpdate(AtNode*, bool plane_distance);

void AiMsgInfo(const char* fmt, ...);
void AiMsgWarning(const char* fmt, ...);
void AiMsgError(const char* fmt, ...);
void AiRenderAbort();

void* AiMalloc(size_t bytes);
void  AiFree(void* p);
void  AiAddMemUsage(int64_t bytes, const AtString category);

bool AiTextureGetResolution(const AtString path, unsigned int* w, unsigned int* h);
bool AiTextureGetNumChannels(const AtString path, unsigned int* nch);
bool AiTextureLoad(const AtString path, bool use_float, unsigned int mipmap, void* out);
