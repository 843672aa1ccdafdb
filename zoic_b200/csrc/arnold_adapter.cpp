// arnold_adapter.cpp -- the Arnold-shaped surface of the camera: libzoic_arnold.so.
//
// Exports the one C symbol an Arnold plugin exports, NodeLoader, and behind it the same six node callbacks
// with the same node name ("zoic"), node type, parameter names and defaults as the reference plugin
// (reference src/zoic.cpp:61, :1547-1572, :1575-1720, :1723-1749, :1752-1995, :1999-2007).  Every callback
// forwards to the C ABI of libzoicb (include/zoicb.h):
//
//     node_initialize    -> allocates the node's local data
//     node_update        -> zoicb_create (re-created when a parameter changed; the reference's own
//                           lensChanged()/bokehChanged() shortcut, :595-611)
//     camera_create_ray  -> zoicb_generate_one (Arnold hands over one sample per call; renderers that can
//                           batch call zoicb_generate directly)
//     node_finish        -> zoicb_get_stats (the counters the reference prints, :1729-1732) + zoicb_destroy
//     camera_reverse_ray -> false, as in the reference (:1992-1995)
//
// Error convention: where the reference logs AiMsgError + AiRenderAbort() and carries on with undefined
// state, this adapter logs the zoicb error text, aborts the render and leaves the node without a camera, in
// which case camera_create_ray returns a zero-weight ray.
//
// Built against include/arnold_shim/ai.h when no Arnold SDK is present.
#include <ai.h>

#include <atomic>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/zoicb.h"

AI_CAMERA_NODE_EXPORT_METHODS(zoicB200Methods)

namespace {

const char* kLensModelNames[] = {"THINLENS", "RAYTRACED", NULL};

struct NodeParams {
    float sensorWidth = 0, sensorHeight = 0, focalLength = 0, fStop = 0, focalDistance = 0;
    bool useImage = false;
    std::string bokehPath;
    int lensModel = -1;
    std::string lensDataPath;
    bool kolbSamplingLUT = false, useDof = false;
    float opticalVignettingDistance = 0, opticalVignettingRadius = 0, exposureControl = 0;

    void read(AtNode* node) {
        sensorWidth = AiNodeGetFlt(node, "sensorWidth");
        sensorHeight = AiNodeGetFlt(node, "sensorHeight");
        focalLength = AiNodeGetFlt(node, "focalLength");
        fStop = AiNodeGetFlt(node, "fStop");
        focalDistance = AiNodeGetFlt(node, "focalDistance");
        useImage = AiNodeGetBool(node, "useImage");
        bokehPath = AiNodeGetStr(node, "bokehPath").c_str();
        lensModel = AiNodeGetInt(node, "lensModel");
        lensDataPath = AiNodeGetStr(node, "lensDataPath").c_str();
        kolbSamplingLUT = AiNodeGetBool(node, "kolbSamplingLUT");
        useDof = AiNodeGetBool(node, "useDof");
        opticalVignettingDistance = AiNodeGetFlt(node, "opticalVignettingDistance");
        opticalVignettingRadius = AiNodeGetFlt(node, "opticalVignettingRadius");
        exposureControl = AiNodeGetFlt(node, "exposureControl");
    }
    bool operator==(const NodeParams& o) const {
        return sensorWidth == o.sensorWidth && sensorHeight == o.sensorHeight && focalLength == o.focalLength &&
               fStop == o.fStop && focalDistance == o.focalDistance && useImage == o.useImage &&
               bokehPath == o.bokehPath && lensModel == o.lensModel && lensDataPath == o.lensDataPath &&
               kolbSamplingLUT == o.kolbSamplingLUT && useDof == o.useDof &&
               opticalVignettingDistance == o.opticalVignettingDistance &&
               opticalVignettingRadius == o.opticalVignettingRadius && exposureControl == o.exposureControl;
    }
};

struct NodeData {
    zoicb_ctx* ctx = nullptr;
    NodeParams params;
    std::atomic<uint64_t> next_index{0};  // the retry stream of a sample is keyed by its arrival number
};

const uint64_t kAdapterSeed = 0;

}  // namespace

node_parameters {
    AiParameterFlt("sensorWidth", 3.6f);
    AiParameterFlt("sensorHeight", 2.4f);
    AiParameterFlt("focalLength", 2.0f);
    AiParameterFlt("fStop", 4.0f);
    AiParameterFlt("focalDistance", 100.0f);
    AiParameterBool("useImage", false);
    AiParameterStr("bokehPath", "");
    AiParameterEnum("lensModel", ZOICB_RAYTRACED, kLensModelNames);
    AiParameterStr("lensDataPath", "");
    AiParameterBool("kolbSamplingLUT", true);
    AiParameterBool("useDof", true);
    AiParameterFlt("opticalVignettingDistance", 0.0f);
    AiParameterFlt("opticalVignettingRadius", 1.0f);
    AiParameterFlt("exposureControl", 0.0f);
    (void)nentry;
}

node_initialize {
    AiCameraInitialize(node);
    AiNodeSetLocalData(node, new NodeData());
}

node_update {
    AiCameraUpdate(node, false);
    NodeData* data = (NodeData*)AiNodeGetLocalData(node);
    NodeParams now;
    now.read(node);
    if (data->ctx && now == data->params) {
        AiMsgWarning("[ZOIC] Skipping node update, parameters didn't change.");
        return;
    }
    if (data->ctx) { zoicb_destroy(data->ctx); data->ctx = nullptr; }
    data->params = now;

    std::vector<float> pixels;
    unsigned w = 0, h = 0, nch = 0;
    if (now.useImage) {
        const AtString path(now.bokehPath.c_str());
        if (!AiTextureGetResolution(path, &w, &h) || !AiTextureGetNumChannels(path, &nch)) {
            AiMsgError("[ZOIC] Couldn't open bokeh image!");
            AiRenderAbort();
            return;
        }
        pixels.resize((size_t)w * h * nch);
        if (!AiTextureLoad(path, true, 0, pixels.data())) {
            AiMsgError("[ZOIC] Couldn't open bokeh image!");
            AiRenderAbort();
            return;
        }
    }
    zoicb_params p;
    zoicb_default_params(&p);
    p.sensorWidth = now.sensorWidth; p.sensorHeight = now.sensorHeight; p.focalLength = now.focalLength;
    p.fStop = now.fStop; p.focalDistance = now.focalDistance; p.useImage = now.useImage ? 1 : 0;
    p.lensModel = now.lensModel; p.kolbSamplingLUT = now.kolbSamplingLUT ? 1 : 0; p.useDof = now.useDof ? 1 : 0;
    p.opticalVignettingDistance = now.opticalVignettingDistance;
    p.opticalVignettingRadius = now.opticalVignettingRadius; p.exposureControl = now.exposureControl;
    p.lensDataPath = now.lensDataPath.c_str();
    p.bokehPath = now.bokehPath.c_str();
    zoicb_status rc = zoicb_create(&p, pixels.empty() ? nullptr : pixels.data(), (int)w, (int)h, (int)nch, 0, &data->ctx);
    if (rc != ZOICB_OK) {
        AiMsgError("[ZOIC] %s", zoicb_last_error());
        AiRenderAbort();
        data->ctx = nullptr;
        return;
    }
    zoicb_constants c;
    zoicb_get_constants(data->ctx, &c);
    if (now.lensModel == ZOICB_RAYTRACED) {
        AiMsgInfo("%-40s %12d", "[ZOIC] Aperture is lens element number", c.apertureElement);
        AiMsgInfo("%-40s %12.8f", "[ZOIC] Adj. Raytraced Focal Length [cm]", c.tracedFocalLength[1]);
        AiMsgInfo("%-40s %12.8f", "[ZOIC] User aperture radius [cm]", c.userApertureRadius);
        AiMsgInfo("%-40s %12.8f", "[ZOIC] Image distance [cm]", c.originShift);
        AiMsgInfo("%-40s %12.8f", "[ZOIC] Aperture distance [cm]", c.apertureDistance);
    }
}

node_finish {
    NodeData* data = (NodeData*)AiNodeGetLocalData(node);
    if (data->ctx) {
        zoicb_stats s;
        if (zoicb_get_stats(data->ctx, &s) == ZOICB_OK) {
            AiMsgInfo("%-40s %12llu", "[ZOIC] Succesful rays", (unsigned long long)s.success);
            AiMsgInfo("%-40s %12llu", "[ZOIC] Vignetted rays", (unsigned long long)s.vignetted);
            AiMsgInfo("%-40s %12.8f", "[ZOIC] Vignetted Percentage",
                      100.0 * (double)s.vignetted / (double)(s.success + s.vignetted ? s.success + s.vignetted : 1));
            AiMsgInfo("%-40s %12llu", "[ZOIC] Total internal reflection cases", (unsigned long long)s.total_internal_reflection);
        }
        zoicb_destroy(data->ctx);
    }
    delete data;
}

camera_create_ray {
    (void)tid;
    NodeData* data = (NodeData*)AiNodeGetLocalData(node);
    if (!data->ctx) { output.weight = 0.0f; return; }
    const float sample[4] = {input.sx, input.sy, input.lensx, input.lensy};
    zoicb_ray r;
    const uint64_t index = data->next_index.fetch_add(1, std::memory_order_relaxed);
    if (zoicb_generate_one(data->ctx, sample, index, kAdapterSeed, &r) != ZOICB_OK) { output.weight = 0.0f; return; }
    output.origin = AtVector(r.origin[0], r.origin[1], r.origin[2]);
    output.dir = AtVector(r.dir[0], r.dir[1], r.dir[2]);
    // the batched ABI starts from weight 1; Arnold's incoming weight is multiplied in (reference :1983-1986)
    output.weight *= r.weight;
    if (r.tries > 0.0f) {  // the reference's derivative workaround for re-sampled rays (:1974-1977)
        output.dOdy = output.origin;
        output.dDdy = output.dir;
    }
}

camera_reverse_ray {
    (void)node; (void)Po; (void)Ro; (void)relative_time; (void)Ps;
    return false;
}

node_loader {
    if (i > 0) return false;
    node->methods = zoicB200Methods;
    node->output_type = AI_TYPE_NONE;
    node->name = "zoic";
    node->node_type = AI_NODE_CAMERA;
    strcpy(node->version, AI_VERSION);
    return true;
}
