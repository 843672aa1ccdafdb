"""The benchmark / parity configurations of BASELINE.json, as camera parameters + synthetic sample grids.

A W x H x spp frame is laid out in `passes` passes of spp / passes samples per pixel:
sample index i = pass * (W*H*spp/passes) + pixel * (spp/passes) + s, i.e. zoicb_synth_samples with spp_per_pass samples
per pixel, whose pixel index wraps once per pass.  The layout is FIXED per workload (8 passes for the big frames), not a
function of the GPU count: G GPUs take passes [r*8/G, (r+1)*8/G) = the contiguous index range [r N/G, (r+1) N/G), every
rank sees the whole film (vignetting is balanced), and the rays of a job are the same bits for every G.  The four
uniforms of sample i come from a counter hash of (seed, i), so any slice of any configuration can
be regenerated on any machine, CPU or GPU, without storing it.
"""
import os

from .synth import hex_bokeh_image

LENS_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "lenses")

# file -> (f-number of the design, focal length used by config 5)
LENSES = {
    "petzval_f1.25.dat": (1.25, 5.0),
    "petzval_f1.6.dat": (1.6, 5.0),
    "double_gauss_f2.0.dat": (2.0, 5.0),
    "triplet_f2.5.dat": (2.5, 5.0),
    "mori_f2.8.dat": (2.8, 5.0),
    "tessar_f2.8.dat": (2.8, 5.0),
    "fisheye_muller_f4.0.dat": (4.0, 1.0),
    "telephoto_f5.0.dat": (5.0, 5.0),
}


def lens_path(name):
    return os.path.join(LENS_DIR, name)


COMMON = dict(sensorWidth=3.6, sensorHeight=2.4, focalDistance=100.0, useDof=1, exposureControl=0.0)


class Workload:
    def __init__(self, name, W, H, spp, seed, params, image=None, note="", passes=8):
        self.name, self.W, self.H, self.spp, self.seed = name, W, H, spp, seed
        self.passes = passes if spp % passes == 0 else 1
        self.params = dict(COMMON)
        self.params.update(params)
        self._image = image
        self.note = note

    @property
    def n(self):
        return self.W * self.H * self.spp

    @property
    def spp_per_pass(self):
        return self.spp // self.passes

    def synth_args(self):
        """(W, H, spp_per_pass, seed): the first four arguments of synth_samples for this frame."""
        return self.W, self.H, self.spp_per_pass, self.seed

    def image(self):
        return hex_bokeh_image(255) if self._image == "hex255" else None

    def describe(self):
        p = {k: (os.path.basename(v) if k == "lensDataPath" else v) for k, v in self.params.items()}
        return {"workload": self.name, "W": self.W, "H": self.H, "spp": self.spp, "passes": self.passes,
                "samples": self.n, "seed": self.seed, "params": p, "bokeh_image": self._image}


def _kolb(lens, focal, fstop):
    return dict(lensModel=1, lensDataPath=lens_path(lens), focalLength=focal, fStop=fstop, kolbSamplingLUT=1)


def config1():
    return Workload("config1: thin-lens 1920x1080x1", 1920, 1080, 1, 0x200C + 1,
                    dict(lensModel=0, focalLength=3.5, fStop=2.8), passes=1)


def config2(spp=64):
    return Workload("config2: Kolb double-gauss f/2.0 3840x2160x%d" % spp, 3840, 2160, spp, 0x200C + 2,
                    _kolb("double_gauss_f2.0.dat", 5.0, 2.0))


def headline():
    w = config2(256)
    w.name = "headline: Kolb double-gauss f/2.0 3840x2160x256 (4Kx256spp)"
    return w


def config3():
    return Workload("config3: thin-lens + optical vignetting + hex bokeh image 3840x2160x256", 3840, 2160, 256,
                    0x200C + 3, dict(lensModel=0, focalLength=3.5, fStop=2.8, opticalVignettingDistance=2.0,
                                     opticalVignettingRadius=1.0, useImage=1), image="hex255")


def config4():
    return Workload("config4: Kolb fisheye f/4.0 7680x4320x128", 7680, 4320, 128, 0x200C + 4,
                    _kolb("fisheye_muller_f4.0.dat", 1.0, 4.0))


def config5(lens):
    fnum, focal = LENSES[lens]
    return Workload("config5: Kolb %s 7680x4320x1024" % lens, 7680, 4320, 1024, 0x200C + 5, _kolb(lens, focal, fnum))


def _config5_factory(lens):
    return lambda: config5(lens)


BY_NAME = {"config1": config1, "config2": config2, "headline": headline, "config3": config3, "config4": config4}
# config 5 = the sweep over all eight lens tables: "config5:<lens file>" each, "config5" = the first of the sweep
for _lens in sorted(LENSES):
    BY_NAME["config5:" + _lens] = _config5_factory(_lens)
BY_NAME["config5"] = BY_NAME["config5:" + sorted(LENSES)[0]]
CONFIG5 = ["config5:" + _lens for _lens in sorted(LENSES)]
