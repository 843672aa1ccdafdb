"""zoic_b200 -- batched, B200-native camera-ray generation with the behaviour of the zoic Arnold camera.

Only what the hot path needs lives here: csrc/ (CUDA kernels, host setup, the C ABI and the Arnold-shaped
adapter), the ctypes binding (capi), the host-side mirror of the camera node (camera) and synthetic
workloads (synth).
"""
from .camera import (ZoicCamera, Gather, build_bokeh_tables, debug_lut_boxes, host_setup, make_params, nccl_unique_id,  # noqa: F401
                     split_rays, unpack_planes, THINLENS, RAYTRACED, MODE_EXACT, MODE_GUARDED)
