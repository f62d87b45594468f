"""ctypes binding of oracle/_ref/libref_ocl.so — the reference's UNMODIFIED OpenCL kernels run on a real OpenCL device
(the B200 itself, through the NVIDIA driver's OpenCL runtime, when the GPU box has one), with the reference's host sequences
restated around them (oracle/ref_build/ocl_host.cpp).  TEST INFRASTRUCTURE: tests/ and tests/probes/ref_opencl_bench.py only.

available() is False when the library was not built or no OpenCL platform / device can be opened (this container)."""
import ctypes as C
import os

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(_ROOT, "oracle", "_ref", "libref_ocl.so")
_lib = None
_ok = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(SO)
        _lib.ocl_last_error.restype = C.c_char_p
        _lib.ocl_info.restype = C.c_char_p
    return _lib


def error():
    return lib().ocl_last_error().decode(errors="replace")


def available():
    global _ok
    if _ok is None:
        _ok = os.path.exists(SO) and lib().ocl_init() == 0
    return _ok


def info():
    return lib().ocl_info().decode(errors="replace")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _check(rc):
    if rc != 0:
        raise RuntimeError("reference OpenCL host: " + error())


def set_nearest(on):
    """kernels built from now on: CLK_FILTER_LINEAR -> CLK_FILTER_NEAREST (the filter OpenCL defines for integer images) or, with
    on=False, the text as shipped (linear filtering of integer images: undefined by the specification)"""
    lib().ocl_set_nearest(1 if on else 0)


def fetch_stats(vol):
    vol = np.ascontiguousarray(vol, dtype=np.int16)
    nz, ny, nx = vol.shape
    st = (C.c_int32 * 4)()
    ms = C.c_double(0)
    _check(lib().ocl_fetch_stats(_p(vol), nx, ny, nz, st, C.byref(ms)))
    return list(st), ms.value


def histogram(vol, width, height, rng):
    """tf_sort_values (renderer.cpp:49-61); rng minima must be the volume's own stats (no negative index)"""
    vol = np.ascontiguousarray(vol, dtype=np.int16)
    nz, ny, nx = vol.shape
    bins = np.zeros(width * height, dtype=np.uint32)
    ms = C.c_double(0)
    _check(lib().ocl_histogram(_p(vol), nx, ny, nz, width, height, (C.c_float * 4)(*[float(x) for x in rng]), _p(bins), C.byref(ms)))
    return bins, ms.value


def bilateral(vol):
    vol = np.ascontiguousarray(vol, dtype=np.int16)
    nz, ny, nx = vol.shape
    out = np.zeros_like(vol)
    ms = C.c_double(0)
    _check(lib().ocl_bilateral(_p(vol), nx, ny, nz, _p(out), C.byref(ms)))
    return out, ms.value


def clip(vol, start, size):
    vol = np.ascontiguousarray(vol, dtype=np.int16)
    nz, ny, nx = vol.shape
    out = np.zeros((size[2], size[1], size[0]), dtype=np.int16)
    ms = C.c_double(0)
    _check(lib().ocl_clip(_p(vol), nx, ny, nz, (C.c_uint * 3)(*start), (C.c_uint * 3)(*size), _p(out), C.byref(ms)))
    return out, ms.value


def probe_sample(vol, coords):
    """[n, 3] ints: read_imagei on a SIGNED_INT16 3-D image with {linear sampler + float coords, nearest sampler + float coords,
    linear sampler + int coords} at the float coordinates `coords` [n, 3] (x, y, z) — our own probe kernel"""
    vol = np.ascontiguousarray(vol, dtype=np.int16)
    nz, ny, nx = vol.shape
    c = np.zeros((len(coords), 4), dtype=np.float32)
    c[:, :3] = np.asarray(coords, dtype=np.float32)
    out = np.zeros((len(coords), 3), dtype=np.int32)
    _check(lib().ocl_probe_sample(_p(vol), nx, ny, nz, _p(c), len(coords), _p(out)))
    return out


class Scene:
    """reference_volume + env_map + renderer after flush_changes(): volume / env / frame images, the voxel cache, the JIT-compiled
    `render` kernel with `tf_src` (the generated is_event_gen text) prepended, and the SDF built by the reference's host loop."""

    def __init__(self, vol, env_rgba, tf_src, W, H):
        vol = np.ascontiguousarray(vol, dtype=np.int16)
        env = np.ascontiguousarray(env_rgba, dtype=np.uint8)
        self.nz, self.ny, self.nx = vol.shape
        self.W, self.H = W, H
        self.h = C.c_void_p()
        _check(lib().ocl_scene_create(_p(vol), self.nx, self.ny, self.nz, _p(env), env.shape[1], env.shape[0], tf_src.encode(), W, H,
                                      C.byref(self.h)))
        it = C.c_int(0)
        ms = (C.c_double * 3)()
        lib().ocl_scene_timings(self.h, C.byref(it), ms)
        self.sdf_iterations, self.sdf_ms, self.sdf_jit_ms, self.render_jit_ms = it.value, ms[0], ms[1], ms[2]
        self.reset()

    def reset(self):
        _check(lib().ocl_scene_reset_cache(self.h))

    def render(self, pos, direction, seeds, pull_every_frame=True, readback=True):
        """returns (last frame or None, host wall ms of the loop)"""
        seeds = (C.c_int32 * len(seeds))(*[int(s) for s in seeds])
        out = np.zeros((self.H, self.W, 4), dtype=np.uint8) if readback else None
        ms = C.c_double(0)
        _check(lib().ocl_scene_render(self.h, (C.c_float * 3)(*[float(v) for v in pos]), (C.c_float * 3)(*[float(v) for v in direction]),
                                      seeds, len(seeds), 1 if pull_every_frame else 0, _p(out) if readback else None, C.byref(ms)))
        return out, ms.value

    def cache(self):
        out = np.empty(self.nx * self.ny * self.nz * 4, dtype=np.uint16)
        _check(lib().ocl_scene_cache_download(self.h, _p(out)))
        return out

    def sdf(self):
        out = np.empty((self.nz, self.ny, self.nx), dtype=np.int8)
        _check(lib().ocl_scene_sdf_download(self.h, _p(out)))
        return out

    def close(self):
        if self.h:
            lib().ocl_scene_destroy(self.h)
            self.h = C.c_void_p()
