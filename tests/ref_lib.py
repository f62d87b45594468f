"""ctypes binding of oracle/_ref/libref.so — the reference's own OpenCL kernels compiled for the host CPU
(oracle/ref_build/).  TEST INFRASTRUCTURE: used by tests/ and by bench.py --impl reference only.

The library is built in the build container from /root/reference (which the GPU box does not have); the prebuilt .so
travels with the repository snapshot.  available() is False when it is missing."""
import ctypes as C
import os

import numpy as np

from oracle_lib import TfRect, tf_rects, _p

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(_ROOT, "oracle", "_ref", "libref.so")
_lib = None


def available():
    return os.path.exists(SO)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(SO)
        _lib.ref_sdf_build.restype = C.c_int
        _lib.ref_histogram.restype = C.c_int
        _lib.ref_tf_color_frame.restype = C.c_int
        _lib.ref_num_threads.restype = C.c_int
    return _lib


def num_threads():
    return lib().ref_num_threads()


def sdf_build(vol, tf, threads=0):
    vol = np.ascontiguousarray(vol, dtype=np.int16)
    nz, ny, nx = vol.shape
    out = np.empty(vol.shape, dtype=np.int8)
    arr, n = tf_rects(tf)
    it = lib().ref_sdf_build(_p(vol), nx, ny, nz, arr, n, _p(out), threads)
    return out, it


def fetch_stats(vol, threads=0):
    vol = np.ascontiguousarray(vol, dtype=np.int16)
    nz, ny, nx = vol.shape
    st = (C.c_int32 * 4)()
    lib().ref_fetch_stats(_p(vol), nx, ny, nz, st, threads)
    return list(st)


def histogram(vol, width, height, rng, threads=0):
    vol = np.ascontiguousarray(vol, dtype=np.int16)
    nz, ny, nx = vol.shape
    bins = np.zeros(width * height + height + 2, dtype=np.uint32)
    rc = lib().ref_histogram(_p(vol), nx, ny, nz, width, height, C.c_float(rng[0]), C.c_float(rng[1]), C.c_float(rng[2]),
                             C.c_float(rng[3]), _p(bins), threads)
    if rc != 0:
        raise ValueError("range would make the reference kernel index below 0")
    return bins[: width * height].copy()


def tf_color_frame(bins, width, height):
    b = np.zeros(width * height + height + 2, dtype=np.uint32)
    b[: width * height] = bins
    out = np.zeros((height, width, 4), dtype=np.uint8)
    n = lib().ref_tf_color_frame(_p(b), width, height, _p(out))
    return out, b[: width * height].copy(), n


def bilateral(vol, threads=0):
    vol = np.ascontiguousarray(vol, dtype=np.int16)
    nz, ny, nx = vol.shape
    out = np.empty_like(vol)
    lib().ref_bilateral(_p(vol), nx, ny, nz, _p(out), threads)
    return out


def clip(vol, start, size):
    vol = np.ascontiguousarray(vol, dtype=np.int16)
    nz, ny, nx = vol.shape
    out = np.empty((size[2], size[1], size[0]), dtype=np.int16)
    lib().ref_clip(_p(vol), nx, ny, nz, (C.c_int * 3)(*start), (C.c_int * 3)(*size), _p(out))
    return out


def image_filter2d(rgba, kernel_size, sigma, threads=0):
    """the reference's 2d_image_filter.cl bilateral_filter over w x h work-items; reads see the input frame"""
    rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
    h, w = rgba.shape[:2]
    out = np.empty_like(rgba)
    lib().ref_image_filter2d(_p(rgba), w, h, int(kernel_size), C.c_float(sigma), _p(out), threads)
    return out


class Renderer:
    """The reference `render` kernel over a CPU NDRange; token cap is the kernel's hard-coded 256."""

    def __init__(self, vol, env_rgba, tf, W, H, sdf):
        self.vol = np.ascontiguousarray(vol, dtype=np.int16)
        self.nz, self.ny, self.nx = self.vol.shape
        self.env = np.ascontiguousarray(env_rgba, dtype=np.uint8)
        self.tf, self.ntf = tf_rects(tf)
        self.W, self.H = W, H
        self.sdf = np.ascontiguousarray(sdf, dtype=np.int8)
        self.cache = np.zeros(self.vol.size * 4, dtype=np.uint16)
        self.reset()

    def reset(self):
        self.cache[:] = 0xFFFF
        lib().ref_buffer_reset(_p(self.cache), self.nx, self.ny, self.nz)

    def render_frame(self, cam_pos, cam_dir, seed, window=None, threads=1):
        frame = np.zeros((self.H, self.W, 4), dtype=np.uint8)
        x0, y0, x1, y1 = window if window else (0, 0, self.W, self.H)
        lib().ref_render_frame(_p(self.vol), self.nx, self.ny, self.nz, _p(self.sdf), _p(self.env), self.env.shape[1],
                               self.env.shape[0], self.tf, self.ntf, _p(self.cache), self.W, self.H, x0, y0, x1, y1,
                               (C.c_float * 3)(*[float(v) for v in cam_pos]), (C.c_float * 3)(*[float(v) for v in cam_dir]),
                               C.c_int32(seed), _p(frame), threads)
        return frame
