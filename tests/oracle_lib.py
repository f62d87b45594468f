"""ctypes binding of oracle/liboracle.so — the CPU restatement of the reference kernels.

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
The product (cl_volume_renderer_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ORACLE_DIR = os.path.join(_ROOT, "oracle")

TF_USE_GRADIENT = 1
TF_THRESHOLD = 2


class TfRect(C.Structure):
    _fields_ = [("min_v", C.c_float), ("max_v", C.c_float), ("min_g", C.c_float), ("max_g", C.c_float),
                ("flags", C.c_int32), ("rgba", C.c_int32 * 4)]


def tf_rects(specs):
    """specs: list of dicts {min_v,max_v,min_g,max_g,flags,rgba} -> ctypes array"""
    arr = (TfRect * max(len(specs), 1))()
    for i, s in enumerate(specs):
        arr[i].min_v = s.get("min_v", 0.0)
        arr[i].max_v = s.get("max_v", 0.0)
        arr[i].min_g = s.get("min_g", 0.0)
        arr[i].max_g = s.get("max_g", 0.0)
        arr[i].flags = s.get("flags", 0)
        for k in range(4):
            arr[i].rgba[k] = s.get("rgba", (0, 0, 0, 0))[k]
    return arr, len(specs)


def tf_threshold(k):
    return [{"min_v": float(k), "flags": TF_THRESHOLD}]


def tf_default():
    # ui.cpp:195 — rect(500,1200,0,4000), white, alpha 1.0; gradient clause omitted when it covers the stats range
    return [{"min_v": 500.0, "max_v": 1200.0, "min_g": 0.0, "max_g": 4000.0, "flags": 0, "rgba": (255, 255, 255, 255)}]


def _build():
    so = os.path.join(_ORACLE_DIR, "liboracle.so")
    src = os.path.join(_ORACLE_DIR, "oracle.cpp")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _ORACLE_DIR, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build())
        _lib.orc_hash.restype = C.c_uint32
        _lib.orc_hash.argtypes = [C.c_uint32]
        _lib.orc_sdf_build.restype = C.c_int
        _lib.orc_num_threads.restype = C.c_int
        _lib.orc_tf_color_frame.restype = C.c_int
    return _lib


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t)


def hash_u32(s):
    return int(lib().orc_hash(C.c_uint32(s & 0xFFFFFFFF)))


def rng_triple(seed, gx, gy):
    ra = (C.c_int32 * 3)()
    comp = (C.c_int32 * 3)()
    lib().orc_rng_triple(C.c_int32(seed), C.c_uint32(gx), C.c_uint32(gy), ra, comp)
    return list(ra), list(comp)


def hemisphere(normal, seed, rough, gx, gy):
    n = (C.c_float * 3)(*normal)
    o = (C.c_float * 3)()
    lib().orc_hemisphere(n, C.c_int32(seed), C.c_float(rough), C.c_uint32(gx), C.c_uint32(gy), o)
    return np.array(list(o), dtype=np.float32)


def camera_dir(alpha, beta):
    o = (C.c_float * 3)()
    lib().orc_camera_dir(C.c_double(alpha), C.c_double(beta), o)
    return np.array(list(o), dtype=np.float32)


def sdf_build(vol, tf):
    vol = np.ascontiguousarray(vol, dtype=np.int16)
    nz, ny, nx = vol.shape
    out = np.empty(vol.shape, dtype=np.int8)
    arr, n = tf_rects(tf)
    iters = lib().orc_sdf_build(_p(vol), nx, ny, nz, arr, n, _p(out))
    return out, iters


def fetch_stats(vol):
    vol = np.ascontiguousarray(vol, dtype=np.int16)
    nz, ny, nx = vol.shape
    st = (C.c_int32 * 4)()
    lib().orc_fetch_stats(_p(vol), nx, ny, nz, st)
    return list(st)


def histogram(vol, width, height, rng):
    vol = np.ascontiguousarray(vol, dtype=np.int16)
    nz, ny, nx = vol.shape
    bins = np.zeros(width * height, dtype=np.uint32)
    lib().orc_histogram(_p(vol), nx, ny, nz, width, height, C.c_float(rng[0]), C.c_float(rng[1]), C.c_float(rng[2]),
                        C.c_float(rng[3]), _p(bins))
    return bins


def tf_color_frame(bins, width, height):
    bins = np.array(bins, dtype=np.uint32, copy=True)
    out = np.zeros((height, width, 4), dtype=np.uint8)
    n = lib().orc_tf_color_frame(_p(bins), width, height, _p(out))
    return out, bins, n


def bilateral(vol):
    vol = np.ascontiguousarray(vol, dtype=np.int16)
    nz, ny, nx = vol.shape
    out = np.empty_like(vol)
    lib().orc_bilateral(_p(vol), nx, ny, nz, _p(out))
    return out


def clip(vol, start, size):
    vol = np.ascontiguousarray(vol, dtype=np.int16)
    nz, ny, nx = vol.shape
    out = np.empty((size[2], size[1], size[0]), dtype=np.int16)
    lib().orc_clip(_p(vol), nx, ny, nz, (C.c_int * 3)(*start), (C.c_int * 3)(*size), _p(out))
    return out


def image_filter2d(rgba, kernel_size, sigma):
    """2d_image_filter.cl bilateral_filter, literal (reads see the input frame)"""
    rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
    h, w = rgba.shape[:2]
    out = np.empty_like(rgba)
    lib().orc_image_filter2d(_p(rgba), w, h, int(kernel_size), C.c_float(sigma), _p(out))
    return out


def image_bilateral2d(rgba, kernel_size, sigma):
    """the corrected 2-D bilateral (our definition, oracle.cpp)"""
    rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
    h, w = rgba.shape[:2]
    out = np.empty_like(rgba)
    lib().orc_image_bilateral2d(_p(rgba), w, h, int(kernel_size), C.c_float(sigma), _p(out))
    return out


def set_sampling(mode):
    """0: NEAREST (default); 1: the linear filtering NVIDIA hardware applies to the reference's integer images (oracle.cpp
    hw_linear_fetch).  Applies to Renderer.render_frame from now on."""
    lib().orc_set_sampling(int(mode))


def hw_linear_fetch(vol, coords):
    vol = np.ascontiguousarray(vol, dtype=np.int16)
    nz, ny, nx = vol.shape
    c = np.ascontiguousarray(coords, dtype=np.float32)
    out = np.empty(len(c), dtype=np.int32)
    lib().orc_hw_linear_fetch(_p(vol), nx, ny, nz, _p(c), len(c), _p(out))
    return out


def fetch_stats_shipped(vol, edge):
    """fetch_stats with the hardware filter (oracle.cpp orc_fetch_stats_shipped); edge: what texel -1 reads under the sampler without
    an addressing mode (0 border colour, 1 nearest edge texel)"""
    vol = np.ascontiguousarray(vol, dtype=np.int16)
    nz, ny, nx = vol.shape
    st = (C.c_int32 * 4)()
    lib().orc_fetch_stats_shipped(_p(vol), nx, ny, nz, int(edge), st)
    return list(st)


def histogram_shipped(vol, edge, width, height, rng):
    vol = np.ascontiguousarray(vol, dtype=np.int16)
    nz, ny, nx = vol.shape
    bins = np.zeros(width * height, dtype=np.uint32)
    lib().orc_histogram_shipped(_p(vol), nx, ny, nz, int(edge), width, height, C.c_float(rng[0]), C.c_float(rng[1]), C.c_float(rng[2]),
                                C.c_float(rng[3]), _p(bins))
    return bins


def bilateral_shipped(vol):
    vol = np.ascontiguousarray(vol, dtype=np.int16)
    nz, ny, nx = vol.shape
    out = np.empty_like(vol)
    lib().orc_bilateral_shipped(_p(vol), nx, ny, nz, _p(out))
    return out


def quiet_cells(vol, tf):
    """per 2x2x2 cell: can an interpolated value there meet a TF clause?  1 = no (oracle.cpp orc_quiet_cells); (nz+1, ny+1, nx+1)"""
    vol = np.ascontiguousarray(vol, dtype=np.int16)
    nz, ny, nx = vol.shape
    arr, n = tf_rects(tf)
    q = np.zeros((nz + 1, ny + 1, nx + 1), dtype=np.uint8)
    lib().orc_quiet_cells(_p(vol), nx, ny, nz, arr, n, _p(q))
    return q


def quiet_cells27(vol, tf):
    """per voxel cell floor(p) (0..n inclusive): can a value interpolated anywhere in it meet a TF clause?  1 = no; covers the 3x3x3
    texels around the cell (oracle.cpp orc_quiet_cells27) — the flags the CUDA path packs into its step field"""
    vol = np.ascontiguousarray(vol, dtype=np.int16)
    nz, ny, nx = vol.shape
    arr, n = tf_rects(tf)
    q = np.zeros((nz + 1, ny + 1, nx + 1), dtype=np.uint8)
    lib().orc_quiet_cells27(_p(vol), nx, ny, nz, arr, n, _p(q))
    return q


def set_quiet_mode(mode):
    """0: flags per hardware cell (quiet_cells), 1: flags per voxel cell (quiet_cells27)"""
    lib().orc_set_quiet_mode(int(mode))


def set_quiet_cells(q):
    """install (or, with None, remove) the flags: linear event tests then count how many could be skipped and verify none is an event"""
    lib().orc_set_quiet_cells(_p(q) if q is not None else None)


def quiet_stats():
    st = (C.c_uint64 * 3)()
    lib().orc_quiet_stats(st)
    return {"event_tests": int(st[0]), "skippable": int(st[1]), "violations": int(st[2])}


def env_lookup(env_rgba, dirs):
    env_rgba = np.ascontiguousarray(env_rgba, dtype=np.uint8)
    h, w = env_rgba.shape[:2]
    dirs = np.ascontiguousarray(dirs, dtype=np.float32)
    n = dirs.shape[0]
    out = np.empty((n, 4), dtype=np.uint8)
    txy = np.empty((n, 2), dtype=np.int32)
    lib().orc_env_lookup(_p(env_rgba), w, h, _p(dirs), n, _p(out), _p(txy))
    return out, txy


def primary_ray(dims, W, H, x, y, cam_pos, cam_dir):
    d = (C.c_float * 3)()
    c = (C.c_float * 3)()
    ic = C.c_int()
    lib().orc_primary_ray(dims[0], dims[1], dims[2], W, H, x, y, (C.c_float * 3)(*cam_pos), (C.c_float * 3)(*cam_dir),
                          d, c, C.byref(ic))
    return np.array(list(d), np.float32), np.array(list(c), np.float32), bool(ic.value)


class Renderer:
    """Stateful CPU renderer: holds volume, SDF, env map, TF and the voxel cache."""

    def __init__(self, vol, env_rgba, tf, W, H, token_cap=256, sdf=None):
        self.vol = np.ascontiguousarray(vol, dtype=np.int16)
        self.nz, self.ny, self.nx = self.vol.shape
        self.env = np.ascontiguousarray(env_rgba, dtype=np.uint8)
        self.tf_spec = tf
        self.tf, self.ntf = tf_rects(tf)
        self.W, self.H = W, H
        self.token_cap = token_cap
        self.sdf = sdf if sdf is not None else sdf_build(self.vol, tf)[0]
        self.sdf = np.ascontiguousarray(self.sdf, dtype=np.int8)
        self.cache = np.zeros(self.vol.size * 4, dtype=np.uint16)
        self.counters = np.zeros(6, dtype=np.uint64)

    def reset(self):
        lib().orc_buffer_reset(_p(self.cache), self.nx, self.ny, self.nz)

    def render_frame(self, cam_pos, cam_dir, seed, window=None, want_frame=True):
        frame = np.zeros((self.H, self.W, 4), dtype=np.uint8) if want_frame else None
        x0, y0, x1, y1 = window if window else (0, 0, self.W, self.H)
        lib().orc_render_frame(_p(self.vol), self.nx, self.ny, self.nz, _p(self.sdf), _p(self.env), self.env.shape[1],
                               self.env.shape[0], self.tf, self.ntf, _p(self.cache), self.token_cap, self.W, self.H,
                               x0, y0, x1, y1, (C.c_float * 3)(*[float(v) for v in cam_pos]),
                               (C.c_float * 3)(*[float(v) for v in cam_dir]), C.c_int32(seed),
                               _p(frame) if frame is not None else None, _p(self.counters))
        return frame


def render_frame_immediate(r, cam_pos, cam_dir, seed, window=None):
    """single-phase, row-major sequential execution (see orc_render_frame_immediate) on Renderer r's state"""
    frame = np.zeros((r.H, r.W, 4), dtype=np.uint8)
    x0, y0, x1, y1 = window if window else (0, 0, r.W, r.H)
    lib().orc_render_frame_immediate(_p(r.vol), r.nx, r.ny, r.nz, _p(r.sdf), _p(r.env), r.env.shape[1], r.env.shape[0],
                                     r.tf, r.ntf, _p(r.cache), r.token_cap, r.W, r.H, x0, y0, x1, y1,
                                     (C.c_float * 3)(*[float(v) for v in cam_pos]),
                                     (C.c_float * 3)(*[float(v) for v in cam_dir]), C.c_int32(seed), _p(frame))
    return frame


def num_threads():
    return lib().orc_num_threads()
