"""ctypes binding of libvr.so (include/vr.h) — the Python view of the C-ABI used by tests/ and bench.py.

The product is the CUDA library; this module only marshals numpy buffers into it.  There is no CPU fallback:
if libvr.so is missing, or no B200 is visible, the calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# VR_LIB: tools/ and the A/B tests point this harness at tools/ab/libvr_ab.so (the -DVR_AB build); the product library is libvr.so
LIB_PATH = os.environ.get("VR_LIB") or os.path.join(_HERE, "libvr.so")
VR_COMM_ID_BYTES = 128

VR_TF_USE_GRADIENT = 1
VR_TF_THRESHOLD = 2
VR_TF_MAX_RECTS = 16
VR_SAMPLING_NEAREST = 0    # the filter OpenCL defines for integer images (default)
VR_SAMPLING_HW_LINEAR = 1  # the reference's samplers as NVIDIA hardware executes them (texture-unit interpolation)
VR_FILTER2D_REFERENCE = 0  # 2d_image_filter.cl as written
VR_FILTER2D_BILATERAL = 1  # the corrected bilateral


class VrError(RuntimeError):
    pass


class TfRect(C.Structure):
    _fields_ = [("min_v", C.c_float), ("max_v", C.c_float), ("min_g", C.c_float), ("max_g", C.c_float),
                ("flags", C.c_int32), ("rgba", C.c_int32 * 4)]

    def as_dict(self):
        return {"min_v": self.min_v, "max_v": self.max_v, "min_g": self.min_g, "max_g": self.max_g,
                "flags": self.flags, "rgba": tuple(self.rgba)}


# every symbol include/vr.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "vr_last_error": (C.c_char_p, []),
    "vr_ctx_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "vr_ctx_destroy": (None, [_P]),
    "vr_ctx_synchronize": (C.c_int, [_P]),
    "vr_ctx_stream": (_P, [_P]),
    "vr_ctx_launch_count": (C.c_uint64, [_P]),
    "vr_ctx_array_count": (C.c_int, [_P]),
    "vr_tf_parse": (C.c_int, [C.c_char_p, C.POINTER(TfRect), C.c_int, C.POINTER(C.c_int)]),
    "vr_tf_format": (C.c_int, [C.POINTER(TfRect), C.c_int, C.c_char_p, C.c_size_t]),
    "vr_volume_upload": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "vr_volume_upload_async": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "vr_volume_wait": (C.c_int, [_P]),
    "vr_volume_upload_device": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "vr_sdf_checksum": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "vr_volume_checksum": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "vr_volume_device_ptr": (_P, [_P]),
    "vr_volume_destroy": (None, [_P]),
    "vr_volume_stats": (C.c_int, [_P, C.POINTER(C.c_int32)]),
    "vr_volume_set_value_clip": (C.c_int, [_P, C.c_int, C.c_int]),
    "vr_volume_set_gradient_clip": (C.c_int, [_P, C.c_int, C.c_int]),
    "vr_volume_clipped_stats": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "vr_volume_dims": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "vr_volume_clip": (C.c_int, [_P, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "vr_volume_filter": (C.c_int, [_P]),
    "vr_volume_download": (C.c_int, [_P, _P]),
    "vr_histogram": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_float), _P]),
    "vr_volume_upload_slab": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "vr_volume_download_planes": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "vr_sdf_slab_create": (C.c_int, [_P, _P, C.POINTER(TfRect), C.c_int, C.c_int, C.POINTER(_P)]),
    "vr_sdf_slab_advance": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int)]),
    "vr_sdf_slab_bits": (_P, [_P]),
    "vr_sdf_slab_plane_words": (C.c_size_t, [_P]),
    "vr_sdf_slab_mark_imported": (C.c_int, [_P]),
    "vr_sdf_slab_finished": (C.c_int, [_P]),
    "vr_sdf_slab_download": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "vr_sdf_slab_destroy": (None, [_P]),
    "vr_envmap_bind": (C.c_int, [_P, _P, C.c_int, C.c_int, C.POINTER(_P)]),
    "vr_envmap_destroy": (None, [_P]),
    "vr_sdf_build": (C.c_int, [_P, _P, C.POINTER(TfRect), C.c_int, C.POINTER(_P)]),
    "vr_sdf_destroy": (None, [_P]),
    "vr_sdf_download": (C.c_int, [_P, _P]),
    "vr_sdf_levels": (C.c_int, [_P]),
    "vr_renderer_create": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(_P)]),
    "vr_renderer_destroy": (None, [_P]),
    "vr_renderer_set_scene": (C.c_int, [_P, _P, _P]),
    "vr_renderer_set_tf": (C.c_int, [_P, C.POINTER(TfRect), C.c_int]),
    "vr_renderer_set_tf_code": (C.c_int, [_P, C.c_char_p]),
    "vr_renderer_flush": (C.c_int, [_P]),
    "vr_renderer_last_flush_kept_fields": (C.c_int, [_P]),
    "vr_renderer_reset_cache": (C.c_int, [_P]),
    "vr_render_frame": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int32, _P]),
    "vr_render_frames": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int32), C.c_int, _P]),
    "vr_render_tf": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "vr_renderer_set_sampling": (C.c_int, [_P, C.c_int]),
    "vr_volume_set_sampling": (C.c_int, [_P, C.c_int]),
    "vr_renderer_filter_frame": (C.c_int, [_P, C.c_int, C.c_float, C.c_int, _P]),
    "vr_image_filter": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, _P]),
    "vr_renderer_host_frame": (_P, [_P]),
    "vr_cache_download": (C.c_int, [_P, _P]),
    "vr_renderer_sdf": (_P, [_P]),
    "vr_renderer_hit_download": (C.c_int, [_P, _P]),
    "vr_cache_download_at": (C.c_int, [_P, _P, C.c_size_t, _P]),
    "vr_renderer_set_token_cap": (C.c_int, [_P, C.c_int]),
    "vr_renderer_set_rows": (C.c_int, [_P, C.c_int, C.c_int]),
    "vr_renderer_cache_device_ptr": (_P, [_P]),
    "vr_renderer_cache_bytes": (C.c_size_t, [_P]),
    "vr_renderer_frame_device_ptr": (_P, [_P]),
    "vr_renderer_resolve": (C.c_int, [_P, _P]),
    "vr_renderer_xchg_gather": (C.c_int, [_P]),
    "vr_renderer_xchg_scatter": (C.c_int, [_P]),
    "vr_renderer_xchg_device_ptr": (_P, [_P]),
    "vr_renderer_xchg_bytes": (C.c_size_t, [_P]),
    "vr_renderer_enable_counters": (C.c_int, [_P, C.c_int]),
    "vr_renderer_counters": (C.c_int, [_P, C.POINTER(C.c_uint64), C.c_int]),
    "vr_renderer_set_trace_mode": (C.c_int, [_P, C.c_int]),
    "vr_renderer_set_primary_reuse": (C.c_int, [_P, C.c_int]),
    "vr_renderer_enable_timing": (C.c_int, [_P, C.c_int]),
    "vr_renderer_kernel_times": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_int]),
    "vr_renderer_filtered_device_ptr": (_P, [_P]),
    "vr_renderer_set_tuning": (C.c_int, [_P, C.c_char_p, C.c_int]),
    "vr_renderer_quiet_download": (C.c_int, [_P, _P]),
    "vr_debug_rng_dump": (C.c_int, [_P, _P, _P, _P, C.c_int, _P, _P, _P]),
    "vr_debug_linear_fetch": (C.c_int, [_P, _P, C.c_int, _P]),
    "vr_comm_unique_id": (C.c_int, [_P]),
    "vr_comm_init": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "vr_comm_destroy": (None, [_P]),
    "vr_comm_rank": (C.c_int, [_P]),
    "vr_comm_size": (C.c_int, [_P]),
    "vr_comm_barrier": (C.c_int, [_P]),
    "vr_comm_allreduce_host": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int]),
    "vr_comm_slab": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "vr_cache_allreduce": (C.c_int, [_P, _P]),
    "vr_renderer_set_row_blocks": (C.c_int, [_P, C.c_int, C.c_int, C.c_int]),
    "vr_frame_allgather": (C.c_int, [_P, _P]),
    "vr_volume_upload_sharded": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "vr_volume_upload_sharded_async": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "vr_sdf_build_sharded": (C.c_int, [_P, _P, C.POINTER(TfRect), C.c_int, C.POINTER(_P)]),
    "vr_renderer_set_sharded_build": (C.c_int, [_P, C.c_int]),
    "vr_sdf_build_slab_only": (C.c_int, [_P, _P, C.POINTER(TfRect), C.c_int, C.POINTER(_P)]),
    "vr_histogram_sharded": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_float), _P]),
    "vr_volume_filter_sharded": (C.c_int, [_P]),
}

_lib = None


def lib():
    """Loads libvr.so; raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VrError(f"{LIB_PATH} is missing: build it with `make` (there is no CPU fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def _check(status):
    if status != 0:
        raise VrError(f"vr status {status}: {lib().vr_last_error().decode(errors='replace')}")


def make_rects(specs):
    arr = (TfRect * max(len(specs), 1))()
    for i, s in enumerate(specs):
        arr[i].min_v = s.get("min_v", 0.0)
        arr[i].max_v = s.get("max_v", 0.0)
        arr[i].min_g = s.get("min_g", 0.0)
        arr[i].max_g = s.get("max_g", 0.0)
        arr[i].flags = s.get("flags", 0)
        col = s.get("rgba", (0, 0, 0, 0))
        for k in range(4):
            arr[i].rgba[k] = col[k]
    return arr, len(specs)


def tf_parse(src):
    out = (TfRect * VR_TF_MAX_RECTS)()
    n = C.c_int(0)
    _check(lib().vr_tf_parse(src.encode(), out, VR_TF_MAX_RECTS, C.byref(n)))
    return [out[i].as_dict() for i in range(n.value)]


def tf_format(specs):
    arr, n = make_rects(specs)
    buf = C.create_string_buffer(8192)
    _check(lib().vr_tf_format(arr, n, buf, 8192))
    return buf.value.decode()


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


class Context:
    def __init__(self, device=0):
        self.h = _P()
        _check(lib().vr_ctx_create(device, C.byref(self.h)))

    def close(self):
        if self.h:
            lib().vr_ctx_destroy(self.h)
            self.h = _P()

    def synchronize(self):
        _check(lib().vr_ctx_synchronize(self.h))

    @property
    def stream(self):
        return lib().vr_ctx_stream(self.h)

    def image_filter(self, rgba, kernel_size, sigma, mode=VR_FILTER2D_REFERENCE):
        """2d_image_filter.cl bilateral_filter(frame, kernel_size, sigma) on a host RGBA8 image [h, w, 4]"""
        rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
        assert rgba.ndim == 3 and rgba.shape[2] == 4
        out = np.empty_like(rgba)
        _check(lib().vr_image_filter(self.h, _vp(rgba), rgba.shape[1], rgba.shape[0], int(kernel_size), C.c_float(sigma),
                                     int(mode), _vp(out)))
        return out

    @property
    def launches(self):
        return int(lib().vr_ctx_launch_count(self.h))

    @property
    def array_count(self):
        return int(lib().vr_ctx_array_count(self.h))

    # ---- multi-GPU: NCCL communicator behind the C-ABI (include/vr.h) ----
    @staticmethod
    def comm_unique_id():
        buf = (C.c_uint8 * VR_COMM_ID_BYTES)()
        _check(lib().vr_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, rank, nranks, uid):
        buf = (C.c_uint8 * VR_COMM_ID_BYTES).from_buffer_copy(uid)
        _check(lib().vr_comm_init(self.h, rank, nranks, buf))

    @property
    def comm_rank(self):
        return lib().vr_comm_rank(self.h)

    @property
    def comm_size(self):
        return lib().vr_comm_size(self.h)

    def comm_barrier(self):
        _check(lib().vr_comm_barrier(self.h))

    def comm_allreduce(self, values, op="sum"):
        """small host array (int32 / uint32 / float64) reduced over the ranks in place; returns it"""
        a = np.ascontiguousarray(values)
        dt = {np.dtype(np.int32): 0, np.dtype(np.uint32): 1, np.dtype(np.float64): 2}[a.dtype]
        _check(lib().vr_comm_allreduce_host(self.h, _vp(a), a.size, dt, {"sum": 0, "min": 1, "max": 2}[op]))
        return a

    def comm_slab(self, nz, rank=None):
        z0, z1 = C.c_int(0), C.c_int(0)
        _check(lib().vr_comm_slab(self.h, nz, self.comm_rank if rank is None else rank, C.byref(z0), C.byref(z1)))
        return z0.value, z1.value

    def rng_dump(self, seeds, gids, normal_rough):
        """device RNG known answers: -> (ra [n,3] int32, comp [n,3] int32, dir [n,3] float32)"""
        seeds = np.ascontiguousarray(seeds, dtype=np.int32)
        gids = np.ascontiguousarray(gids, dtype=np.uint32).reshape(-1, 2)
        nr = np.ascontiguousarray(normal_rough, dtype=np.float32).reshape(-1, 4)
        n = seeds.size
        ra, comp, d = np.empty((n, 3), np.int32), np.empty((n, 3), np.int32), np.empty((n, 3), np.float32)
        _check(lib().vr_debug_rng_dump(self.h, _vp(seeds), _vp(gids), _vp(nr), n, _vp(ra), _vp(comp), _vp(d)))
        return ra, comp, d


class Volume:
    """reference_volume (app/reference_volume.hpp:27-63)"""

    @classmethod
    def from_device(cls, ctx, device_ptr, nx, ny, nz):
        """voxels already in the context's device memory (vr_volume_upload_device)"""
        self = cls.__new__(cls)
        self.ctx, self.h, self._keep = ctx, _P(), None
        _check(lib().vr_volume_upload_device(ctx.h, C.c_void_p(device_ptr), nx, ny, nz, C.byref(self.h)))
        return self

    def checksum(self):
        c = C.c_uint64(0)
        _check(lib().vr_volume_checksum(self.h, C.byref(c)))
        return int(c.value)

    def __init__(self, ctx, voxels, interior=None, async_upload=False, sharded_dims=None):
        """interior=(z_lo, z_hi): `voxels` is a z-slab with halo planes; stats / histogram cover planes [z_lo, z_hi) only.
        async_upload: return at once (vr_volume_upload_async); `voxels` must not change until wait() / first use.
        sharded_dims=(nx, ny, nz): `voxels` holds only this rank's planes ctx.comm_slab(nz) of that volume (vr_volume_upload_sharded)."""
        voxels = np.ascontiguousarray(voxels, dtype=np.int16)
        assert voxels.ndim == 3, "volume must be [nz, ny, nx]"
        nz, ny, nx = voxels.shape
        self.ctx = ctx
        self.h = _P()
        self._keep = voxels
        if sharded_dims is not None:
            gx, gy, gz = sharded_dims
            z0, z1 = ctx.comm_slab(gz)
            assert (nx, ny, nz) == (gx, gy, z1 - z0), "sharded upload: pass exactly this rank's planes"
            fn = lib().vr_volume_upload_sharded_async if async_upload else lib().vr_volume_upload_sharded
            _check(fn(ctx.h, _vp(voxels), gx, gy, gz, C.byref(self.h)))
        elif async_upload:
            assert interior is None
            _check(lib().vr_volume_upload_async(ctx.h, _vp(voxels), nx, ny, nz, C.byref(self.h)))
        elif interior is None:
            _check(lib().vr_volume_upload(ctx.h, _vp(voxels), nx, ny, nz, C.byref(self.h)))
        else:
            _check(lib().vr_volume_upload_slab(ctx.h, _vp(voxels), nx, ny, nz, interior[0], interior[1], C.byref(self.h)))

    def wait(self):
        _check(lib().vr_volume_wait(self.h))
        self._keep = None

    def download_planes(self, z0, nplanes):
        nx, ny, _ = self.dims()
        out = np.empty((nplanes, ny, nx), dtype=np.int16)
        _check(lib().vr_volume_download_planes(self.h, z0, nplanes, _vp(out)))
        return out

    def close(self):
        if self.h:
            lib().vr_volume_destroy(self.h)
            self.h = _P()

    def stats(self):
        s = (C.c_int32 * 4)()
        _check(lib().vr_volume_stats(self.h, s))
        return list(s)

    def set_sampling(self, mode):
        """VR_SAMPLING_HW_LINEAR: stats (recomputed now), histogram / TF image and the bilateral filter read the volume the way NVIDIA
        hardware serves the reference's CLK_FILTER_LINEAR samplers"""
        _check(lib().vr_volume_set_sampling(self.h, mode))

    def set_value_clip(self, lo, hi):
        _check(lib().vr_volume_set_value_clip(self.h, lo, hi))

    def set_gradient_clip(self, lo, hi):
        _check(lib().vr_volume_set_gradient_clip(self.h, lo, hi))

    def clipped_stats(self):
        s = (C.c_float * 4)()
        _check(lib().vr_volume_clipped_stats(self.h, s))
        return list(s)

    def dims(self):
        d = (C.c_int * 3)()
        _check(lib().vr_volume_dims(self.h, d))
        return list(d)

    def clip(self, mn, mx):
        _check(lib().vr_volume_clip(self.h, (C.c_uint32 * 3)(*mn), (C.c_uint32 * 3)(*mx)))

    def filter(self):
        _check(lib().vr_volume_filter(self.h))

    def filter_sharded(self):
        _check(lib().vr_volume_filter_sharded(self.h))

    def histogram_sharded(self, width, height, rng):
        bins = np.empty(width * height, dtype=np.uint32)
        _check(lib().vr_histogram_sharded(self.h, width, height, (C.c_float * 4)(*rng), _vp(bins)))
        return bins

    def download(self):
        nx, ny, nz = self.dims()
        out = np.empty((nz, ny, nx), dtype=np.int16)
        _check(lib().vr_volume_download(self.h, _vp(out)))
        return out

    def histogram(self, width, height, rng):
        bins = np.empty(width * height, dtype=np.uint32)
        _check(lib().vr_histogram(self.h, width, height, (C.c_float * 4)(*rng), _vp(bins)))
        return bins


class EnvMap:
    def __init__(self, ctx, rgba):
        rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
        assert rgba.ndim == 3 and rgba.shape[2] == 4
        self.h = _P()
        _check(lib().vr_envmap_bind(ctx.h, _vp(rgba), rgba.shape[1], rgba.shape[0], C.byref(self.h)))

    def close(self):
        if self.h:
            lib().vr_envmap_destroy(self.h)
            self.h = _P()


class Sdf:
    """signed_distance_field (app/signed_distance_field.hpp:5-12)"""

    def __init__(self, ctx, volume, tf_specs, sharded=False):
        """sharded: False single GPU; True z-slab build + gather; "slab" z-slab build only (own planes valid)"""
        arr, n = make_rects(tf_specs)
        self.h = _P()
        self.dims = volume.dims()
        fn = lib().vr_sdf_build_slab_only if sharded == "slab" else (lib().vr_sdf_build_sharded if sharded else lib().vr_sdf_build)
        _check(fn(ctx.h, volume.h, arr, n, C.byref(self.h)))

    def close(self):
        if self.h:
            lib().vr_sdf_destroy(self.h)
            self.h = _P()

    def download(self):
        nx, ny, nz = self.dims
        out = np.empty((nz, ny, nx), dtype=np.int8)
        _check(lib().vr_sdf_download(self.h, _vp(out)))
        return out

    @property
    def levels(self):
        return lib().vr_sdf_levels(self.h)

    def checksum(self):
        c = C.c_uint64(0)
        _check(lib().vr_sdf_checksum(self.h, C.byref(c)))
        return int(c.value)


class SdfSlab:
    """SDF of a z-slab with halo planes, advanced level by level (include/vr.h, z-slab sharding); driver: parallel.py"""

    def __init__(self, ctx, ext_volume, tf_specs, max_it_global):
        arr, n = make_rects(tf_specs)
        self.ctx = ctx
        self.h = _P()
        self.dims = ext_volume.dims()
        _check(lib().vr_sdf_slab_create(ctx.h, ext_volume.h, arr, n, max_it_global, C.byref(self.h)))

    def advance(self, nlevels):
        done = C.c_int(0)
        _check(lib().vr_sdf_slab_advance(self.h, nlevels, C.byref(done)))
        return done.value

    @property
    def finished(self):
        return bool(lib().vr_sdf_slab_finished(self.h))

    @property
    def bits_ptr(self):
        return lib().vr_sdf_slab_bits(self.h)

    @property
    def plane_words(self):
        return int(lib().vr_sdf_slab_plane_words(self.h))

    def mark_imported(self):
        _check(lib().vr_sdf_slab_mark_imported(self.h))

    def download(self, z0, nplanes):
        nx, ny, nz = self.dims
        out = np.empty((nplanes, ny, nx), dtype=np.int8)
        _check(lib().vr_sdf_slab_download(self.h, self.ctx.h, nx, ny, nz, z0, nplanes, _vp(out)))
        return out

    def close(self):
        if self.h:
            lib().vr_sdf_slab_destroy(self.h)
            self.h = _P()


class Renderer:
    """renderer : frame_emitter (app/renderer.hpp:10-29)"""

    def __init__(self, ctx, width, height):
        self.ctx = ctx
        self.W, self.H = width, height
        self.h = _P()
        self.volume = None
        _check(lib().vr_renderer_create(ctx.h, width, height, C.byref(self.h)))

    def close(self):
        if self.h:
            lib().vr_renderer_destroy(self.h)
            self.h = _P()

    def image_set(self, volume, env):
        self.volume = volume
        self.env = env
        _check(lib().vr_renderer_set_scene(self.h, volume.h, env.h))

    def set_tf(self, tf_specs):
        arr, n = make_rects(tf_specs)
        _check(lib().vr_renderer_set_tf(self.h, arr, n))

    def next_event_code_set(self, cl_code):
        _check(lib().vr_renderer_set_tf_code(self.h, cl_code.encode()))

    def flush_changes(self):
        _check(lib().vr_renderer_flush(self.h))

    @property
    def last_flush_kept_fields(self):
        return bool(lib().vr_renderer_last_flush_kept_fields(self.h))

    def reset_cache(self):
        _check(lib().vr_renderer_reset_cache(self.h))

    def host_frame(self):
        """numpy view of the renderer-owned pinned frame buffer"""
        p = lib().vr_renderer_host_frame(self.h)
        buf = (C.c_uint8 * (self.W * self.H * 4)).from_address(p)
        return np.frombuffer(buf, dtype=np.uint8).reshape(self.H, self.W, 4)

    def render_frame(self, pos, direction, seed, readback=True, out=None):
        if readback:
            if out is None:
                out = np.empty((self.H, self.W, 4), dtype=np.uint8)
            _check(lib().vr_render_frame(self.h, _f3(pos), _f3(direction), C.c_int32(seed), _vp(out)))
            return out
        _check(lib().vr_render_frame(self.h, _f3(pos), _f3(direction), C.c_int32(seed), None))
        return None

    def render_frames(self, pos, direction, seeds, readback=True, out=None):
        seeds = (C.c_int32 * len(seeds))(*[int(s) for s in seeds])
        if readback and out is None:
            out = np.empty((self.H, self.W, 4), dtype=np.uint8)
        _check(lib().vr_render_frames(self.h, _f3(pos), _f3(direction), seeds, len(seeds),
                                      _vp(out) if readback else None))
        return out if readback else None

    def resolve(self, readback=True):
        out = np.empty((self.H, self.W, 4), dtype=np.uint8) if readback else None
        _check(lib().vr_renderer_resolve(self.h, _vp(out) if readback else None))
        return out

    def filter_frame(self, kernel_size, sigma, mode=VR_FILTER2D_REFERENCE, readback=True):
        """2d_image_filter.cl over the renderer's current device frame"""
        out = np.empty((self.H, self.W, 4), dtype=np.uint8) if readback else None
        _check(lib().vr_renderer_filter_frame(self.h, int(kernel_size), C.c_float(sigma), int(mode),
                                              _vp(out) if readback else None))
        return out

    def render_tf(self, width, height):
        out = np.empty((height, width, 4), dtype=np.uint8)
        _check(lib().vr_render_tf(self.h, width, height, _vp(out)))
        return out

    def cache_download(self):
        n = lib().vr_renderer_cache_bytes(self.h) // 2
        out = np.empty(n, dtype=np.uint16)
        _check(lib().vr_cache_download(self.h, _vp(out)))
        return out

    def hit_download(self):
        out = np.empty((self.H, self.W), dtype=np.uint32)
        _check(lib().vr_renderer_hit_download(self.h, _vp(out)))
        return out

    def cache_download_at(self, voxels):
        voxels = np.ascontiguousarray(voxels, dtype=np.uint32)
        out = np.empty((voxels.size, 4), dtype=np.uint16)
        _check(lib().vr_cache_download_at(self.h, _vp(voxels), voxels.size, _vp(out)))
        return out

    def sdf_download(self):
        nx, ny, nz = self.volume.dims()
        out = np.empty((nz, ny, nx), dtype=np.int8)
        _check(lib().vr_sdf_download(lib().vr_renderer_sdf(self.h), _vp(out)))
        return out

    def set_token_cap(self, cap):
        _check(lib().vr_renderer_set_token_cap(self.h, cap))

    def set_rows(self, y0, y1):
        _check(lib().vr_renderer_set_rows(self.h, y0, y1))

    @property
    def cache_device_ptr(self):
        return lib().vr_renderer_cache_device_ptr(self.h)

    @property
    def cache_bytes(self):
        return int(lib().vr_renderer_cache_bytes(self.h))

    @property
    def frame_device_ptr(self):
        return lib().vr_renderer_frame_device_ptr(self.h)

    def enable_counters(self, on=True):
        _check(lib().vr_renderer_enable_counters(self.h, 1 if on else 0))

    def counters(self, reset=False):
        c = (C.c_uint64 * 6)()
        _check(lib().vr_renderer_counters(self.h, c, 1 if reset else 0))
        return dict(zip(["steps", "normals", "env", "primary_hits", "admitted", "samples"], [int(v) for v in c]))

    def enable_timing(self, on=True):
        _check(lib().vr_renderer_enable_timing(self.h, 1 if on else 0))

    def kernel_times(self, reset=True):
        """(trace_ms_total, resolve_ms_total, frames) measured with CUDA events on the launching stream"""
        ms = (C.c_double * 2)()
        n = C.c_int(0)
        _check(lib().vr_renderer_kernel_times(self.h, ms, C.byref(n), 1 if reset else 0))
        return ms[0], ms[1], n.value

    def xchg_gather(self):
        _check(lib().vr_renderer_xchg_gather(self.h))

    def xchg_scatter(self):
        _check(lib().vr_renderer_xchg_scatter(self.h))

    @property
    def xchg_device_ptr(self):
        return lib().vr_renderer_xchg_device_ptr(self.h)

    @property
    def xchg_bytes(self):
        return int(lib().vr_renderer_xchg_bytes(self.h))

    def set_sampling(self, mode):
        """takes effect at the next flush_changes()"""
        _check(lib().vr_renderer_set_sampling(self.h, mode))

    def set_tuning(self, key, value):
        _check(lib().vr_renderer_set_tuning(self.h, key.encode(), int(value)))

    def quiet_download(self):
        """hw-linear step field: quiet-octant byte per voxel cell [nz, ny, nx]"""
        nx, ny, nz = self.volume.dims()
        out = np.empty((nz, ny, nx), dtype=np.uint8)
        _check(lib().vr_renderer_quiet_download(self.h, _vp(out)))
        return out

    def linear_fetch(self, coords):
        """hw-linear value at float positions [n,3] through the renderer's texture (vr_debug_linear_fetch)"""
        xyz = np.ascontiguousarray(coords, dtype=np.float32).reshape(-1, 3)
        out = np.empty(xyz.shape[0], dtype=np.int32)
        _check(lib().vr_debug_linear_fetch(self.h, _vp(xyz), xyz.shape[0], _vp(out)))
        return out

    def cache_allreduce(self, readback=False, out=None):
        """spp split: sum the touched cache entries over the ranks of the context's communicator and resolve"""
        if readback and out is None:
            out = np.empty((self.H, self.W, 4), dtype=np.uint8)
        _check(lib().vr_cache_allreduce(self.h, _vp(out) if readback else None))
        return out if readback else None

    def set_row_blocks(self, block_rows, rank, nranks):
        _check(lib().vr_renderer_set_row_blocks(self.h, block_rows, rank, nranks))

    def frame_allgather(self, readback=False, out=None):
        if readback and out is None:
            out = np.empty((self.H, self.W, 4), dtype=np.uint8)
        _check(lib().vr_frame_allgather(self.h, _vp(out) if readback else None))
        return out if readback else None

    def set_sharded_build(self, on=True):
        _check(lib().vr_renderer_set_sharded_build(self.h, 1 if on else 0))

    def set_trace_mode(self, mode):
        _check(lib().vr_renderer_set_trace_mode(self.h, mode))

    def set_primary_reuse(self, level):
        _check(lib().vr_renderer_set_primary_reuse(self.h, level))
