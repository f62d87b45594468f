"""Synthetic inputs for tests and the bench.  The reference ships no procedural volume; these are ours.

synth_ct(n)        `short` CT-like volume, x fastest ([nz,ny,nx] numpy), values 0..4095: hash-noise background (0..60),
                   solid ellipsoids and hollow shells with densities 600..1500 and a 3-voxel ramp at surfaces.
synth_env(w,h)     RGBA8 environment map: sky gradient + sun disc + ground tint.
GLIBC_RAND         first values of glibc rand() from its default state — the seeds renderer.cpp:142 would draw.
camera_dir(a,b)    Position3D(a, b, 0, {1,0,0}) (common.hpp:7-12), the view direction renderer.cpp:140 builds.
"""
import math

import numpy as np

# std::rand() sequence with the default seed (renderer.cpp:142 never calls srand)
GLIBC_RAND_HEAD = [1804289383, 846930886, 1681692777, 1714636915, 1957747793, 424238335]


def glibc_rand(n):
    """glibc TYPE_3 random(): r[i] = r[i-3] + r[i-31], output (r[i] >> 1); default seed 1."""
    r = [0] * (344 + n)
    r[0] = 1
    for i in range(1, 31):
        r[i] = (16807 * r[i - 1]) % 2147483647
    for i in range(31, 34):
        r[i] = r[i - 31]
    for i in range(34, 344 + n):
        r[i] = (r[i - 31] + r[i - 3]) & 0xFFFFFFFF
    return [(r[344 + k] & 0xFFFFFFFF) >> 1 for k in range(n)]


def camera_dir(alpha, beta):
    # gamma = 0, base = (1,0,0): only the first column of the rotation survives
    v = np.array([math.cos(alpha) * math.cos(beta), -math.sin(beta), math.sin(alpha) * math.cos(beta)], dtype=np.float32)
    length = np.float32(math.sqrt(float(v[0]) ** 2 + float(v[1]) ** 2 + float(v[2]) ** 2))
    return (v / length).astype(np.float32)


DEFAULT_LOOK = (0.9, 6.183)  # ui.cpp:178


def default_camera(n):
    """pos (-200,200,-200) scaled with the volume (ui.cpp:178 is tuned for ~256^3), default look angles"""
    s = n / 256.0
    return np.array([-200.0 * s, 200.0 * s, -200.0 * s], dtype=np.float32), camera_dir(*DEFAULT_LOOK)


def closeup_camera(n):
    """A second, harder view for the bench: camera 0.9*n in front of the volume centre, looking slightly down, so that
    about half of the pixels are shaded (the default UI view shades < 10 %)."""
    d = camera_dir(0.785398, 0.35)
    c = np.array([n / 2.0, n / 2.0, n / 2.0], dtype=np.float32)
    return (c - d * np.float32(0.9 * n)).astype(np.float32), d


class _SplitMix64:
    def __init__(self, seed):
        self.s = seed & 0xFFFFFFFFFFFFFFFF

    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)

    def uniform(self, lo, hi):
        return lo + (hi - lo) * (self.next() >> 11) / float(1 << 53)


def _noise(nz, ny, nx, z0):
    """integer hash of the voxel index -> 0..60"""
    z = np.arange(z0, z0 + nz, dtype=np.uint32)[:, None, None]
    y = np.arange(ny, dtype=np.uint32)[None, :, None]
    x = np.arange(nx, dtype=np.uint32)[None, None, :]
    h = (x * np.uint32(73856093)) ^ (y * np.uint32(19349663)) ^ (z * np.uint32(83492791))
    h ^= h >> np.uint32(13)
    h = h * np.uint32(0x5BD1E995)
    h ^= h >> np.uint32(15)
    return (h % np.uint32(61)).astype(np.int16)


def synth_ct(n, seed=0, dims=None, scale_to_u8=False):
    """CT-like `short` volume [nz,ny,nx].  dims overrides the cube (nx,ny,nz)."""
    nx, ny, nz = dims if dims else (n, n, n)
    rng = _SplitMix64(0xC0FFEE + n + seed)
    vol = np.empty((nz, ny, nx), dtype=np.int16)
    slab = max(1, min(nz, (1 << 24) // max(nx * ny, 1)))
    for z0 in range(0, nz, slab):
        vol[z0:z0 + slab] = _noise(min(slab, nz - z0), ny, nx, z0)
    m = float(min(nx, ny, nz))
    objs = []
    for _ in range(8):  # solid ellipsoids
        c = [rng.uniform(0.2, 0.8) * d for d in (nx, ny, nz)]
        r = [rng.uniform(0.06, 0.2) * m for _ in range(3)]
        objs.append(("solid", c, r, rng.uniform(600, 1500), 0.0))
    for _ in range(3):  # hollow spherical shells
        c = [rng.uniform(0.3, 0.7) * d for d in (nx, ny, nz)]
        R = rng.uniform(0.15, 0.3) * m
        objs.append(("shell", c, [R, R, R], rng.uniform(600, 1500), rng.uniform(4, 8)))
    ramp = 3.0
    for kind, c, r, dens, wall in objs:
        lo = [max(0, int(math.floor(c[k] - r[k] - ramp - 1))) for k in range(3)]
        hi = [min(d, int(math.ceil(c[k] + r[k] + ramp + 2))) for k, d in enumerate((nx, ny, nz))]
        if any(hi[k] <= lo[k] for k in range(3)):
            continue
        X = (np.arange(lo[0], hi[0], dtype=np.float32) - np.float32(c[0]))[None, None, :]
        Y = (np.arange(lo[1], hi[1], dtype=np.float32) - np.float32(c[1]))[None, :, None]
        Z = (np.arange(lo[2], hi[2], dtype=np.float32) - np.float32(c[2]))[:, None, None]
        if kind == "solid":
            # approximate distance to the ellipsoid surface in voxels: (1 - q) * r_min, q = normalised radius
            q = np.sqrt((X / np.float32(r[0])) ** 2 + (Y / np.float32(r[1])) ** 2 + (Z / np.float32(r[2])) ** 2)
            d = (1.0 - q) * np.float32(min(r))
        else:
            rad = np.sqrt(X * X + Y * Y + Z * Z)
            d = np.float32(wall / 2.0) - np.abs(rad - np.float32(r[0] - wall / 2.0))
        w = np.clip(d / np.float32(ramp) + 1.0, 0.0, 1.0)  # 0 outside .. 1 at >= the surface, 3-voxel ramp
        val = (w * np.float32(dens)).astype(np.int16)
        sub = vol[lo[2]:hi[2], lo[1]:hi[1], lo[0]:hi[0]]
        np.maximum(sub, val, out=sub)
    if scale_to_u8:
        vol = (vol.astype(np.int32) * 255 // 1500).clip(0, 255).astype(np.int16)
    return vol


def synth_env(w=2048, h=1024):
    """RGBA8 [h,w,4]: row 0 is v=0 (straight down in the reference's mapping v = asin(-dy)/pi + 0.5)."""
    v = (np.arange(h, dtype=np.float32) + 0.5) / h
    u = (np.arange(w, dtype=np.float32) + 0.5) / w
    V, U = np.meshgrid(v, u, indexing="ij")
    elev = (0.5 - V) * np.float32(math.pi)  # dy = -sin((v-0.5)*pi) -> elevation of the direction
    sky = np.clip(0.35 + 0.65 * np.sin(np.clip(elev, 0, None)), 0, 1)
    r = np.where(elev >= 0, 120 + 100 * (1 - sky), 70.0)
    g = np.where(elev >= 0, 150 + 80 * (1 - sky), 60.0)
    b = np.where(elev >= 0, 200 + 55 * sky, 50.0)
    # sun disc
    su, sv = 0.3, 0.3
    dist = np.sqrt(((U - su) * 2.0) ** 2 + (V - sv) ** 2)
    sun = np.clip(1.0 - dist / 0.04, 0, 1)
    r = r * (1 - sun) + 255 * sun
    g = g * (1 - sun) + 250 * sun
    b = b * (1 - sun) + 220 * sun
    out = np.stack([r, g, b, np.full_like(r, 255.0)], axis=-1)
    return np.clip(out, 0, 255).astype(np.uint8)


def default_tf():
    """ui.cpp:195: tf_rect_selection(0, 500, 1200, 0, 4000), colour 1,1,1,1 -> {255,255,255,255}.  The gradient clause is
    emitted only when the rectangle is tighter than the stats range (tf_part.cpp:65); with the UI's gradient clip
    [0,4000] it is not."""
    return [{"min_v": 500.0, "max_v": 1200.0, "min_g": 0.0, "max_g": 4000.0, "flags": 0, "rgba": (255, 255, 255, 255)}]


def threshold_tf(k=800):
    """tests/sdf/sdf_test.cpp:22"""
    return [{"min_v": float(k), "flags": 2}]
