"""CPU suite for the ingest row: cl_volume_renderer_b200/host/vr_io.hpp (nrrd_loader, hdre_loader)
against numpy restatements, the reference's own fixture (tests/golden/sdf_ref.npz = tests/sdf/testdata.nrrd decoded) and —
where oracle/_ref/ref_loaders exists — the reference's own loaders (app/nrrd_loader.cpp, app/hdre_loader.cpp + stb_image
v2.25) compiled from their sources and run on the same files."""
import gzip
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROBE = os.path.join(ROOT, "tests", "cpp", "io_probe")
REF = os.path.join(ROOT, "oracle", "_ref", "ref_loaders")


def _run(exe, *args):
    out = subprocess.run([exe, *map(str, args)], capture_output=True, text=True, timeout=120)
    return out.returncode, out.stdout.strip(), out.stderr


def write_nrrd(path, vol, encoding="raw", directions="(1,0,0) (0,1,0) (0,0,1)", extra=""):
    """the header keys the reference parses (app/nrrd_loader.cpp:50-110; cf. tests/sdf/testdata.nrrd:1-12)"""
    nz, ny, nx = vol.shape
    hdr = ("NRRD0004\n# synthetic\ntype: short\ndimension: 3\nspace: left-posterior-superior\n"
           f"sizes: {nx} {ny} {nz}\nspace directions: {directions}\nkinds: domain domain domain\nendian: little\n"
           f"encoding: {encoding}\n{extra}space origin: (0,0,0)\n\n").encode()
    payload = vol.astype("<i2").tobytes()
    if encoding == "gzip":
        payload = gzip.compress(payload, 6)
    with open(path, "wb") as f:
        f.write(hdr + payload)


def _nrrd_line(vol, sx=1.0, sy=1.0, sz=1.0):
    flat = vol.reshape(-1).astype(np.int64)
    w = np.arange(flat.size, dtype=np.int64) % 1009
    nz, ny, nx = vol.shape
    return f"{nx} {ny} {nz} {sx:g} {sy:g} {sz:g} {int(flat.sum())} {int((flat * w).sum())} {int(flat[0])}"


def _rgbe(rgb):
    """float RGB -> Radiance RGBE bytes"""
    m = rgb.max(axis=-1)
    e = np.zeros_like(m, dtype=np.int32)
    mant, ex = np.frexp(m)
    scale = np.where(m > 1e-32, mant * 256.0 / np.maximum(m, 1e-38), 0.0)
    out = np.zeros(rgb.shape[:-1] + (4,), dtype=np.uint8)
    out[..., :3] = (rgb * scale[..., None]).astype(np.uint8)
    out[..., 3] = np.where(m > 1e-32, ex + 128, 0).astype(np.uint8)
    del e
    return out


def _rle_channel(row):
    """new-style Radiance RLE of one channel of one scanline"""
    out = bytearray()
    i, n = 0, len(row)
    while i < n:
        run = 1
        while i + run < n and run < 127 and row[i + run] == row[i]:
            run += 1
        if run >= 4:
            out += bytes([128 + run, row[i]])
            i += run
            continue
        j = i
        while j < n and j - i < 128:
            r = 1
            while j + r < n and r < 4 and row[j + r] == row[j]:
                r += 1
            if r >= 4:
                break
            j += 1
        out += bytes([j - i]) + bytes(row[i:j])
        i = j
    return bytes(out)


def write_hdr(path, rgbe, rle):
    h, w, _ = rgbe.shape
    with open(path, "wb") as f:
        f.write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n" + f"-Y {h} +X {w}\n".encode())
        for y in range(h):
            if rle:
                f.write(bytes([2, 2, w >> 8, w & 255]))
                for c in range(4):
                    f.write(_rle_channel(rgbe[y, :, c].tolist()))
            else:
                f.write(rgbe[y].tobytes())


def hdr_to_ldr(rgbe):
    """stbi__hdr_convert + stbi__hdr_to_ldr with gamma 2.2, scale 1 (stb_image.h:1783-1808, hdre_loader.cpp:11-12)"""
    e = rgbe[..., 3].astype(np.int32)
    f1 = np.ldexp(np.float32(1.0), e - 136).astype(np.float32)
    v = np.where(e[..., None] != 0, rgbe[..., :3].astype(np.float32) * f1[..., None], np.float32(0)).astype(np.float32)
    g = np.float64(np.float32(1.0) / np.float32(2.2))
    z = np.power(v.astype(np.float64), g).astype(np.float32) * np.float32(255) + np.float32(0.5)
    z = np.clip(z, 0, 255)
    out = np.empty(rgbe.shape[:2] + (4,), dtype=np.uint8)
    out[..., :3] = z.astype(np.int32).astype(np.uint8)
    out[..., 3] = 255
    return out


@pytest.fixture(scope="module")
def probe():
    assert os.path.exists(PROBE), "build first: make host"
    return PROBE


def _vol(shape, seed):
    rng = np.random.default_rng(seed)
    return rng.integers(-2000, 4000, size=shape, dtype=np.int16)


@pytest.mark.parametrize("encoding", ["raw", "gzip"])
def test_nrrd_loader_matches_numpy(tmp_path, probe, encoding):
    vol = _vol((13, 21, 34), 1)
    p = tmp_path / f"v_{encoding}.nrrd"
    write_nrrd(p, vol, encoding, directions="(0.5,0,0) (0,0.75,0) (0,0,2)")
    rc, out, err = _run(probe, "nrrd", p)
    assert rc == 0, err
    assert out == _nrrd_line(vol, 1.0, 1.5, 4.0)
    if os.path.exists(REF):
        assert _run(REF, "nrrd", p)[1] == out


def test_nrrd_loader_reference_fixture(tmp_path, probe):
    """the reference's own test volume (tests/sdf/testdata.nrrd: 38x35x38, sum 26559114, first voxel -837), re-encoded"""
    vol = np.load(os.path.join(ROOT, "tests", "golden", "sdf_ref.npz"))["volume"]
    p = tmp_path / "testdata.nrrd"
    write_nrrd(p, vol, "gzip")
    rc, out, err = _run(probe, "nrrd", p)
    assert rc == 0, err
    assert out == _nrrd_line(vol)
    assert out.split()[6] == "26559114" and out.split()[8] == "-837"
    orig = "/root/reference/tests/sdf/testdata.nrrd"  # only in the build container
    if os.path.exists(orig):
        assert _run(probe, "nrrd", orig)[1] == out
        if os.path.exists(REF):
            assert _run(REF, "nrrd", orig)[1] == out


def test_nrrd_loader_short_payload_and_errors(tmp_path, probe):
    vol = _vol((4, 5, 6), 2)
    p = tmp_path / "short.nrrd"
    write_nrrd(p, vol, "raw")
    data = open(p, "rb").read()
    open(p, "wb").write(data[:-20])  # ten voxels missing: they stay 0 (the reference's read() leaves the vector's zeros)
    want = vol.copy().reshape(-1)
    want[-10:] = 0
    rc, out, _ = _run(probe, "nrrd", p)
    assert rc == 0 and out == _nrrd_line(want.reshape(vol.shape))
    for bad, msg in [("type: short", "type: float"), ("endian: little", "endian: big"), ("dimension: 3", "dimension: 4"),
                     ("encoding: raw", "encoding: bzip2"), ("sizes: 6 5 4", "sizes: 6 5")]:
        q = tmp_path / "bad.nrrd"
        open(q, "wb").write(data.replace(bad.encode(), msg.encode()))
        rc, _, err = _run(probe, "nrrd", q)
        assert rc == 1 and "Error" in err, (msg, rc, err)   # fail-hard like nrrd_loader.cpp:56-93
    rc, _, err = _run(probe, "nrrd", tmp_path / "missing.nrrd")
    assert rc == 1 and "Error" in err


@pytest.mark.parametrize("rle", [False, True])
def test_hdr_loader_matches_stb_conversion(tmp_path, probe, rle):
    h, w = 16, 40
    rng = np.random.default_rng(7)
    rgb = rng.random((h, w, 3)) ** 3 * 4.0
    rgb[3:6, 5:30] = rgb[3, 5]       # runs, so that the RLE path sees both run and literal packets
    rgb[8, :] = 0.0                   # exponent 0 pixels
    rgbe = _rgbe(rgb)
    p = tmp_path / ("rle.hdr" if rle else "flat.hdr")
    write_hdr(p, rgbe, rle)
    rc, out, err = _run(probe, "env", p, tmp_path / "o.rgba")
    assert rc == 0, err
    assert out == f"{w} {h} 4"
    got = np.fromfile(tmp_path / "o.rgba", dtype=np.uint8).reshape(h, w, 4)
    assert np.array_equal(got, hdr_to_ldr(rgbe))
    if os.path.exists(REF):
        rc, rout, _ = _run(REF, "env", p, tmp_path / "r.rgba")
        assert rc == 0 and rout == out
        assert np.array_equal(np.fromfile(tmp_path / "r.rgba", dtype=np.uint8).reshape(h, w, 4), got)


def test_ppm_and_raw_env_maps(tmp_path, probe):
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, size=(9, 14, 3), dtype=np.uint8)
    p = tmp_path / "e.ppm"
    open(p, "wb").write(b"P6\n# c\n14 9\n255\n" + img.tobytes())
    rc, out, err = _run(probe, "env", p, tmp_path / "o.rgba")
    assert rc == 0 and out == "14 9 4", err
    got = np.fromfile(tmp_path / "o.rgba", dtype=np.uint8).reshape(9, 14, 4)
    assert np.array_equal(got[..., :3], img) and (got[..., 3] == 255).all()
    if os.path.exists(REF):  # stb reads P6 as well
        _run(REF, "env", p, tmp_path / "r.rgba")
        assert np.array_equal(np.fromfile(tmp_path / "r.rgba", dtype=np.uint8).reshape(9, 14, 4), got)
    rgba = rng.integers(0, 256, size=(6, 10, 4), dtype=np.uint8)
    q = tmp_path / "env.10x6.rgba"
    rgba.tofile(q)
    rc, out, _ = _run(probe, "env", q, tmp_path / "o2.rgba")
    assert rc == 0 and out == "10 6 4"
    assert np.array_equal(np.fromfile(tmp_path / "o2.rgba", dtype=np.uint8).reshape(6, 10, 4), rgba)
    rc, _, err = _run(probe, "env", tmp_path / "nothing.hdr")
    assert rc == 1 and "Error" in err
