"""GPU suite: the C++ host shim (cl_volume_renderer_b200/host/vr_host.hpp) driven like the reference's app
(ui::run start-up, render loop, render_tf, and the reference's SDF test) against the CPU oracle."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as o
from cl_volume_renderer_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "tests", "cpp", "frame_emitter_driver")


def test_frame_emitter_flow_matches_oracle(tmp_path):
    assert os.path.exists(DRIVER), "build the driver: make host"
    n, W, H, frames = 64, 160, 120, 4
    v = synth.synth_ct(n)
    env = synth.synth_env(128, 64)
    v.tofile(tmp_path / "vol.raw")
    env.tofile(tmp_path / "env.raw")
    out = subprocess.run([DRIVER, str(tmp_path / "vol.raw"), str(n), str(n), str(n), str(tmp_path / "env.raw"), "128", "64",
                          str(W), str(H), str(frames), str(tmp_path)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr + out.stdout
    assert "EVERYTHING FINE" in out.stdout
    # oracle: same flow — default TF, seeds = glibc rand() stream, camera (-200,200,-200)*n/256, look (0.9, 6.183)
    tf = synth.default_tf()
    ref = o.Renderer(v, env, tf, W, H)
    pos, d = synth.default_camera(n)
    for s in synth.glibc_rand(frames):
        want = ref.render_frame(pos, d, s)
    got = np.fromfile(tmp_path / "frame.bin", dtype=np.uint8).reshape(H, W, 4)
    assert np.array_equal(got[..., 3], want[..., 3])
    mse = np.mean((got[..., :3].astype(np.float64) - want[..., :3].astype(np.float64)) ** 2)
    assert mse == 0 or 10 * np.log10(255.0 ** 2 / mse) >= 45.0
    # render_tf with the UI's clips (value [-2000,3000], gradient [0,4000])
    st = o.fetch_stats(v)
    rng = [float(max(-2000, st[0])), float(min(3000, st[1])), float(max(0, st[2])), float(min(4000, st[3]))]
    # the driver calls render_tf(100, 80); the implementation names its parameters (height, width) (renderer.cpp:45 vs
    # renderer.hpp:26), so the image is 80 wide and 100 tall
    want_tf, _, _ = o.tf_color_frame(o.histogram(v, 80, 100, rng), 80, 100)
    got_tf = np.fromfile(tmp_path / "tf.bin", dtype=np.uint8).reshape(100, 80, 4)
    assert np.array_equal(got_tf, want_tf)
    # the reference's SDF test flow (`value > 800`)
    got_sdf = np.fromfile(tmp_path / "sdf.bin", dtype=np.int8).reshape(n, n, n)
    assert np.array_equal(got_sdf, o.sdf_build(v, o.tf_threshold(800))[0])
    assert f"stats {rng[0]:g} {rng[1]:g} {rng[2]:g} {rng[3]:g}" in out.stdout


def test_headless_cli_matches_oracle(tmp_path):
    """vr_headless = app/main.cpp + ui::run without the window: NRRD (gzip) + Radiance .hdr from disk, the
    UI's start-up sequence, `--spp` frames with the std::rand() seeds, frame written upright as PPM and raw RGBA."""
    from test_io_cpu import hdr_to_ldr, write_hdr, write_nrrd, _rgbe
    cli = os.path.join(ROOT, "cl_volume_renderer_b200", "host", "vr_headless")
    assert os.path.exists(cli), "build first: make host"
    n, W, H, spp = 64, 160, 120, 3
    v = synth.synth_ct(n)
    write_nrrd(tmp_path / "vol.nrrd", v, "gzip")
    rgbe = _rgbe(synth.synth_env(128, 64)[..., :3].astype(np.float64) / 255.0 * 1.5)
    write_hdr(tmp_path / "env.hdr", rgbe, rle=True)
    env = hdr_to_ldr(rgbe)
    out = subprocess.run([cli, str(tmp_path / "vol.nrrd"), str(tmp_path / "env.hdr"), "--w", str(W), "--h", str(H), "--spp", str(spp),
                          "--out", str(tmp_path / "f.ppm"), "--raw", str(tmp_path / "f.rgba"), "--tf-image", str(tmp_path / "tf.ppm")],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr + out.stdout
    ref = o.Renderer(v, env, synth.default_tf(), W, H)
    pos, d = synth.default_camera(n)
    for s in synth.glibc_rand(spp):
        want = ref.render_frame(pos, d, s)
    got = np.fromfile(tmp_path / "f.rgba", dtype=np.uint8).reshape(H, W, 4)
    assert np.array_equal(got[..., 3], want[..., 3])
    mse = np.mean((got[..., :3].astype(np.float64) - want[..., :3].astype(np.float64)) ** 2)
    assert mse == 0 or 10 * np.log10(255.0 ** 2 / mse) >= 45.0
    ppm = open(tmp_path / "f.ppm", "rb").read()
    head = f"P6\n{W} {H}\n255\n".encode()
    assert ppm.startswith(head)
    img = np.frombuffer(ppm[len(head):], dtype=np.uint8).reshape(H, W, 3)
    assert np.array_equal(img, got[::-1, :, :3])  # the frame's row 0 is the bottom of the view
    tf = open(tmp_path / "tf.ppm", "rb").read()
    assert tf.startswith(b"P6\n500 500\n255\n") and len(tf) == 15 + 500 * 500 * 3
    # --frame-filter: 2d_image_filter.cl over the final frame (reference mode = the kernel as written, bit-exact)
    out2 = subprocess.run([cli, str(tmp_path / "vol.nrrd"), str(tmp_path / "env.hdr"), "--w", str(W), "--h", str(H), "--spp", str(spp),
                           "--out", str(tmp_path / "g.ppm"), "--raw", str(tmp_path / "g.rgba"), "--frame-filter", "2", "1.5", "reference"],
                          capture_output=True, text=True, timeout=120)
    assert out2.returncode == 0, out2.stderr + out2.stdout
    got2 = np.fromfile(tmp_path / "g.rgba", dtype=np.uint8).reshape(H, W, 4)
    assert np.array_equal(got2, o.image_filter2d(got, 2, 1.5))
