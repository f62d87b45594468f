"""GPU parity suite (-m gpu): every kernel of the hot path through the C-ABI against the CPU oracle."""
import os

import numpy as np
import pytest

import oracle_lib as o
from cl_volume_renderer_b200 import api, synth

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _ragged(seed=3):
    # ragged, non-multiple-of-tile dims; values span negative and positive like the reference fixture
    rng = np.random.default_rng(seed)
    v = synth.synth_ct(0, dims=(45, 37, 29)).astype(np.int32)
    v += rng.integers(-900, 200, size=v.shape)
    return v.clip(-2000, 4095).astype(np.int16)


# ---- SDF: bit exact ------------------------------------------------------------------------------------------
def test_sdf_reference_golden_vector(vr_ctx):
    g = np.load(os.path.join(GOLDEN, "sdf_ref.npz"))
    vol = api.Volume(vr_ctx, g["volume"])
    sdf = api.Sdf(vr_ctx, vol, synth.threshold_tf(int(g["threshold"])))
    assert np.array_equal(sdf.download(), g["sdf"])
    sdf.close(); vol.close()


@pytest.mark.parametrize("dims,tf", [((45, 37, 29), "default"), ((64, 64, 64), "default"), ((96, 80, 72), "thr"),
                                     ((8, 8, 8), "default"), ((2, 2, 2), "thr"), ((1, 5, 3), "thr"),
                                     ((130, 20, 20), "grad"),
                                     # rows of a multiple of 4 words run the 128-voxels-per-thread level kernel (k_sdf_wave8): one quad,
                                     # lanes beyond the row, 8 lanes along x, rows wider than a warp tile (edge loads), ragged y / z
                                     ((128, 40, 36), "default"), ((256, 23, 21), "thr"), ((640, 19, 18), "default"),
                                     ((1152, 13, 11), "thr"), ((384, 70, 9), "grad")])
def test_sdf_matches_oracle(vr_ctx, dims, tf):
    v = synth.synth_ct(0, dims=dims)
    tfs = {"default": synth.default_tf(), "thr": synth.threshold_tf(300),
           "grad": [{"min_v": 100.0, "max_v": 1400.0, "min_g": 50.0, "max_g": 900.0, "flags": 1, "rgba": (255, 0, 0, 128)}]}[tf]
    want, _ = o.sdf_build(v, tfs)
    vol = api.Volume(vr_ctx, v)
    sdf = api.Sdf(vr_ctx, vol, tfs)
    got = sdf.download()
    assert np.array_equal(got, want)
    sdf.close(); vol.close()


def test_sdf_empty_and_full_volumes(vr_ctx):
    for fill in (0, 800):  # no event anywhere / event everywhere: homogeneous -> +-max_it
        v = np.full((20, 24, 28), fill, dtype=np.int16)
        vol = api.Volume(vr_ctx, v)
        sdf = api.Sdf(vr_ctx, vol, synth.default_tf())
        got = sdf.download()
        assert np.array_equal(got, o.sdf_build(v, synth.default_tf())[0])
        assert (np.abs(got) == 14).all()
        sdf.close(); vol.close()


# ---- stats / histogram / clip / filter --------------------------------------------------------------------------
def test_fetch_stats_bit_exact(vr_ctx):
    for v in (_ragged(), synth.synth_ct(64), np.load(os.path.join(GOLDEN, "sdf_ref.npz"))["volume"]):
        vol = api.Volume(vr_ctx, v)
        assert vol.stats() == o.fetch_stats(v)
        vol.close()


def test_fetch_stats_integer_path_and_its_fp32_fallback(vr_ctx):
    """nx % 8 == 0 takes the speculative integer kernel: exact while neighbouring voxels are less than 4096 apart and the squared
    gradient stays below 2^24; volumes beyond that (full-range noise, a single 5000 step, values straddling zero by 4096) must come
    out of the fp32 rerun — all bit-exact against the oracle, blocking and asynchronous upload alike."""
    rng = np.random.default_rng(7)
    vols = [synth.synth_ct(64),
            rng.integers(-32768, 32768, size=(24, 16, 40)).astype(np.int16),            # every difference huge: fallback
            rng.integers(0, 4095, size=(16, 24, 32)).astype(np.int16),                  # range 4094 but sums up to 3*4094^2 > 2^24: fallback
            rng.integers(-100, 100, size=(16, 16, 16)).astype(np.int16),                 # small: integer path
            np.full((8, 8, 8), 4095, np.int16), np.full((8, 8, 16), -4096, np.int16)]   # border zeros 4095 / 4096 away
    step = np.zeros((16, 16, 24), np.int16); step[:, :, 12:] = 5000
    vols.append(step)
    for v in vols:
        want = o.fetch_stats(v)
        vol = api.Volume(vr_ctx, v)
        assert vol.stats() == want
        vol.close()
        vol = api.Volume(vr_ctx, v, async_upload=True)
        assert vol.stats() == want
        vol.close()


@pytest.mark.parametrize("which", ["ragged", "x8"])  # nx % 8 == 0 takes the vectorised kernels
def test_histogram_bit_exact(vr_ctx, which):
    v = _ragged() if which == "ragged" else synth.synth_ct(0, dims=(72, 37, 29))
    st = o.fetch_stats(v)
    vol = api.Volume(vr_ctx, v)
    for (w, h, rng) in [(50, 40, st), (500, 500, st), (64, 64, [-2000, 3000, 0, 4000]), (17, 9, [0, 100, 5, 50])]:
        got = vol.histogram(w, h, [float(x) for x in rng])
        want = o.histogram(v, w, h, [float(x) for x in rng])
        assert np.array_equal(got, want)
    vol.close()


def test_histogram_table_path_bit_exact(vr_ctx):
    """nx % 8 == 0 and every voxel difference below 4096: k_histogram_lut (columns / rows from tables of the reference's expressions,
    steep gradients through the per-block table or the expression itself).  Noise of the full admissible span makes squared
    gradients up to 3 * 3999^2 (beyond fp32's exact integers: those voxels are above any admissible max_g in both arithmetics)."""
    rng = np.random.default_rng(5)
    v = rng.integers(-2000, 2000, size=(24, 40, 64)).astype(np.int16)
    v[8:16, 10:30, 16:48] = 1500                                        # a homogeneous block: runs, window hits
    st = o.fetch_stats(v)
    vol = api.Volume(vr_ctx, v)
    for (w, h, r) in [(256, 256, st), (64, 64, [-2000, 3000, 0, 4000]), (100, 300, [-500, 1200, 20, 2500]), (31, 17, [0, 100, 5, 50])]:
        got = vol.histogram(w, h, [float(x) for x in r])
        want = o.histogram(v, w, h, [float(x) for x in r])
        assert np.array_equal(got, want), (w, h, r)
    vol.close()
    v2 = synth.synth_ct(0, dims=(128, 48, 40))
    st2 = o.fetch_stats(v2)
    vol = api.Volume(vr_ctx, v2)
    for (w, h, r) in [(500, 500, st2), (256, 128, [-2000, 3000, 0, 4000])]:
        assert np.array_equal(vol.histogram(w, h, [float(x) for x in r]), o.histogram(v2, w, h, [float(x) for x in r]))
    vol.close()


def test_clip_bit_exact(vr_ctx):
    v = _ragged()
    vol = api.Volume(vr_ctx, v)
    vol.clip((3, 5, 2), (40, 30, 27))
    assert vol.dims() == [37, 25, 25]
    assert np.array_equal(vol.download(), o.clip(v, (3, 5, 2), (37, 25, 25)))
    assert vol.stats() == o.fetch_stats(v)  # stats stay those of the original volume (reference_volume.hpp:33-34)
    vol.close()


def test_bilateral_filter(vr_ctx):
    # fp32 with 125 exp() per voxel: expf differs by an ulp between glibc and CUDA, so the truncated short may
    # differ by 1 where out/wp lands next to an integer (measured: 0.25 % of voxels).  Tolerance: |diff| <= 1, >= 99 % exact.
    v = synth.synth_ct(0, dims=(40, 33, 21))
    vol = api.Volume(vr_ctx, v)
    vol.filter()
    got = vol.download().astype(np.int32)
    want = o.bilateral(v).astype(np.int32)
    assert np.abs(got - want).max() <= 1
    assert (got == want).mean() >= 0.99
    vol.close()


def test_render_tf_image(vr_ctx):
    v = synth.synth_ct(48)
    vol = api.Volume(vr_ctx, v)
    vol.set_value_clip(-2000, 3000); vol.set_gradient_clip(0, 4000)  # ui.cpp:187-188
    env = api.EnvMap(vr_ctx, synth.synth_env(64, 32))
    r = api.Renderer(vr_ctx, 64, 48)
    r.image_set(vol, env)
    got = r.render_tf(100, 80)
    rng = vol.clipped_stats()
    st = o.fetch_stats(v)
    assert rng == [float(max(-2000, st[0])), float(min(3000, st[1])), float(max(0, st[2])), float(min(4000, st[3]))]
    want, _, n = o.tf_color_frame(o.histogram(v, 100, 80, rng), 100, 80)
    assert n > 3 and np.array_equal(got, want)
    r.close(); env.close(); vol.close()


# ---- render ------------------------------------------------------------------------------------------------------
# 2: k_primary once per pixel + k_trace_pt per (pixel, frame) (default), 1: hybrid k_trace + k_trace_pt per frame,
# 0: k_trace alone; tests using `both` run each
TRACE_MODE = [2]


@pytest.fixture(params=[2, 1, 0], ids=["primary_reuse", "hybrid", "k_trace"])
def both(request):
    TRACE_MODE[0] = request.param
    yield request.param
    TRACE_MODE[0] = 2


def _scene(vr_ctx, n, W, H, tf=None, token_cap=256):
    v = synth.synth_ct(n)
    envimg = synth.synth_env(256, 128)
    tf = tf or synth.default_tf()
    vol = api.Volume(vr_ctx, v)
    env = api.EnvMap(vr_ctx, envimg)
    r = api.Renderer(vr_ctx, W, H)
    r.image_set(vol, env)
    r.set_tf(tf)
    r.set_token_cap(token_cap)
    r.set_trace_mode(TRACE_MODE[0])
    r.flush_changes()
    ref = o.Renderer(v, envimg, tf, W, H, token_cap=token_cap)
    return r, ref, (vol, env)


def _psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


def test_render_sdf_built_by_flush_matches(vr_ctx):
    r, ref, keep = _scene(vr_ctx, 48, 64, 48)
    assert np.array_equal(r.sdf_download(), ref.sdf)
    r.close(); [k.close() for k in keep]


@pytest.mark.parametrize("n,W,H,frames", [(64, 160, 120, 6), (96, 200, 136, 3)])
def test_render_cache_and_image_parity(vr_ctx, both, n, W, H, frames):
    # Ray positions are bit-identical by construction (no FMA, IEEE div/sqrt); the only sources of difference are
    # atan2f/asinf ulps at environment-texel boundaries and powf in the tone map.
    # Stated tolerance: voxel cache >= 99.9% of entries identical and max |diff| <= 64 per 16-bit lane;
    # frame PSNR >= 45 dB, max abs pixel diff <= 8, alpha channel identical.
    r, ref, keep = _scene(vr_ctx, n, W, H)
    pos, d = synth.default_camera(n)
    r.enable_counters(True)
    for k in range(frames):
        seed = synth.glibc_rand(frames)[k]
        got = r.render_frame(pos, d, seed)
        want = ref.render_frame(pos, d, seed)
    gc = r.cache_download().astype(np.int32)
    wc = ref.cache.astype(np.int32)
    assert np.array_equal(gc.reshape(-1, 4)[:, 3], wc.reshape(-1, 4)[:, 3])  # token counts: exact
    assert (gc == wc).mean() >= 0.999
    assert np.abs(gc - wc).max() <= 64
    assert np.array_equal(got[..., 3], want[..., 3])
    assert (got[..., 3] == 1).mean() > 0.05
    assert _psnr(got[..., :3], want[..., :3]) >= 45.0
    assert np.abs(got[..., :3].astype(int) - want[..., :3].astype(int)).max() <= 8
    c = r.counters()
    oc = dict(zip(["steps", "normals", "env", "primary_hits", "admitted", "samples"], [int(x) for x in ref.counters]))
    assert c == oc  # per-sample work counters (march steps, hits, env fetches) identical
    r.close(); [k.close() for k in keep]


def test_render_token_cap_saturates(vr_ctx, both):
    r, ref, keep = _scene(vr_ctx, 32, 96, 72, token_cap=3)
    pos, d = synth.default_camera(32)
    for k in range(6):
        r.render_frame(pos, d, 1000 + k, readback=False)
    tokens = r.cache_download().reshape(-1, 4)[:, 3]
    assert tokens.max() == 3
    r.close(); [k.close() for k in keep]


def test_render_camera_inside_and_axis_aligned(vr_ctx, both):
    # camera inside the volume, looking along +x exactly: zero direction components exercise the inf/NaN paths of
    # cut() and positions that land exactly on integer coordinates
    r, ref, keep = _scene(vr_ctx, 48, 64, 64)
    for pos, d in [((24.0, 24.0, 24.0), (1.0, 0.0, 0.0)), ((-30.0, 24.0, 24.0), (1.0, 0.0, 0.0)),
                   ((24.0, 100.0, 24.0), (0.0, -1.0, 0.0)), ((200.0, 200.0, 200.0), (0.577, 0.577, 0.577))]:
        r.reset_cache(); ref.reset()
        got = r.render_frame(pos, d, 77)
        want = ref.render_frame(pos, d, 77)
        assert np.array_equal(got[..., 3], want[..., 3])
        assert _psnr(got[..., :3], want[..., :3]) >= 45.0
    r.close(); [k.close() for k in keep]


def test_render_threshold_tf_and_multi_clause_tf(vr_ctx, both):
    tf2 = [{"min_v": 900.0, "max_v": 1200.0, "min_g": 0.0, "max_g": 0.0, "flags": 0, "rgba": (255, 64, 32, 128)},
           {"min_v": 500.0, "max_v": 1500.0, "min_g": 100.0, "max_g": 2000.0, "flags": 1, "rgba": (40, 200, 255, 255)}]
    for tf in (synth.threshold_tf(700), tf2):
        r, ref, keep = _scene(vr_ctx, 48, 96, 64, tf=tf)
        pos, d = synth.default_camera(48)
        for k in range(3):
            got = r.render_frame(pos, d, 5 + k)
            want = ref.render_frame(pos, d, 5 + k)
        assert np.array_equal(r.sdf_download(), ref.sdf)
        assert np.array_equal(got[..., 3], want[..., 3])
        assert _psnr(got[..., :3], want[..., :3]) >= 45.0
        gc, wc = r.cache_download().astype(np.int32), ref.cache.astype(np.int32)
        assert (gc == wc).mean() >= 0.999
        r.close(); [k.close() for k in keep]


def test_render_rows_window_and_resolve(vr_ctx, both):
    # image-tile split hook: tracing rows [y0,y1) only touches those pixels
    r, ref, keep = _scene(vr_ctx, 48, 64, 48)
    pos, d = synth.default_camera(48)
    r.set_rows(16, 32)
    got = r.render_frame(pos, d, 9)
    want = ref.render_frame(pos, d, 9, window=(0, 16, 64, 32))
    assert np.array_equal(got[16:32, :, 3], want[16:32, :, 3])
    assert _psnr(got[16:32, :, :3], want[16:32, :, :3]) >= 45.0
    assert (got[:16] == 0).all() and (got[32:] == 0).all()
    r.close(); [k.close() for k in keep]


def test_tf_code_entry_point_equals_table(vr_ctx):
    r, ref, keep = _scene(vr_ctx, 32, 64, 48)
    pos, d = synth.default_camera(32)
    a = r.render_frame(pos, d, 3)
    r.next_event_code_set(api.tf_format(synth.default_tf()))
    r.flush_changes()
    b = r.render_frame(pos, d, 3)
    assert np.array_equal(a, b)
    r.close(); [k.close() for k in keep]


def test_render_frames_batch_equals_sequential_oracle(vr_ctx, both):
    # vr_render_frames traces the whole batch in one launch (gridDim.z = frame) and resolves once; below the token cap
    # the samples commute (integer atomics), so cache and final frame must equal the oracle's frame-by-frame result.
    r, ref, keep = _scene(vr_ctx, 64, 128, 96)
    pos, d = synth.default_camera(64)
    seeds = synth.glibc_rand(10)
    got = r.render_frames(pos, d, seeds)
    for s in seeds:
        want = ref.render_frame(pos, d, s)
    gc, wc = r.cache_download().astype(np.int32), ref.cache.astype(np.int32)
    assert gc.reshape(-1, 4)[:, 3].max() < 256
    assert np.array_equal(gc.reshape(-1, 4)[:, 3], wc.reshape(-1, 4)[:, 3])
    assert (gc == wc).mean() >= 0.999
    assert np.array_equal(got[..., 3], want[..., 3])
    assert _psnr(got[..., :3], want[..., :3]) >= 45.0
    r.close(); [k.close() for k in keep]


def test_primary_reuse_across_calls_and_invalidation(vr_ctx):
    """vr_renderer_set_primary_reuse(2): per-frame calls with an unchanged camera reuse the primary records; a camera move, a
    row-range change or a flush re-marches.  Results equal the oracle frame by frame."""
    r, ref, keep = _scene(vr_ctx, 64, 128, 96)
    r.set_primary_reuse(2)
    pos, d = synth.default_camera(64)
    pos2 = pos + np.array([6.0, -3.0, 2.0], dtype=np.float32)
    seeds = synth.glibc_rand(9)
    l0 = vr_ctx.launches
    for k, s in enumerate(seeds):
        cam = pos if k < 4 or k >= 7 else pos2   # 4 frames, move, 3 frames, move back, 2 frames
        got = r.render_frame(cam, d, s)
        want = ref.render_frame(cam, d, s)
        assert np.array_equal(got[..., 3], want[..., 3]), k
        assert _psnr(got[..., :3], want[..., :3]) >= 45.0
    assert vr_ctx.launches - l0 == 9 * 2 + 3   # trace + resolve per frame, k_primary only after the three camera changes
    gc, wc = r.cache_download().astype(np.int32), ref.cache.astype(np.int32)
    assert np.array_equal(gc.reshape(-1, 4)[:, 3], wc.reshape(-1, 4)[:, 3])
    assert (gc == wc).mean() >= 0.999
    r.flush_changes(); ref.reset()
    l0 = vr_ctx.launches
    got = r.render_frames(pos, d, seeds[:3]); [ref.render_frame(pos, d, s) for s in seeds[:2]]
    want = ref.render_frame(pos, d, seeds[2])
    assert vr_ctx.launches - l0 == 3   # k_primary (flush invalidated the records), one k_trace_pt, one k_resolve
    assert np.array_equal(got[..., 3], want[..., 3]) and _psnr(got[..., :3], want[..., :3]) >= 45.0
    r.close(); [k.close() for k in keep]


@pytest.mark.parametrize("mode", [2, 1, 0])
def test_render_far_face_positions(vr_ctx, mode):
    # rays that land exactly on the far faces (coordinate == dim): axis-aligned steps of 0.5 from integer origins.
    # The SDF apron must behave like the reference's border read (0 -> step 0.5) and the TF must see value 0.
    n = 16
    v = np.zeros((n, n, n), dtype=np.int16)
    v[:, :, n - 1] = 900  # a slab touching the far x face, so the gradient tap at x == n sees it
    envimg = synth.synth_env(64, 32)
    tf = [{"min_v": -10.0, "max_v": 10.0, "min_g": 500.0, "max_g": 2000.0, "flags": 1, "rgba": (200, 100, 50, 255)},
          {"min_v": 500.0, "max_v": 1200.0, "min_g": 0.0, "max_g": 0.0, "flags": 0, "rgba": (255, 255, 255, 255)}]
    vol = api.Volume(vr_ctx, v); env = api.EnvMap(vr_ctx, envimg)
    r = api.Renderer(vr_ctx, 32, 32); r.image_set(vol, env); r.set_tf(tf); r.set_trace_mode(mode); r.flush_changes()
    ref = o.Renderer(v, envimg, tf, 32, 32)
    assert np.array_equal(r.sdf_download(), ref.sdf)
    for pos, d in [((-4.0, 8.0, 8.0), (1.0, 0.0, 0.0)), ((8.0, 8.0, -4.0), (0.0, 0.0, 1.0)), ((20.0, 8.0, 8.0), (-1.0, 0.0, 0.0))]:
        r.reset_cache(); ref.reset()
        got = r.render_frame(pos, d, 5); want = ref.render_frame(pos, d, 5)
        assert np.array_equal(got[..., 3], want[..., 3])
        assert np.array_equal(r.cache_download().reshape(-1, 4)[:, 3], ref.cache.reshape(-1, 4)[:, 3])
    r.close(); env.close(); vol.close()


@pytest.mark.parametrize("mode", ["front_overflow", "front", "reg", "wave1", "wave2", "wave3", "wave4", "warp"])
def test_sdf_alternative_builds_bit_exact(mode):
    """The schedules of the SDF build that were tried on the way live in the A/B build of the library (tools/ab/libvr_ab.so, `make ab`;
    selected with VR_SDF_MODE; a frontier list that overflows restarts with the dense per-level kernel).  All must reproduce the
    reference's golden vector and the oracle on a synthetic volume (own process: the mode is read once per process)."""
    import subprocess, sys, os
    env = dict(os.environ)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env["VR_LIB"] = os.path.join(root, "tools", "ab", "libvr_ab.so")
    assert os.path.exists(env["VR_LIB"]), "build the A/B library: make ab"
    if mode == "front_overflow":
        env["VR_SDF_MODE"] = "front"
        env["VR_SDF_FRONT_CAP"] = "64"
    else:
        env["VR_SDF_MODE"] = mode
    code = (
        "import sys, numpy as np\n"
        "sys.path.insert(0, 'tests')\n"
        "import oracle_lib as o\n"
        "from cl_volume_renderer_b200 import api, synth\n"
        "ctx = api.Context(0)\n"
        "G = np.load('tests/golden/sdf_ref.npz')\n"
        "vol = api.Volume(ctx, G['volume']); s = api.Sdf(ctx, vol, synth.threshold_tf(int(G['threshold'])))\n"
        "assert np.array_equal(s.download(), G['sdf'])\n"
        "v = synth.synth_ct(72, dims=(70, 45, 96)); vol2 = api.Volume(ctx, v)\n"
        "for tf in (synth.default_tf(), synth.threshold_tf(700)):\n"
        "    s2 = api.Sdf(ctx, vol2, tf)\n"
        "    assert np.array_equal(s2.download(), o.sdf_build(v, tf)[0])\n"
        "print('OK')\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr


@pytest.mark.parametrize("knobs", [{"VR_SDF_FLOW": "1"}, {"VR_SDF_FLOW": "0"}, {"VR_SDF_FLOW": "0", "VR_SDF_PDL": "0", "VR_SDF_VARIANT": "0"},
                                   {"VR_SDF_FLOW": "1", "VR_SDF_VARIANT": "4"}, {"VR_SDF_WAVE": "6"}, {"VR_SDF_WAVE": "5"}],
                         ids=["flow", "per_level", "per_level_2x8_no_pdl", "flow_1x8", "wave6", "wave5"])
def test_sdf_level_kernels_bit_exact(knobs):
    """Every way the product can run the level wave — tiles synchronised point to point in one cooperative launch (k_sdf_flow, what
    volumes above 512^3 take), a launch per level (k_sdf_wave9), the two-volume kernel for odd row widths (k_sdf_wave6) — and round 1's
    k_sdf_wave5, forced through the A/B build's knobs at sizes the oracle finishes in seconds: rows of 1, 2, 5 and 9 quads (lanes beyond
    the row, 8 lanes along x, edge loads), several tiles per axis, ragged y / z.  Own process: the knobs are read once per process."""
    import subprocess, sys, os
    env = dict(os.environ)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env["VR_LIB"] = os.path.join(root, "tools", "ab", "libvr_ab.so")
    assert os.path.exists(env["VR_LIB"]), "build the A/B library: make ab"
    env.update(knobs)
    code = (
        "import sys, numpy as np\n"
        "sys.path.insert(0, 'tests')\n"
        "import oracle_lib as o\n"
        "from cl_volume_renderer_b200 import api, synth\n"
        "ctx = api.Context(0)\n"
        "G = np.load('tests/golden/sdf_ref.npz')\n"
        "vol = api.Volume(ctx, G['volume']); s = api.Sdf(ctx, vol, synth.threshold_tf(int(G['threshold'])))\n"
        "assert np.array_equal(s.download(), G['sdf'])\n"
        "for dims, tf in (((128, 70, 45), synth.default_tf()), ((256, 23, 21), synth.threshold_tf(300)), ((640, 19, 37), synth.default_tf()),\n"
        "                 ((1152, 13, 11), synth.threshold_tf(700))):\n"
        "    v = synth.synth_ct(0, dims=dims); vol2 = api.Volume(ctx, v)\n"
        "    s2 = api.Sdf(ctx, vol2, tf)\n"
        "    assert np.array_equal(s2.download(), o.sdf_build(v, tf)[0]), dims\n"
        "    s2.close(); vol2.close()\n"
        "print('OK')\n")
    out = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr


def test_ab_build_trace_sm_equals_default_schedule():
    """The A/B build's trace mode 3 (k_trace_sm: sample slots in shared memory, packed batches — measured and not adopted, see
    vr_render.cu) computes the same samples: voxel cache and counters identical to the default schedule, under both samplings."""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    env["VR_LIB"] = os.path.join(root, "tools", "ab", "libvr_ab.so")
    assert os.path.exists(env["VR_LIB"]), "build the A/B library: make ab"
    code = (
        "import numpy as np\n"
        "from cl_volume_renderer_b200 import api, synth\n"
        "ctx = api.Context(0)\n"
        "n, W, H = 64, 160, 120\n"
        "vol = api.Volume(ctx, synth.synth_ct(n)); env = api.EnvMap(ctx, synth.synth_env(128, 64))\n"
        "seeds = synth.glibc_rand(4)\n"
        "for sampling in (api.VR_SAMPLING_NEAREST, api.VR_SAMPLING_HW_LINEAR):\n"
        "    res = []\n"
        "    for mode in (2, 3):\n"
        "        r = api.Renderer(ctx, W, H); r.set_sampling(sampling); r.set_trace_mode(mode)\n"
        "        r.image_set(vol, env); r.set_tf(synth.default_tf()); r.flush_changes(); r.enable_counters(True)\n"
        "        out = []\n"
        "        for cam in (synth.default_camera(n), synth.closeup_camera(n)):\n"
        "            r.reset_cache()\n"
        "            f = r.render_frames(cam[0], cam[1], seeds)\n"
        "            out.append((r.cache_download(), f))\n"
        "        c = r.counters(); r.close()\n"
        "        assert c['admitted'] == c['primary_hits'] > 0   # the token cap is never reached: admissions do not depend on the order\n"
        "        res.append((out, c))\n"
        "    assert res[0][1] == res[1][1]\n"
        "    for k in range(2):\n"
        "        assert np.array_equal(res[0][0][k][0], res[1][0][k][0]) and np.array_equal(res[0][0][k][1], res[1][0][k][1])\n"
        "print('OK')\n")
    out = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr


def test_render_odd_shapes_and_empty_tf(vr_ctx, both):
    """Edge cases the reference leaves to luck: a frame that is not a multiple of its 8x8 work-groups (render has no bounds guard,
    clw_function.hpp:235-237 asserts instead), a volume whose dims are not multiples of the brick / word sizes, an empty transfer
    function (every pixel is environment), and a 1-voxel-thin volume."""
    env = synth.synth_env(96, 48)
    cases = [((45, 37, 29), 50, 37, synth.default_tf()), ((33, 70, 20), 61, 43, synth.threshold_tf(400)),
             ((24, 24, 24), 40, 24, []), ((40, 1, 40), 48, 32, synth.threshold_tf(200))]
    for dims, W, H, tf in cases:
        v = synth.synth_ct(0, dims=dims)
        vol = api.Volume(vr_ctx, v); e = api.EnvMap(vr_ctx, env)
        r = api.Renderer(vr_ctx, W, H); r.image_set(vol, e); r.set_tf(tf); r.set_trace_mode(TRACE_MODE[0]); r.flush_changes()
        ref = o.Renderer(v, env, tf, W, H)
        assert np.array_equal(r.sdf_download(), ref.sdf)
        m = float(max(dims))
        pos = np.array([-0.8 * m, 0.9 * m, -0.7 * m], dtype=np.float32)
        d = synth.camera_dir(0.9, 6.183)
        seeds = synth.glibc_rand(3)
        got = r.render_frames(pos, d, seeds)
        for s in seeds:
            want = ref.render_frame(pos, d, s)
        assert np.array_equal(got[..., 3], want[..., 3])
        if not tf:
            assert (got[..., 3] == 200).all()
        assert _psnr(got[..., :3], want[..., :3]) >= 45.0
        gc, wc = r.cache_download().astype(np.int32), ref.cache.astype(np.int32)
        assert np.array_equal(gc.reshape(-1, 4)[:, 3], wc.reshape(-1, 4)[:, 3])
        r.close(); e.close(); vol.close()


def test_async_volume_upload_equals_blocking(vr_ctx):
    """vr_volume_upload_async: copy + fetch_stats on the copy stream; results (stats, SDF, frame) equal the blocking upload, also
    when a second upload is in flight while the first volume renders."""
    import torch
    v1, v2 = synth.synth_ct(64), synth.synth_ct(48, seed=3)
    p1 = torch.empty(v1.shape, dtype=torch.int16, pin_memory=True); p1.numpy()[...] = v1
    p2 = torch.empty(v2.shape, dtype=torch.int16, pin_memory=True); p2.numpy()[...] = v2
    env = api.EnvMap(vr_ctx, synth.synth_env(128, 64))
    a = api.Volume(vr_ctx, p1.numpy(), async_upload=True)
    b = api.Volume(vr_ctx, p2.numpy(), async_upload=True)     # in flight while `a` is used
    r = api.Renderer(vr_ctx, 96, 64); r.image_set(a, env); r.set_tf(synth.default_tf()); r.flush_changes()
    pos, d = synth.default_camera(64)
    fa = r.render_frames(pos, d, synth.glibc_rand(3))
    sa = r.sdf_download()
    ref = api.Volume(vr_ctx, v1)
    r2 = api.Renderer(vr_ctx, 96, 64); r2.image_set(ref, env); r2.set_tf(synth.default_tf()); r2.flush_changes()
    assert a.stats() == ref.stats() == o.fetch_stats(v1)
    assert np.array_equal(sa, r2.sdf_download())
    assert np.array_equal(fa, r2.render_frames(pos, d, synth.glibc_rand(3)))
    b.wait()
    assert b.stats() == o.fetch_stats(v2)
    assert np.array_equal(b.download(), v2)
    [x.close() for x in (r, r2, a, b, ref, env)]


def test_config1_256_uint8_volume_8spp_640x480(vr_ctx):
    """BASELINE config 1 — the reference's CPU-runnable case: 256^3 "uint8" volume (values 0..255 stored as short, the only
    type the loader accepts), SDF build + 8 spp at 640x480.  Full-size parity against the oracle."""
    n, W, H = 256, 640, 480
    v = synth.synth_ct(n, scale_to_u8=True)
    assert v.min() >= 0 and v.max() <= 255
    tf = [{"min_v": 85.0, "max_v": 204.0, "min_g": 0.0, "max_g": 4000.0, "flags": 0, "rgba": (255, 255, 255, 255)}]  # rect(500,1200) * 255/1500
    envimg = synth.synth_env(512, 256)
    vol = api.Volume(vr_ctx, v); env = api.EnvMap(vr_ctx, envimg)
    r = api.Renderer(vr_ctx, W, H); r.image_set(vol, env); r.set_tf(tf); r.flush_changes()
    ref = o.Renderer(v, envimg, tf, W, H)
    assert np.array_equal(r.sdf_download(), ref.sdf)
    assert vol.stats() == o.fetch_stats(v)
    pos, d = synth.default_camera(n)
    seeds = synth.glibc_rand(8)
    r.enable_counters(True)
    got = r.render_frames(pos, d, seeds)
    for s in seeds:
        want = ref.render_frame(pos, d, s)
    c = r.counters()
    assert c == dict(zip(["steps", "normals", "env", "primary_hits", "admitted", "samples"], [int(x) for x in ref.counters]))
    gc, wc = r.cache_download().astype(np.int32), ref.cache.astype(np.int32)
    assert np.array_equal(gc.reshape(-1, 4)[:, 3], wc.reshape(-1, 4)[:, 3])
    assert (gc == wc).mean() >= 0.999 and np.abs(gc - wc).max() <= 64
    assert np.array_equal(got[..., 3], want[..., 3]) and (got[..., 3] == 1).mean() > 0.03
    assert _psnr(got[..., :3], want[..., :3]) >= 45.0
    r.close(); env.close(); vol.close()


def test_frame_reset_clears_everything_sparse_or_full(vr_ctx, both):
    """vr_renderer_reset_cache zeroes only cache[hit[pix]] when every trace since the last reset used one camera / row window,
    and the whole cache otherwise: after a reset the cache must be all zero in every case, and results after it must equal a
    fresh renderer's."""
    r, ref, keep = _scene(vr_ctx, 48, 96, 64)
    pos, d = synth.default_camera(48)
    pos2 = pos + np.array([5.0, 2.0, -3.0], dtype=np.float32)
    seeds = synth.glibc_rand(4)
    r.render_frames(pos, d, seeds)                      # one camera -> sparse reset
    assert r.cache_download().any()
    r.reset_cache()
    assert not r.cache_download().any()
    r.reset_cache()                                     # nothing touched since: no-op, still zero
    assert not r.cache_download().any()
    r.render_frame(pos, d, 1); r.render_frame(pos2, d, 2)   # two cameras -> full reset
    r.reset_cache()
    assert not r.cache_download().any()
    r.set_rows(8, 40); r.render_frame(pos, d, 3); r.set_rows(0, 64); r.render_frame(pos, d, 4)   # two row windows -> full reset
    r.reset_cache()
    assert not r.cache_download().any()
    got = r.render_frames(pos, d, seeds)
    for s in seeds:
        want = ref.render_frame(pos, d, s)
    assert np.array_equal(r.cache_download().reshape(-1, 4)[:, 3], ref.cache.reshape(-1, 4)[:, 3])
    assert np.array_equal(got[..., 3], want[..., 3]) and _psnr(got[..., :3], want[..., :3]) >= 45.0
    r.close(); [k.close() for k in keep]


# ---- 2-D frame filter: 2d_image_filter.cl -----------------------------------------------------------------------
def test_image_filter2d_reference_mode_bit_exact(vr_ctx):
    g = np.load(os.path.join(GOLDEN, "ref_filter2d.npz"))
    for key in g.files:  # outputs of the reference kernel itself (k = 9 takes the untiled path)
        if key.startswith("out_"):
            _, name, k, s = key.split("_")
            got = vr_ctx.image_filter(g["in_" + name], int(k[1:]), float(s[1:]), api.VR_FILTER2D_REFERENCE)
            assert np.array_equal(got, g[key]), key
    rs = np.random.default_rng(2)
    for (h, w, k, sigma) in [(61, 97, 2, 1.5), (8, 200, 8, 3.0), (130, 33, 4, 0.7), (3, 3, 0, 1.0), (70, 70, 20, 9.0)]:
        img = rs.integers(0, 256, (h, w, 4), dtype=np.uint8)
        img[: h // 2, : w // 4] = img[0, 0]
        got = vr_ctx.image_filter(img, k, sigma, api.VR_FILTER2D_REFERENCE)
        assert np.array_equal(got, o.image_filter2d(img, k, sigma)), (h, w, k, sigma)


def test_image_filter2d_bilateral_mode(vr_ctx):
    # expf differs by an ulp between CUDA and glibc: a rounded channel may move by one count
    rs = np.random.default_rng(4)
    for (h, w, k, sigma) in [(61, 97, 2, 12.0), (40, 64, 8, 30.0), (33, 130, 12, 6.0), (9, 9, 0, 1.0), (16, 16, 15, 50.0)]:
        img = rs.integers(0, 256, (h, w, 4), dtype=np.uint8)
        got = vr_ctx.image_filter(img, k, sigma, api.VR_FILTER2D_BILATERAL)
        want = o.image_bilateral2d(img, k, sigma)
        d = np.abs(got.astype(int) - want.astype(int))
        assert d.max() <= 1 and (d == 0).mean() >= 0.99, (h, w, k, sigma, d.max(), (d == 0).mean())
        assert np.array_equal(got[..., 3], img[..., 3])


def test_renderer_filter_frame_and_argument_checks(vr_ctx):
    n, W, H = 48, 96, 64
    vol, env = api.Volume(vr_ctx, synth.synth_ct(n)), api.EnvMap(vr_ctx, synth.synth_env(128, 64))
    r = api.Renderer(vr_ctx, W, H)
    r.image_set(vol, env)
    r.set_tf(synth.default_tf())
    r.flush_changes()
    pos, d = synth.default_camera(n)
    frame = r.render_frame(pos, d, 1)
    got = r.filter_frame(2, 1.5)
    assert np.array_equal(got, o.image_filter2d(frame, 2, 1.5))
    assert np.array_equal(r.render_frame(pos, d, 2)[..., 3], frame[..., 3])  # the next frame overwrites the filtered one
    for bad in [(-1, 1.0, 0), (65, 1.0, 0), (16, 1.0, 1), (2, 0.0, 0), (2, float("nan"), 0), (2, 1.0, 7)]:
        with pytest.raises(api.VrError):
            r.filter_frame(*bad)
    r.close(); env.close(); vol.close()


# ---- sampling modes -------------------------------------------------------------------------------------------------------
def test_sampling_mode_switch_and_texture_lifecycle(vr_ctx):
    """VR_SAMPLING_HW_LINEAR needs a flush (textures), renders, differs from the NEAREST reading on interpolated data, and switching
    back restores the NEAREST results exactly (parity of the linear mode itself: tests/test_ref_opencl_gpu.py)."""
    n, W, H = 48, 96, 64
    v, envimg, tf = synth.synth_ct(n), synth.synth_env(128, 64), synth.default_tf()
    vol, env = api.Volume(vr_ctx, v), api.EnvMap(vr_ctx, envimg)
    r = api.Renderer(vr_ctx, W, H)
    r.image_set(vol, env)
    r.set_tf(tf)
    r.flush_changes()
    pos, d = synth.default_camera(n)
    seeds = synth.glibc_rand(3)
    for s in seeds:
        near_frame = r.render_frame(pos, d, s)
    near_cache = r.cache_download()
    r.set_sampling(api.VR_SAMPLING_HW_LINEAR)
    with pytest.raises(api.VrError):
        r.render_frame(pos, d, seeds[0])  # no textures yet
    r.flush_changes()
    for s in seeds:
        lin_frame = r.render_frame(pos, d, s)
    lin_cache = r.cache_download()
    assert lin_cache.any() and not np.array_equal(lin_cache, near_cache)
    assert (lin_frame[..., 3] == near_frame[..., 3]).mean() > 0.9  # same silhouette up to the interpolation at the surface
    with pytest.raises(api.VrError):
        r.set_sampling(7)
    r.set_sampling(api.VR_SAMPLING_NEAREST)
    r.flush_changes()
    for s in seeds:
        again = r.render_frame(pos, d, s)
    assert np.array_equal(r.cache_download(), near_cache) and np.array_equal(again, near_frame)
    r.close(); env.close(); vol.close()


@pytest.mark.parametrize("n,W,H,frames,cam", [(64, 160, 120, 4, "default"), (96, 200, 136, 2, "closeup")])
def test_hw_linear_sampling_matches_oracle_model(vr_ctx, both, n, W, H, frames, cam):
    """VR_SAMPLING_HW_LINEAR (texture unit) against the oracle's model of that filter (oracle.cpp hw_linear_fetch, pinned bit-exactly
    against 874 545 samples of NVIDIA's OpenCL runtime, tests/test_ref_pinning_cpu.py).  The volume reads decide which voxels are hit
    and which samples are admitted: those must be identical.  The bilinear environment lookup of the oracle models the scaling of
    the normalised coordinate as an exact product (not probed separately); measured: the whole cache is identical."""
    v, envimg, tf = synth.synth_ct(n), synth.synth_env(128, 64), synth.default_tf()
    pos, d = synth.default_camera(n) if cam == "default" else synth.closeup_camera(n)
    seeds = synth.glibc_rand(frames)
    o.set_sampling(1)
    try:
        ref = o.Renderer(v, envimg, tf, W, H)
        for s in seeds:
            want = ref.render_frame(pos, d, s)
    finally:
        o.set_sampling(0)
    vol, env = api.Volume(vr_ctx, v), api.EnvMap(vr_ctx, envimg)
    r = api.Renderer(vr_ctx, W, H)
    r.set_sampling(api.VR_SAMPLING_HW_LINEAR)
    r.set_trace_mode(TRACE_MODE[0])
    r.image_set(vol, env)
    r.set_tf(tf)
    r.flush_changes()
    r.enable_counters(True)
    for s in seeds:
        got = r.render_frame(pos, d, s)
    c = r.counters()
    assert [c[k] for k in ("steps", "normals", "env", "primary_hits", "admitted", "samples")] == [int(x) for x in ref.counters]
    a = r.cache_download().astype(np.int32).reshape(-1, 4)
    b = ref.cache.astype(np.int32).reshape(-1, 4)
    assert np.array_equal(a[:, 3], b[:, 3])                      # tokens: same hits, same admissions
    assert np.array_equal(got[..., 3], want[..., 3])             # hit / miss classification of every pixel
    diff = np.abs(a[:, :3] - b[:, :3])
    assert (diff <= np.maximum(b[:, 3:4], 1)).all()              # at most one count per admitted sample
    touched = b[:, 3] > 0
    print("hw-linear vs oracle model: colour lanes identical", float((diff[touched] == 0).mean()), "max diff", int(diff.max()))
    assert (diff[touched] == 0).mean() >= 0.99                   # measured on B200: 1.0, max diff 0, on both scenes
    r.close(); env.close(); vol.close()


def test_volume_kernels_hw_linear_match_oracle_model(vr_ctx):
    """fetch_stats, tf_sort_values and bilateral_filter under vr_volume_set_sampling(VR_SAMPLING_HW_LINEAR) against the oracle's
    restatement with the hardware-filter model (orc_*_shipped), which tests/test_ref_pinning_cpu.py pins against the outputs of the
    reference's OpenCL kernels as shipped, recorded on the B200 (same ragged volume)."""
    v = synth.synth_ct(0, dims=(45, 37, 29))
    vol = api.Volume(vr_ctx, v)
    near_stats = vol.stats()
    vol.set_sampling(api.VR_SAMPLING_HW_LINEAR)
    assert vol.stats() == o.fetch_stats_shipped(v, 1)
    assert vol.stats() != near_stats
    rng = [float(x) for x in near_stats]
    assert np.array_equal(vol.histogram(500, 500, rng), o.histogram_shipped(v, 1, 500, 500, rng))
    # the table path under this reading (k_histogram_lut<LINEAR>): clipped ranges, a coarse grid, rows below gradient 0
    for (w, h, r2) in [(64, 64, [-2000.0, 3000.0, 0.0, 4000.0]), (31, 17, [0.0, 100.0, 5.0, 50.0]), (100, 300, [200.0, 1200.0, 20.0, 900.0])]:
        assert np.array_equal(vol.histogram(w, h, r2), o.histogram_shipped(v, 1, w, h, r2)), (w, h, r2)
    v8 = synth.synth_ct(3, dims=(64, 24, 20))   # rows of whole octets: the last octet of a row ends at the volume's face
    vol8 = api.Volume(vr_ctx, v8)
    st8 = [float(x) for x in vol8.stats()]
    vol8.set_sampling(api.VR_SAMPLING_HW_LINEAR)
    assert np.array_equal(vol8.histogram(256, 128, st8), o.histogram_shipped(v8, 1, 256, 128, st8))
    vol8.close()
    vol.filter()
    got, want = vol.download(), o.bilateral_shipped(v)
    dd = np.abs(got.astype(np.int32) - want.astype(np.int32))
    assert dd.max() <= 1 and (dd == 0).mean() >= 0.99
    vol.set_sampling(api.VR_SAMPLING_NEAREST)   # back: the NEAREST stats of the (now filtered) volume
    assert vol.stats() == o.fetch_stats(got)
    vol.close()


def test_hw_linear_quiet_field_equals_oracle_analysis(vr_ctx):
    """The step field of the hw-linear path (vr_quiet.cu): bit (ux + 2 uy + 4 uz) of voxel cell (x, y, z) must be the oracle's verdict
    for the hardware cell (x-1+ux, y-1+uy, z-1+uz) — orc_quiet_cells stores cell c at index c+1 — for three transfer functions."""
    v = synth.synth_ct(0, dims=(45, 37, 29))
    envimg = synth.synth_env(64, 32)
    tf2 = [{"min_v": 900.0, "max_v": 1200.0, "min_g": 0.0, "max_g": 0.0, "flags": 0, "rgba": (255, 64, 32, 128)},
           {"min_v": 500.0, "max_v": 1500.0, "min_g": 100.0, "max_g": 2000.0, "flags": 1, "rgba": (40, 200, 255, 255)}]
    vol, env = api.Volume(vr_ctx, v), api.EnvMap(vr_ctx, envimg)
    r = api.Renderer(vr_ctx, 64, 48)
    r.set_sampling(api.VR_SAMPLING_HW_LINEAR)
    r.image_set(vol, env)
    nz, ny, nx = v.shape
    for tf in (synth.default_tf(), tf2, synth.threshold_tf(800)):
        r.set_tf(tf)
        r.flush_changes()
        got = r.quiet_download()
        q = o.quiet_cells(v, tf)  # [nz+1, ny+1, nx+1]
        want = np.zeros_like(got)
        for oct_ in range(8):
            ux, uy, uz = oct_ & 1, (oct_ >> 1) & 1, oct_ >> 2
            want |= (q[uz:uz + nz, uy:uy + ny, ux:ux + nx] << oct_).astype(np.uint8)
        assert np.array_equal(got, want)
        assert 0.2 < (got == 255).mean() < 1.0
    r.close(); env.close(); vol.close()


def test_hw_linear_batch_thresholds_and_gradient_clauses(vr_ctx, both):
    """hw-linear path: a 16-frame batch (primary reuse inside one call), a TF with a gradient clause and a colourless threshold
    TF — tokens, hit / miss classification and the per-sample counters equal the oracle's model; colour lanes within one count."""
    n, W, H = 64, 128, 96
    v, envimg = synth.synth_ct(n), synth.synth_env(128, 64)
    tf2 = [{"min_v": 900.0, "max_v": 1200.0, "min_g": 0.0, "max_g": 0.0, "flags": 0, "rgba": (255, 64, 32, 128)},
           {"min_v": 500.0, "max_v": 1500.0, "min_g": 100.0, "max_g": 2000.0, "flags": 1, "rgba": (40, 200, 255, 255)}]
    vol, env = api.Volume(vr_ctx, v), api.EnvMap(vr_ctx, envimg)
    r = api.Renderer(vr_ctx, W, H)
    r.set_sampling(api.VR_SAMPLING_HW_LINEAR)
    r.set_trace_mode(TRACE_MODE[0])
    r.image_set(vol, env)
    for tf, cam, frames in ((tf2, synth.closeup_camera(n), 16), (synth.threshold_tf(800), synth.default_camera(n), 3)):
        seeds = synth.glibc_rand(frames)
        o.set_sampling(1)
        try:
            ref = o.Renderer(v, envimg, tf, W, H)
            for s in seeds:
                want = ref.render_frame(cam[0], cam[1], s)
        finally:
            o.set_sampling(0)
        r.set_tf(tf)
        r.flush_changes()
        r.enable_counters(True)
        r.counters(reset=True)
        got = r.render_frames(cam[0], cam[1], seeds)
        c = r.counters(reset=True)
        r.enable_counters(False)
        assert [c[k] for k in ("steps", "normals", "env", "primary_hits", "admitted", "samples")] == [int(x) for x in ref.counters]
        a = r.cache_download().astype(np.int32).reshape(-1, 4)
        b = ref.cache.astype(np.int32).reshape(-1, 4)
        assert np.array_equal(a[:, 3], b[:, 3])
        assert np.array_equal(got[..., 3], want[..., 3])
        diff = np.abs(a[:, :3] - b[:, :3])
        assert (diff <= np.maximum(b[:, 3:4], 1)).all()
        touched = b[:, 3] > 0
        assert not touched.any() or (diff[touched] == 0).mean() >= 0.99
    r.close(); env.close(); vol.close()


# ---- RNG (row a9): the device functions of the trace kernels, known answers ------------------------------------------------------
def test_device_rng_known_answers(vr_ctx):
    """utility_sampling.cl:13-21,40-50 on the DEVICE: the integer triples (ra_x, ra_y, ra_z) and components over a (seed, gid) grid
    equal SURVEY A.4's known answers and the oracle bit for bit; the sampled directions equal the oracle's (same fp32 operation
    order, no FMA, one shared reciprocal that is checked to give the correctly rounded quotients)."""
    s0 = 1804289383
    seeds, gids = [s0 + 1, s0 + 1], [(0, 0), (959, 539)]
    rng = np.random.default_rng(5)
    for s in synth.glibc_rand(6) + [0, -1, 2 ** 31 - 1, -2 ** 31]:
        for o_ in (1, 2, 9, 12):
            for g in [(0, 0), (1919, 1079), (3839, 2159), (7, 5), (5, 7)] + [tuple(int(x) for x in rng.integers(0, 4096, 2)) for _ in range(40)]:
                seeds.append(int(np.int32(np.uint32((s + o_) & 0xFFFFFFFF))))
                gids.append(g)
    n = len(seeds)
    nr = np.zeros((n, 4), np.float32)
    nrm = rng.normal(size=(n, 3)).astype(np.float32)
    nr[:, :3] = nrm / np.linalg.norm(nrm, axis=1, keepdims=True).astype(np.float32)
    nr[:, 3] = rng.choice(np.array([1.0, 0.5, 128.0 / 255.0, 0.0, 76.0 / 255.0], np.float32), n)
    ra, comp, dirs = vr_ctx.rng_dump(seeds, gids, nr)
    # SURVEY.md A.4 known answers
    assert ra[0].tolist() == [1742764452, 1176753706, 884373136] and comp[0].tolist() == [-604, 554, 656]
    assert ra[1].tolist() == [1954444043, -1525972536, -2110478077] and comp[1].tolist() == [-245, -2616, -2813]
    for i in range(n):
        wra, wcomp = o.rng_triple(seeds[i], gids[i][0], gids[i][1])
        assert ra[i].tolist() == list(wra) and comp[i].tolist() == list(wcomp)
        want = o.hemisphere(nr[i, :3], seeds[i], float(nr[i, 3]), gids[i][0], gids[i][1])
        assert np.array_equal(dirs[i].view(np.uint32), np.asarray(want, np.float32).view(np.uint32)), (i, dirs[i], want)
    assert comp.min() >= -3071 and comp.max() <= 1023


# ---- API state checks ------------------------------------------------------------------------------------------
def test_filter_frame_leaves_the_traced_frame_alone_with_primary_reuse(vr_ctx):
    """vr_renderer_filter_frame writes to its own buffer: with primary reuse across calls the environment pixels are written once
    per camera, so filtering in place would leave a blurred background in every later frame."""
    n, W, H = 48, 96, 64
    v, envimg, tf = synth.synth_ct(n), synth.synth_env(128, 64), synth.default_tf()
    vol, env = api.Volume(vr_ctx, v), api.EnvMap(vr_ctx, envimg)
    r = api.Renderer(vr_ctx, W, H)
    r.image_set(vol, env); r.set_tf(tf); r.set_primary_reuse(2); r.flush_changes()
    pos, d = synth.default_camera(n)
    seeds = synth.glibc_rand(3)
    f0 = r.render_frame(pos, d, seeds[0])
    filt = r.filter_frame(3, 2.0, api.VR_FILTER2D_BILATERAL)
    assert not np.array_equal(filt, f0)
    f1 = r.render_frame(pos, d, seeds[1])          # same camera: k_primary is skipped, env pixels are NOT rewritten
    bg = f0[..., 3] == 200
    assert bg.any() and np.array_equal(f1[bg], f0[bg])
    ref = o.Renderer(v, envimg, tf, W, H)
    for s in seeds[:2]:
        want = ref.render_frame(pos, d, s)
    assert np.array_equal(f1[..., 3], want[..., 3]) and _psnr(f1[..., :3], want[..., :3]) >= 45.0
    r.close(); env.close(); vol.close()


def test_trace_after_clip_or_filter_requires_a_flush(vr_ctx):
    """SDF, cache and hit buffer belong to the flushed scene: after vr_volume_clip / vr_volume_filter / set_scene the renderer
    refuses to trace or resolve until it is flushed (the reference always flushes after set_clipping, ui.cpp:273-278)."""
    n, W, H = 48, 64, 48
    v, envimg, tf = synth.synth_ct(n), synth.synth_env(64, 32), synth.default_tf()
    vol, env = api.Volume(vr_ctx, v), api.EnvMap(vr_ctx, envimg)
    r = api.Renderer(vr_ctx, W, H)
    r.image_set(vol, env); r.set_tf(tf); r.flush_changes()
    pos, d = synth.default_camera(n)
    r.render_frame(pos, d, 1)
    vol.clip((0, 0, 0), (40, 48, 48))
    for call in (lambda: r.render_frame(pos, d, 2), lambda: r.render_frames(pos, d, [3, 4]), lambda: r.resolve(), lambda: r.xchg_gather()):
        with pytest.raises(api.VrError, match="flush required"):
            call()
    r.flush_changes()
    r.render_frame(pos, d, 2)
    vol.filter()
    with pytest.raises(api.VrError, match="flush required"):
        r.render_frame(pos, d, 3)
    r.flush_changes()
    got = r.render_frame(pos, d, 3)
    ref = o.Renderer(vol.download(), envimg, tf, W, H)
    want = ref.render_frame(pos, d, 3)
    assert np.array_equal(got[..., 3], want[..., 3])
    vol2 = api.Volume(vr_ctx, v)
    r.image_set(vol2, env)
    with pytest.raises(api.VrError, match="flush required"):
        r.render_frame(pos, d, 4)
    r.close(); env.close(); vol.close(); vol2.close()


def test_tf_colours_outside_0_255_are_rejected(vr_ctx):
    v, envimg = synth.synth_ct(32), synth.synth_env(64, 32)
    vol, env = api.Volume(vr_ctx, v), api.EnvMap(vr_ctx, envimg)
    r = api.Renderer(vr_ctx, 32, 32)
    for bad in ((256, 0, 0, 0), (0, -1, 0, 0), (0, 0, 0, 1000)):
        tf = [{"min_v": 500.0, "max_v": 1200.0, "min_g": 0.0, "max_g": 0.0, "flags": 0, "rgba": bad}]
        with pytest.raises(api.VrError, match="outside"):
            r.set_tf(tf)
        with pytest.raises(api.VrError, match="outside"):
            api.Sdf(vr_ctx, vol, tf)
    r.close(); env.close(); vol.close()


def test_array_cache_stays_bounded_over_many_clip_boxes(vr_ctx):
    """the context recycles the 3-D arrays (SDF surface, hw-linear step field and volume texture) by size and keeps at most four unused"""
    v, envimg, tf = synth.synth_ct(48), synth.synth_env(64, 32), synth.default_tf()
    vol, env = api.Volume(vr_ctx, v), api.EnvMap(vr_ctx, envimg)
    r = api.Renderer(vr_ctx, 32, 32)
    r.image_set(vol, env); r.set_tf(tf)
    base = vr_ctx.array_count
    for k in range(12):
        vol.clip((0, 0, 0), (24 + 2 * k, 40, 40))
        r.flush_changes()
        assert vr_ctx.array_count <= base + 5   # the renderer's current SDF + at most four cached
    r.set_sampling(api.VR_SAMPLING_HW_LINEAR)
    for k in range(6):
        vol.clip((0, 0, 0), (24 + 2 * k, 40, 40))
        r.flush_changes()
        assert vr_ctx.array_count <= base + 7   # + the 16-bit step field and the volume texture array
    r.close(); env.close(); vol.close()


def test_incremental_host_frame_pull_equals_full_pull(vr_ctx):
    """Pulls into the renderer-owned host frame copy only the bounding box of the shaded pixels while the primary records stay valid
    (vr_renderer_set_primary_reuse(r, 2)); the buffer must equal a full pull after every call — across camera changes, a frame
    filter into the same buffer, a flush and a cache reset."""
    n, W, H = 48, 200, 120
    v, envimg, tf = synth.synth_ct(n), synth.synth_env(128, 64), synth.default_tf()
    vol, env = api.Volume(vr_ctx, v), api.EnvMap(vr_ctx, envimg)
    r = api.Renderer(vr_ctx, W, H)
    r.image_set(vol, env); r.set_tf(tf); r.set_primary_reuse(2); r.flush_changes()
    hf = r.host_frame()
    cams = [synth.default_camera(n), synth.closeup_camera(n), ((500.0, 500.0, 500.0), (0.577, 0.577, 0.577))]  # the last one sees no voxel
    seeds = synth.glibc_rand(12)
    k = 0
    for rep in range(2):
        for pos, d in cams:
            for j in range(3):
                r.render_frame(pos, d, seeds[k % 12], out=hf); k += 1
                full = r.resolve()                       # full pull into a separate buffer
                assert np.array_equal(hf, full), (rep, j)
            if rep == 0:
                r.filter_frame(2, 2.0, api.VR_FILTER2D_BILATERAL, readback=False)
                api._check(api.lib().vr_renderer_filter_frame(r.h, 2, api.C.c_float(2.0), api.VR_FILTER2D_BILATERAL, api._vp(hf)))
                r.render_frame(pos, d, seeds[k % 12], out=hf); k += 1   # after a filter into the host frame: a full pull again
                assert np.array_equal(hf, r.resolve())
        r.reset_cache()
        r.flush_changes()
    r.close(); env.close(); vol.close()


def test_hw_linear_fetch_equals_oracle_model_incl_ties_and_large_values(vr_ctx):
    """The value the hw-linear path reads at a float position (texture unit + tex_value) against the oracle's model of the hardware
    filter (hw_linear_fetch, pinned bit for bit on 874 545 samples of NVIDIA's OpenCL runtime): random positions, positions on the
    1/256 grid (where the rounding ties live), positions outside the volume (border 0), and a volume with values up to +-32767
    (the fp64 route of tex_value)."""
    rng = np.random.default_rng(11)
    envimg = synth.synth_env(64, 32)
    env = api.EnvMap(vr_ctx, envimg)
    for which in ("ct", "extreme"):
        if which == "ct":
            v = synth.synth_ct(0, dims=(45, 37, 29))
        else:
            v = rng.integers(-32768, 32768, size=(19, 23, 31)).astype(np.int16)
        nz, ny, nx = v.shape
        vol = api.Volume(vr_ctx, v)
        r = api.Renderer(vr_ctx, 32, 32)
        r.set_sampling(api.VR_SAMPLING_HW_LINEAR)
        r.image_set(vol, env); r.set_tf(synth.default_tf()); r.flush_changes()
        n = 200000
        dims = np.array([nx, ny, nz], np.float32)
        a = (rng.random((n, 3), dtype=np.float32) * (dims + 4.0) - 2.0).astype(np.float32)          # anywhere, incl. outside
        b = (rng.integers(-256, (dims.max() + 1) * 256, size=(n, 3)) / 256.0).astype(np.float32)     # on the fixed-point grid
        c = (rng.integers(0, dims.max() * 2, size=(n, 3)) / 2.0).astype(np.float32)                  # integer and half positions
        pts = np.concatenate([a, b, c])
        got = r.linear_fetch(pts)
        want = o.hw_linear_fetch(v, pts)
        assert np.array_equal(got, want), (which, int((got != want).sum()))
        r.close(); vol.close()
    env.close()


@pytest.mark.parametrize("sampling", [api.VR_SAMPLING_NEAREST, api.VR_SAMPLING_HW_LINEAR], ids=["nearest", "hw_linear"])
def test_flush_after_a_colour_edit_keeps_the_fields_and_renders_the_new_colours(vr_ctx, sampling):
    """Incremental rebuild: a flush whose transfer function differs from the last one in colours only keeps the SDF
    (and the hw-linear step field); a threshold edit, a new volume generation or a sampling switch rebuilds.  Either way the
    frames are those of a fresh renderer with that transfer function."""
    n, W, H = 48, 96, 64
    v, envimg = synth.synth_ct(n), synth.synth_env(128, 64)
    pos, d = synth.default_camera(n)
    seeds = synth.glibc_rand(3)

    def tf(lo, col):
        return [{"min_v": lo, "max_v": 1200.0, "min_g": 0.0, "max_g": 4000.0, "flags": 0, "rgba": col}]

    def fresh(t):
        vol2, env2 = api.Volume(vr_ctx, v), api.EnvMap(vr_ctx, envimg)
        r2 = api.Renderer(vr_ctx, W, H); r2.set_sampling(sampling); r2.image_set(vol2, env2); r2.set_tf(t); r2.flush_changes()
        f = r2.render_frames(pos, d, seeds); c = r2.cache_download()
        r2.close(); env2.close(); vol2.close()
        return f, c

    vol, env = api.Volume(vr_ctx, v), api.EnvMap(vr_ctx, envimg)
    r = api.Renderer(vr_ctx, W, H); r.set_sampling(sampling); r.image_set(vol, env)
    r.set_tf(tf(500.0, (255, 255, 255, 255))); r.flush_changes()
    assert not r.last_flush_kept_fields
    r.render_frames(pos, d, seeds)
    for t, kept in ((tf(500.0, (255, 64, 32, 128)), True), (tf(500.0, (10, 200, 255, 255)), True), (tf(650.0, (10, 200, 255, 255)), False),
                    (tf(650.0, (255, 255, 255, 40)), True)):
        r.set_tf(t); r.flush_changes()
        assert r.last_flush_kept_fields == kept
        f = r.render_frames(pos, d, seeds)
        wf, wc = fresh(t)
        assert np.array_equal(r.cache_download(), wc) and np.array_equal(f, wf)
    vol.filter()                      # new volume generation: rebuild
    r.flush_changes()
    assert not r.last_flush_kept_fields
    r.close(); env.close(); vol.close()
