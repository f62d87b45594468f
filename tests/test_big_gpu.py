"""BASELINE config 4 sanity (1024^3 volume, 2 GiB; 8 GiB voxel cache; 4K frame, row-tile split) — the sizes the
reference cannot run at all (int32 overflow in nrrd_loader.hpp:18-19 and utility.cl:21) — and the bench's own
configuration (512^3, 1920x1080, batched 64-frame steps), incl. the regime where the token cap is reached inside a batch.
Size-independent properties plus oracle parity on windows; the oracle renders with the GPU-built SDF (pinned bit-exact by
tests/test_parity_gpu.py), and cache entries are compared at the touched voxels only (no 8 GiB transfers)."""
import numpy as np
import pytest

import oracle_lib as o
from cl_volume_renderer_b200 import api, synth

pytestmark = pytest.mark.gpu


def test_1024_cube_4k_rows(vr_ctx):
    n, W, H = 1024, 3840, 2160
    v = synth.synth_ct(512).repeat(2, axis=0).repeat(2, axis=1).repeat(2, axis=2)   # 2 GiB: every voxel of the 512^3 CT as a 2^3 block
    tf = synth.default_tf()
    envimg = synth.synth_env(2048, 1024)
    vol = api.Volume(vr_ctx, v)
    env = api.EnvMap(vr_ctx, envimg)
    r = api.Renderer(vr_ctx, W, H)
    r.image_set(vol, env); r.set_tf(tf); r.flush_changes()
    sdf = r.sdf_download()
    ev = (v >= 500) & (v <= 1200)
    assert np.array_equal(sdf < 0, ev)          # sign(sdf) <=> event, every voxel
    assert np.abs(sdf).max() == 127 and not (sdf == 0).any()
    del ev
    # oracle SDF on the top z-slab (byte offsets beyond 2^31 in the volume): values below the cap cannot depend on anything
    # farther than 127 planes away, so the slab plus a 130-plane apron reproduces them exactly
    z0 = 992
    want_slab = o.sdf_build(v[z0 - 130:], tf)[0][130:]
    got_slab = sdf[z0:]
    near = np.abs(want_slab) < 127
    assert near.any() and np.array_equal(got_slab[near], want_slab[near])
    # image-tile split: this "rank" traces rows [1000, 1128)
    pos, d = synth.closeup_camera(n)
    r.set_rows(1000, 1128)
    r.enable_counters(True)
    got = r.render_frame(pos, d, 424238335)
    c = r.counters()
    ref = o.Renderer(v, envimg, tf, W, H, sdf=sdf)   # its 8 GiB cache is zero pages until touched
    want = ref.render_frame(pos, d, 424238335, window=(0, 1000, W, 1128))
    assert np.array_equal(got[1000:1128, :, 3], want[1000:1128, :, 3])
    assert (got[1000:1128, :, 3] == 1).mean() > 0.2
    diff = np.abs(got[1000:1128, :, :3].astype(int) - want[1000:1128, :, :3].astype(int))
    assert diff.max() <= 8
    assert c["steps"] == int(ref.counters[0]) and c["primary_hits"] == int(ref.counters[3]) and c["admitted"] == int(ref.counters[4])
    # 64-bit cache indexing lands on the same voxels: entries at the hit voxels are the oracle's, and they hold every admitted token
    hit = r.hit_download()
    assert (hit[:1000] == 0xFFFFFFFF).all() and (hit[1128:] == 0xFFFFFFFF).all()
    vox = np.unique(hit[hit != 0xFFFFFFFF])
    assert vox.max() > 2 ** 29                      # entry offsets beyond 2^32 bytes
    mine = r.cache_download_at(vox).astype(np.int64)
    theirs = ref.cache.reshape(-1, 4)[vox.astype(np.int64)].astype(np.int64)
    assert np.array_equal(mine[:, 3], theirs[:, 3])
    assert (mine == theirs).mean() >= 0.999
    assert int(mine[:, 3].sum()) == c["admitted"] == int(theirs[:, 3].sum())
    r.close(); env.close(); vol.close()


@pytest.fixture(scope="module")
def bench_scene(vr_ctx):
    n, W, H = 512, 1920, 1080
    v, envimg, tf = synth.synth_ct(n), synth.synth_env(2048, 1024), synth.default_tf()
    vol, env = api.Volume(vr_ctx, v), api.EnvMap(vr_ctx, envimg)
    r = api.Renderer(vr_ctx, W, H)
    r.image_set(vol, env); r.set_tf(tf); r.flush_changes()
    sdf = r.sdf_download()
    yield v, envimg, tf, r, sdf, W, H
    r.close(); env.close(); vol.close()


def test_bench_config_batched_step_matches_oracle(vr_ctx, bench_scene):
    """bench.py's own configuration — 512^3, 1920x1080, one batched vr_render_frames call in the default schedule (primary reuse +
    persistent warps) — against the oracle's sequential frames: 8 frames, cap 256 not reached."""
    v, envimg, tf, r, sdf, W, H = bench_scene
    pos, d = synth.default_camera(512)
    seeds = synth.glibc_rand(8)
    r.set_token_cap(256)
    r.reset_cache()
    r.enable_counters(True)
    r.counters(reset=True)
    got = r.render_frames(pos, d, seeds)
    c = r.counters(reset=True)
    r.enable_counters(False)
    ref = o.Renderer(v, envimg, tf, W, H, sdf=sdf)
    for s in seeds:
        want = ref.render_frame(pos, d, s)
    oc = dict(zip(["steps", "normals", "env", "primary_hits", "admitted", "samples"], [int(x) for x in ref.counters]))
    assert c == oc
    hit = r.hit_download()
    vox = np.unique(hit[hit != 0xFFFFFFFF])
    mine = r.cache_download_at(vox).astype(np.int64)
    theirs = ref.cache.reshape(-1, 4)[vox.astype(np.int64)].astype(np.int64)
    assert np.array_equal(mine[:, 3], theirs[:, 3]) and int(theirs[:, 3].sum()) == oc["admitted"] and theirs[:, 3].max() < 256
    assert (mine == theirs).mean() >= 0.999 and np.abs(mine - theirs).max() <= 64
    assert np.array_equal(got[..., 3], want[..., 3])
    mse = np.mean((got[..., :3].astype(np.float64) - want[..., :3].astype(np.float64)) ** 2)
    assert mse == 0 or 10 * np.log10(255.0 ** 2 / mse) >= 45.0


def test_bench_config_cap_reached_inside_the_batch(vr_ctx, bench_scene):
    """The regime of an 8-rank spp split: per-rank token cap 32, reached inside one 64-frame batch.  Token counts are
    min(cap, samples offered) whatever the order — exact.  WHICH samples a saturated voxel admits depends on the order (the
    reference has the same freedom inside a frame), so its colour sums are compared as means: two admission orders of the oracle
    itself differ by 6.5 on average, 25 at the 99th percentile, 44 at most (of 255; measured at 96^3) — the bound below."""
    v, envimg, tf, r, sdf, W, H = bench_scene
    pos, d = synth.default_camera(512)
    seeds = synth.glibc_rand(64)
    cap = 32
    r.set_token_cap(cap)
    r.reset_cache()
    got = r.render_frames(pos, d, seeds)
    ref = o.Renderer(v, envimg, tf, W, H, token_cap=cap, sdf=sdf)
    for s in seeds:
        want = ref.render_frame(pos, d, s)
    hit = r.hit_download()
    vox = np.unique(hit[hit != 0xFFFFFFFF])
    mine = r.cache_download_at(vox).astype(np.int64)
    theirs = ref.cache.reshape(-1, 4)[vox.astype(np.int64)].astype(np.int64)
    assert np.array_equal(mine[:, 3], theirs[:, 3])                    # tokens: exact
    assert int(theirs[:, 3].sum()) == int(ref.counters[4]) and theirs[:, 3].max() == cap
    sat = theirs[:, 3] == cap
    assert sat.mean() > 0.5
    if (~sat).any():
        assert (mine[~sat] == theirs[~sat]).mean() >= 0.999            # below the cap every sample was admitted on both sides
    dm = np.abs(mine[sat, :3] / float(cap) - theirs[sat, :3] / float(cap))
    assert dm.mean() <= 12.0 and np.percentile(dm, 99) <= 40.0 and dm.max() <= 96.0
    assert np.array_equal(got[..., 3], want[..., 3])
    r.set_token_cap(256)
