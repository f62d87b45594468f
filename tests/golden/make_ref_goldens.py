"""Generates tests/golden/ref_kernels.npz by RUNNING THE REFERENCE'S OWN KERNELS (oracle/_ref/libref.so — the .cl sources
of /root/reference compiled for the CPU, see oracle/ref_build/) on small seeded inputs.  Run in the build container:

    make -C oracle ref && python tests/golden/make_ref_goldens.py

The inputs are regenerated from cl_volume_renderer_b200.synth inside the tests, so only outputs are stored."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib as o  # noqa: E402  (only for tf_rects marshalling and the TF specs)
import ref_lib as R  # noqa: E402
from cl_volume_renderer_b200 import synth  # noqa: E402

out = {}
# volume kernels on a ragged volume
v = synth.synth_ct(0, dims=(45, 37, 29))
out["stats"] = np.array(R.fetch_stats(v), dtype=np.int32)
rng = [float(x) for x in out["stats"]]
out["hist_50x40"] = R.histogram(v, 50, 40, rng)
img, rounded, n = R.tf_color_frame(out["hist_50x40"], 50, 40)
out["tf_frame_50x40"] = img
out["tf_levels"] = np.int32(n)
out["bilateral"] = R.bilateral(v)
out["clip"] = R.clip(v, (3, 5, 2), (30, 20, 20))
out["sdf_default"] = R.sdf_build(v, synth.default_tf())[0]
tf_grad = [{"min_v": 100.0, "max_v": 1400.0, "min_g": 50.0, "max_g": 900.0, "flags": 1, "rgba": (255, 0, 0, 128)}]
out["sdf_grad"] = R.sdf_build(v, tf_grad)[0]
# render: 3 frames, single-threaded (deterministic single-phase execution), two transfer functions
n = 48
vol = synth.synth_ct(n)
env = synth.synth_env(128, 64)
pos, d = synth.default_camera(n)
tf2 = [{"min_v": 900.0, "max_v": 1200.0, "min_g": 0.0, "max_g": 0.0, "flags": 0, "rgba": (255, 64, 32, 128)},
       {"min_v": 500.0, "max_v": 1500.0, "min_g": 100.0, "max_g": 2000.0, "flags": 1, "rgba": (40, 200, 255, 255)}]
for name, tf in (("default", synth.default_tf()), ("two_clause", tf2)):
    sdf = R.sdf_build(vol, tf)[0]
    r = R.Renderer(vol, env, tf, 96, 64, sdf)
    for seed in synth.glibc_rand(3):
        frame = r.render_frame(pos, d, seed, threads=1)
    nz = np.flatnonzero(r.cache)
    out[f"render_{name}_frame"] = frame
    out[f"render_{name}_cache_idx"] = nz.astype(np.int64)
    out[f"render_{name}_cache_val"] = r.cache[nz]
np.savez_compressed(os.path.join(HERE, "ref_kernels.npz"), **out)
print({k: (v.shape, str(v.dtype)) for k, v in out.items()})

# 2d_image_filter.cl (dead code in the reference: no host call site) — inputs are stored as well
f2 = {}
rs = np.random.default_rng(20261018)
noise = rs.integers(0, 256, (40, 48, 4), dtype=np.uint8)
noise[:9, :11] = noise[0, 0]          # flat patch: red weight sum 0 -> 0/0
noise[20:30, 5:25, 0] = 200           # flat in red only
env_small = synth.synth_env(64, 32)   # smooth gradients + the bright disc
f2["in_noise"], f2["in_env"] = noise, env_small
for name, img in (("noise", noise), ("env", env_small)):
    for k, sigma in ((1, 1.0), (2, 0.6), (3, 2.5), (9, 4.0)):
        f2[f"out_{name}_k{k}_s{sigma}"] = R.image_filter2d(img, k, sigma)
np.savez_compressed(os.path.join(HERE, "ref_filter2d.npz"), **f2)
print({k: (v.shape, str(v.dtype)) for k, v in f2.items()})
