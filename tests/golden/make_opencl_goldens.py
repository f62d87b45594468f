"""Generates tests/golden/opencl_reference_runs.npz by RUNNING THE REFERENCE'S UNMODIFIED OPENCL KERNELS ON THE GPU through the driver's
OpenCL runtime (oracle/_ref/libref_ocl.so, oracle/ref_build/ocl_host.cpp).  Run on a GPU box:

    python tests/golden/make_opencl_goldens.py gpurun_out/opencl_reference_runs.npz      (then copy the file to tests/golden/)

Both readings of the samplers are recorded (tests/test_ref_opencl_gpu.py explains them): "nearest" = CLK_FILTER_LINEAR rewritten to
CLK_FILTER_NEAREST, the filter OpenCL defines for integer images; "shipped" = the text as it is.  Inputs are regenerated from
cl_volume_renderer_b200.synth inside the tests, so only outputs are stored (voxel caches sparsely)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ref_ocl_lib as R  # noqa: E402
from cl_volume_renderer_b200 import api, synth  # noqa: E402  (api only for tf_format: host-side text generation, no GPU call)

assert R.available(), R.error()
out = {"info": np.array(R.info())}
tf_src = api.tf_format(synth.default_tf())
SCENES = {"a": (64, 160, 120, 6, "default"), "b": (96, 200, 136, 3, "closeup")}
for reading, nearest in (("nearest", True), ("shipped", False)):
    R.set_nearest(nearest)
    for key, (n, W, H, frames, cam) in SCENES.items():
        v, envimg = synth.synth_ct(n), synth.synth_env(128, 64)
        pos, d = synth.default_camera(n) if cam == "default" else synth.closeup_camera(n)
        sc = R.Scene(v, envimg, tf_src, W, H)
        frame, _ = sc.render(pos, d, synth.glibc_rand(frames))
        cache = sc.cache()
        nz = np.flatnonzero(cache)
        p = f"{reading}_{key}_"
        out[p + "cache_idx"], out[p + "cache_val"], out[p + "frame"] = nz.astype(np.int64), cache[nz], frame
        out[p + "stats"] = np.array(R.fetch_stats(v)[0], dtype=np.int32)
        if nearest and key == "a":
            out["sdf_a"] = sc.sdf()
        sc.close()
    v = synth.synth_ct(0, dims=(45, 37, 29))
    st, _ = R.fetch_stats(v)
    p = f"{reading}_ragged_"
    out[p + "stats"] = np.array(st, dtype=np.int32)
    # histogram range: the NEAREST-reading stats of the volume in both cases (minima are the volume's own: no negative index)
    rng = [float(x) for x in out["nearest_ragged_stats"]]
    bins, _ = R.histogram(v, 500, 500, rng)
    nzb = np.flatnonzero(bins)
    out[p + "hist_idx"], out[p + "hist_val"] = nzb.astype(np.int32), bins[nzb]
    out[p + "bilateral"] = R.bilateral(v)[0]
    out[p + "clip"] = R.clip(v, (3, 5, 2), (30, 20, 20))[0]
np.savez_compressed(sys.argv[1], **out)
print({k: (a.shape, str(a.dtype)) for k, a in out.items()})
