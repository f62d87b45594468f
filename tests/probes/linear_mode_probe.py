"""Quick look: CUDA VR_SAMPLING_HW_LINEAR against the reference's OpenCL kernels as shipped, small scenes only."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "probes"))
import numpy as np
import ref_ocl_lib as R
from cl_volume_renderer_b200 import api, synth
from ref_opencl_bench import cache_metrics, frame_metrics
assert R.available(), R.error()
R.set_nearest(False)
ctx = api.Context(0)
tf_src = api.tf_format(synth.default_tf())
for (n, W, H, nf, cam) in [(64, 160, 120, 6, "default"), (96, 200, 136, 3, "closeup"), (128, 320, 240, 16, "default")]:
    v, envimg = synth.synth_ct(n), synth.synth_env(128, 64)
    pos, d = synth.default_camera(n) if cam == "default" else synth.closeup_camera(n)
    seeds = synth.glibc_rand(nf)
    sc = R.Scene(v, envimg, tf_src, W, H)
    of, _ = sc.render(pos, d, seeds)
    oc = sc.cache()
    vol, env = api.Volume(ctx, v), api.EnvMap(ctx, envimg)
    r = api.Renderer(ctx, W, H)
    r.set_sampling(api.VR_SAMPLING_HW_LINEAR)
    r.image_set(vol, env); r.next_event_code_set(tf_src); r.flush_changes()
    for s in seeds:
        cf = r.render_frame(pos, d, s)
    cc = r.cache_download()
    print(json.dumps({"scene": f"{n}^3 {W}x{H} {nf} {cam}", "cache": cache_metrics(oc, cc), "frame": frame_metrics(of, cf),
                      "env_pixels_identical": float((of[of[..., 3] == 200] == cf[of[..., 3] == 200]).mean()) if (of[..., 3] == 200).any() else None}), flush=True)
    r.close(); env.close(); vol.close(); sc.close()
