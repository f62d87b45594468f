"""Where does an interactive frame go?  64 x vr_render_frame on the bench scene with primary reuse across calls (the frame_emitter
loop, renderer.cpp:131-158): wall time per call with the pull into the renderer's host frame, without any pull, and the device time
of the trace / resolve kernels.   python tools/interactive_probe.py [n]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from cl_volume_renderer_b200 import api, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
W, H = 1920, 1080
ctx = api.Context(0)
vol = api.Volume(ctx, synth.synth_ct(n)); env = api.EnvMap(ctx, synth.synth_env(2048, 1024))
seeds = synth.glibc_rand(64 * 3)
for sampling, name, surf in ((api.VR_SAMPLING_NEAREST, "nearest", 1), (api.VR_SAMPLING_NEAREST, "nearest_bricked_ldg", 0), (api.VR_SAMPLING_HW_LINEAR, "hw_linear", 1)):
    for cam, (pos, d) in (("default", synth.default_camera(n)), ("closeup", synth.closeup_camera(n))):
        r = api.Renderer(ctx, W, H)
        r.set_sampling(sampling); r.set_primary_reuse(2); r.set_tuning("surf", surf)
        r.image_set(vol, env); r.set_tf(synth.default_tf()); r.flush_changes()
        hf = r.host_frame()
        other = np.empty((H, W, 4), np.uint8)
        row = {"sampling": name, "camera": cam}
        for label, kw in (("pull_host_frame", {"out": hf}), ("pull_other_buffer", {"out": other}), ("no_pull", {"readback": False})):
            r.reset_cache()
            r.render_frame(pos, d, seeds[0], **kw)   # primary records + first full pull
            ctx.synchronize()
            r.enable_timing(True); r.kernel_times(reset=True)
            t0 = time.perf_counter()
            for k in range(64):
                r.render_frame(pos, d, seeds[1 + k], **kw)
            ctx.synchronize()
            dt = time.perf_counter() - t0
            tr, rs, nf = r.kernel_times(reset=True)
            r.enable_timing(False)
            row[label] = {"ms_per_call": 1e3 * dt / 64, "gsamples": W * H * 64 / dt / 1e9, "trace_ms_per_call": tr / 64, "resolve_ms_per_call": rs / 64}
        print(json.dumps(row), flush=True)
        r.close()
