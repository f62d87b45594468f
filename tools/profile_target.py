"""Small fixed workload for ncu: the bench scene (512^3, 1080p), flush (SDF build + cache reset), then one batch of frames.
Usage: python tools/profile_target.py [frames] [n] [camera: default|close] [sampling: nearest|linear] [trace mode 0|1|2] [key=value tuning ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cl_volume_renderer_b200 import api, synth

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
cam = sys.argv[3] if len(sys.argv) > 3 else "default"
sampling = sys.argv[4] if len(sys.argv) > 4 else "nearest"
mode = int(sys.argv[5]) if len(sys.argv) > 5 else 2
W, H = 1920, 1080
ctx = api.Context(0)
vol = api.Volume(ctx, synth.synth_ct(n))
env = api.EnvMap(ctx, synth.synth_env(2048, 1024))
r = api.Renderer(ctx, W, H)
r.set_sampling(api.VR_SAMPLING_HW_LINEAR if sampling == "linear" else api.VR_SAMPLING_NEAREST)
r.set_trace_mode(mode)
for kv in sys.argv[6:]:
    k, v = kv.split("=")
    r.set_tuning(k, int(v))
r.image_set(vol, env)
r.set_tf(synth.default_tf())
r.flush_changes()
pos, d = synth.default_camera(n) if cam == "default" else synth.closeup_camera(n)
seeds = synth.glibc_rand(frames)
r.render_frames(pos, d, seeds, readback=False)
f = r.render_frame(pos, d, 12345)
print("shaded fraction", float((f[..., 3] == 1).mean()), "launches", ctx.launches)
r.close(); env.close(); vol.close(); ctx.close()
