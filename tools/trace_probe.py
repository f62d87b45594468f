"""A/B timing of the trace schedules on the bench scene: python tools/trace_probe.py [n] [modes]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cl_volume_renderer_b200 import api, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
modes = [int(m) for m in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 2]
W, H = 1920, 1080
ctx = api.Context(0)
s = torch.cuda.ExternalStream(ctx.stream)
vol = api.Volume(ctx, synth.synth_ct(n)); env = api.EnvMap(ctx, synth.synth_env(2048, 1024))
r = api.Renderer(ctx, W, H); r.image_set(vol, env); r.set_tf(synth.default_tf()); r.flush_changes()
seeds = synth.glibc_rand(64)
for cam, (pos, d) in (("default", synth.default_camera(n)), ("closeup", synth.closeup_camera(n))):
    for m in modes:
        r.set_trace_mode(m)
        best = 1e9
        for rep in range(4):
            r.reset_cache(); ctx.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s); r.render_frames(pos, d, seeds, readback=False); e1.record(s); e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(f"{cam:8s} mode {m}: {best:7.2f} ms / 64 spp -> {W*H*64/best/1e3:8.1f} Msamples/s", flush=True)
