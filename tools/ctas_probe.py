"""k_trace_pt at 10 / 12 (default) / 14 / 16 CTAs per SM on the bench scene (A/B build: VR_LIB=tools/ab/libvr_ab.so python tools/ctas_probe.py)"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cl_volume_renderer_b200 import api, synth
n=512; W,H=1920,1080
ctx=api.Context(0); st=torch.cuda.ExternalStream(ctx.stream)
vol=api.Volume(ctx,synth.synth_ct(n)); env=api.EnvMap(ctx,synth.synth_env(2048,1024))
r=api.Renderer(ctx,W,H); r.image_set(vol,env); r.set_tf(synth.default_tf()); r.flush_changes()
seeds=synth.glibc_rand(64)
cams={"default":synth.default_camera(n),"closeup":synth.closeup_camera(n)}
def run(cam):
    pos,d=cams[cam]; best=1e9
    for rep in range(4):
        r.reset_cache(); ctx.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(st); r.render_frames(pos,d,seeds,readback=False); e1.record(st); e1.synchronize()
        best=min(best,e0.elapsed_time(e1))
    return best
for c in (0, 10, 14, 16):
    r.set_tuning("pt_ctas", c)
    print(json.dumps({"pt_ctas": c, "default_ms": run("default"), "closeup_ms": run("closeup")}), flush=True)
