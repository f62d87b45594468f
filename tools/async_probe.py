import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cl_volume_renderer_b200 import api, synth
n=512; W,H=1920,1080
ctx=api.Context(0)
v=synth.synth_ct(n)
pin=torch.empty(v.shape,dtype=torch.int16,pin_memory=True); pin.numpy()[...]=v
env=api.EnvMap(ctx, synth.synth_env(2048,1024))
pos,d=synth.default_camera(n); seeds=synth.glibc_rand(64)
def T(): ctx.synchronize(); return time.perf_counter()
for it in range(3):
    t0=T(); a=api.Volume(ctx,pin.numpy()); t1=T(); print(f"blocking upload {1e3*(t1-t0):.2f} ms")
    t0=time.perf_counter(); b=api.Volume(ctx,pin.numpy(),async_upload=True); t1=time.perf_counter(); b.wait(); t2=time.perf_counter()
    print(f"async call returns in {1e3*(t1-t0):.2f} ms, wait {1e3*(t2-t1):.2f} ms")
    r=api.Renderer(ctx,W,H); r.image_set(a,env); r.set_tf(synth.default_tf())
    t0=T(); r.flush_changes(); r.render_frames(pos,d,seeds,readback=False); t1=T(); print(f"compute alone {1e3*(t1-t0):.2f} ms")
    t0=T(); c=api.Volume(ctx,pin.numpy(),async_upload=True); tc=time.perf_counter(); r.flush_changes(); r.render_frames(pos,d,seeds,readback=False); ctx.synchronize(); t1=time.perf_counter(); c.wait(); t2=time.perf_counter()
    print(f"async upload + compute: call {1e3*(tc-t0):.2f}, compute done at {1e3*(t1-t0):.2f}, upload done at {1e3*(t2-t0):.2f} ms")
    for x in (r,a,b,c): x.close()
