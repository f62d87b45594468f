"""host-side breakdown of one pipelined headless job (bench.py e2e): python tools/e2e_probe.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cl_volume_renderer_b200 import api, synth
n=512; W,H=1920,1080
ctx=api.Context(0)
v=synth.synth_ct(n); e=synth.synth_env(2048,1024)
vp=torch.empty(v.shape,dtype=torch.int16,pin_memory=True); vp.numpy()[...]=v
ep=torch.empty(e.shape,dtype=torch.uint8,pin_memory=True); ep.numpy()[...]=e
pos,d=synth.default_camera(n); seeds=synth.glibc_rand(64); tf=api.tf_format(synth.default_tf())
r=api.Renderer(ctx,W,H); hf=r.host_frame()
if len(sys.argv) > 1 and sys.argv[1] == "linear": r.set_sampling(api.VR_SAMPLING_HW_LINEAR)
cur=api.Volume(ctx,vp.numpy(),async_upload=True)
for it in range(6):
    T=[time.perf_counter()]
    def m(): T.append(time.perf_counter())
    en=api.EnvMap(ctx,ep.numpy()); m()
    nxt=api.Volume(ctx,vp.numpy(),async_upload=True); m()
    r.image_set(cur,en); r.next_event_code_set(tf); m()
    r.flush_changes(); m()
    r.render_frames(pos,d,seeds,out=hf); m()
    en.close(); cur.close(); m()
    cur=nxt
    names=["env","async_vol","set","flush","render+readback","close"]
    print(" ".join(f"{k} {1e3*(T[i+1]-T[i]):.2f}" for i,k in enumerate(names)), f"| total {1e3*(T[-1]-T[0]):.2f} ms")
