"""First-contact probe on the GPU box: timings of every kernel family at 256^3 / 512^3 with CUDA events (torch)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cl_volume_renderer_b200 import api, synth

def timed(ctx, fn, n=1):
    s = torch.cuda.ExternalStream(ctx.stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.synchronize()
    e0.record(s)
    for _ in range(n): fn()
    e1.record(s); e1.synchronize()
    return e0.elapsed_time(e1) / n

ctx = api.Context(0)
for n, W, H in [(256, 640, 480), (512, 1920, 1080)]:
    t = time.time(); v = synth.synth_ct(n); print(f"[{n}] synth {time.time()-t:.1f}s", flush=True)
    t = time.time(); vol = api.Volume(ctx, v); print(f"[{n}] upload+stats {1e3*(time.time()-t):.1f} ms", vol.stats())
    env = api.EnvMap(ctx, synth.synth_env(2048, 1024))
    for tfname, tf in [("default", synth.default_tf()), ("thr800", synth.threshold_tf(800))]:
        t = time.time(); s = api.Sdf(ctx, vol, tf); dt = time.time() - t
        print(f"[{n}] sdf build {tfname}: {1e3*dt:.2f} ms wall, levels {s.levels}"); s.close()
    r = api.Renderer(ctx, W, H); r.image_set(vol, env); r.set_tf(synth.default_tf())
    t = time.time(); r.flush_changes(); ctx.synchronize(); print(f"[{n}] flush {1e3*(time.time()-t):.2f} ms")
    pos, d = synth.default_camera(n)
    seeds = synth.glibc_rand(64)
    r.enable_counters(True)
    r.render_frames(pos, d, seeds[:8], readback=False); c = r.counters(reset=True); r.enable_counters(False)
    S = c["samples"]; print(f"[{n}] counters/sample: steps {c['steps']/S:.2f} normals {c['normals']/S:.3f} env {c['env']/S:.3f} hits {c['primary_hits']/S:.3f} admitted {c['admitted']/S:.3f}")
    B = (15*c['steps'] + 12*c['normals'] + 4*c['env'] + 18*c['primary_hits'] + 16*c['admitted'])/S + 4
    for rep in range(3):
        r.reset_cache()
        ms = timed(ctx, lambda: r.render_frames(pos, d, seeds, readback=False))
        print(f"[{n}] 64 spp {W}x{H}: {ms:.2f} ms  -> {W*H*64/ms/1e3:.1f} Msamples/s, alg bytes/sample {B:.1f} -> {B*W*H*64/ms/1e6:.1f} GB/s")
    ms = timed(ctx, lambda: r.render_frames(pos, d, seeds, readback=False))
    print(f"[{n}] next 64 spp (cache 64..128 tokens): {ms:.2f} ms -> {W*H*64/ms/1e3:.1f} Msamples/s")
    ms = timed(ctx, lambda: r.reset_cache(), 5); print(f"[{n}] cache reset {ms:.3f} ms -> {8*n**3/ms/1e6:.0f} GB/s")
    t = time.time(); f = r.render_frame(pos, d, 1); print(f"[{n}] render_frame + readback {1e3*(time.time()-t):.2f} ms; shaded px {(f[...,3]==1).mean():.3f}")
    t = time.time(); img = r.render_tf(500, 500); print(f"[{n}] render_tf {1e3*(time.time()-t):.2f} ms")
    t = time.time(); vol.filter(); ctx.synchronize(); print(f"[{n}] bilateral {1e3*(time.time()-t):.2f} ms")
    r.close(); env.close(); vol.close()
