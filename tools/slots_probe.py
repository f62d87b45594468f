"""k_trace_pt (one sample slot per lane) against k_trace_pt2 (two; A/B build: VR_LIB=tools/ab/libvr_ab.so) on the bench scene: identical voxel caches, ms per 64-frame step.
python tools/slots_probe.py [n]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cl_volume_renderer_b200 import api, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
W, H = 1920, 1080
ctx = api.Context(0)
st = torch.cuda.ExternalStream(ctx.stream)
vol = api.Volume(ctx, synth.synth_ct(n)); env = api.EnvMap(ctx, synth.synth_env(2048, 1024))
r = api.Renderer(ctx, W, H); r.image_set(vol, env); r.set_tf(synth.default_tf()); r.flush_changes()
seeds = synth.glibc_rand(64)
cams = {"default": synth.default_camera(n), "closeup": synth.closeup_camera(n)}
def run(cam):
    pos, d = cams[cam]
    best = 1e9
    for rep in range(4):
        r.reset_cache(); ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); r.render_frames(pos, d, seeds, readback=False); e1.record(st); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
def cache_at_hits(cam):
    pos, d = cams[cam]
    r.reset_cache(); r.render_frames(pos, d, seeds[:16], readback=False)
    hit = r.hit_download().ravel()
    vox = np.unique(hit[hit != 0xFFFFFFFF])
    return vox, r.cache_download_at(vox)
ref = {}
for cam in cams:
    r.set_tuning("pt_slots", 1)
    ref[cam] = cache_at_hits(cam)
    print(json.dumps({"cam": cam, "slots": 1, "ms": run(cam)}), flush=True)
for ctas in (8, 7, 6, 10):
    r.set_tuning("pt_slots", 2); r.set_tuning("pt2_ctas", ctas)
    for cam in cams:
        vox, c = cache_at_hits(cam)
        same = bool(np.array_equal(vox, ref[cam][0]) and np.array_equal(c, ref[cam][1]))
        for rule in ((5, 1), (3, 1), (8, 1)):
            r.set_tuning("rule_a", rule[0]); r.set_tuning("rule_b", rule[1])
            print(json.dumps({"cam": cam, "slots": 2, "ctas": ctas, "rule": rule, "cache_identical": same, "ms": run(cam)}), flush=True)
        r.set_tuning("rule_a", 5); r.set_tuning("rule_b", 1)
