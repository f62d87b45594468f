"""A/B timing of the hw-linear trace schedules on the bench scene (tuning keys of vr_renderer_set_tuning; pt_ctas needs
VR_LIB=tools/ab/libvr_ab.so):   python tools/lin_probe.py [n] [quick]
Prints one JSON line per setting: ms per 64-spp step (reset cache) for the default and the close-up camera."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from cl_volume_renderer_b200 import api, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
quick = len(sys.argv) > 2 and sys.argv[2] == "quick"
W, H = 1920, 1080
ctx = api.Context(0)
s = torch.cuda.ExternalStream(ctx.stream)
vol = api.Volume(ctx, synth.synth_ct(n)); env = api.EnvMap(ctx, synth.synth_env(2048, 1024))
seeds = synth.glibc_rand(64)
cams = {"default": synth.default_camera(n), "closeup": synth.closeup_camera(n)}
ab = "libvr_ab" in api.LIB_PATH


def measure(r, pos, d, reps=3):
    best = 1e9
    for _ in range(reps):
        r.reset_cache(); ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s); r.render_frames(pos, d, seeds, readback=False); e1.record(s); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


for sampling, name in ((api.VR_SAMPLING_NEAREST, "nearest"), (api.VR_SAMPLING_HW_LINEAR, "hw_linear")):
    r = api.Renderer(ctx, W, H)
    r.set_sampling(sampling)
    r.image_set(vol, env); r.set_tf(synth.default_tf())
    ctx.synchronize()
    import time
    t0 = time.perf_counter(); r.flush_changes(); ctx.synchronize(); flush1 = 1e3 * (time.perf_counter() - t0)
    t0 = time.perf_counter(); r.flush_changes(); ctx.synchronize(); flush2 = 1e3 * (time.perf_counter() - t0)
    for mode in ((2, 3, 1, 0) if ab else (2, 1, 0)):
        r.set_trace_mode(mode)
        row = {"sampling": name, "mode": mode, "flush_ms_first": flush1, "flush_ms_again": flush2}
        for cam, (pos, d) in cams.items():
            row[cam + "_ms"] = measure(r, pos, d)
            row[cam + "_gsamples"] = W * H * 64 / row[cam + "_ms"] / 1e6
        print(json.dumps(row), flush=True)
    if ab:
        r.set_trace_mode(3)
    for k, lv in (((8, 16), (16, 8), (16, 24), (32, 16), (70, 16), (16, 1), (8, 24), (24, 20)) if ab else ()):
        r.set_tuning("sm_k", k); r.set_tuning("sm_leave", lv)
        row = {"sampling": name, "mode": 3, "sm_k": k, "sm_leave": lv}
        for cam, (pos, d) in cams.items():
            row[cam + "_ms"] = measure(r, pos, d, 2)
        print(json.dumps(row), flush=True)
    r.set_tuning("sm_k", 16); r.set_tuning("sm_leave", 16)
    r.set_trace_mode(2)
    if sampling == api.VR_SAMPLING_HW_LINEAR and not quick:
        settings = []
        for w in ((4, 2, 2), (4, 1, 2), (8, 2, 2), (2, 2, 2), (4, 2, 1), (4, 2, 3), (4, 4, 2), (6, 2, 2), (3, 2, 2), (4, 2, 5)):
            settings.append({"lin_sched": 0, "lin_w_fast": w[0], "lin_w_slow": w[1], "lin_w_event": w[2]})
        for w in ((4, 3, 1), (2, 2, 1), (4, 2, 1), (3, 2, 1), (3, 3, 1), (4, 4, 1), (2, 3, 1), (6, 4, 1)):
            settings.append({"lin_sched": 1, "lin_w_fast": w[0], "lin_w_slow": w[1], "lin_w_event": w[2]})
        for spc in (2, 3, 6, 8):
            settings.append({"steps_per_check": spc})
        if ab:
            for c in (6, 8, 10):
                settings.append({"pt_ctas": c})
        base = {"lin_sched": 0, "lin_w_fast": 3, "lin_w_slow": 2, "lin_w_event": 2, "steps_per_check": 4}
        for st in settings:
            for k, v in st.items():
                r.set_tuning(k, v)
            row = dict(st)
            for cam, (pos, d) in cams.items():
                row[cam + "_ms"] = measure(r, pos, d, 2)
            print(json.dumps(row), flush=True)
            for k, v in base.items():
                r.set_tuning(k, v)
            if ab:
                r.set_tuning("pt_ctas", 0)
    if sampling == api.VR_SAMPLING_NEAREST and not quick:
        for st in ({"steps_per_check": 2}, {"steps_per_check": 3}, {"steps_per_check": 6}, {"steps_per_check": 8}, {"rule_a": 3}, {"rule_a": 8},
                   {"rule_a": 3, "steps_per_check": 8}, {"pixel_major": 2}, {"pixel_major": 8}):
            for k, v in st.items():
                r.set_tuning(k, v)
            row = dict(st, sampling="nearest")
            for cam, (pos, d) in cams.items():
                row[cam + "_ms"] = measure(r, pos, d, 2)
            print(json.dumps(row), flush=True)
            r.set_tuning("steps_per_check", 4); r.set_tuning("rule_a", 5); r.set_tuning("pixel_major", 1)
    r.close()
