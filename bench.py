#!/usr/bin/env python
"""bench.py — headline benchmark of the volume path tracer (BASELINE.json: path Msamples/s @1080p 512^3).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One STEP = one pass of the hot path over one batch: cache reset (buffer_reset) + 64 render_frame calls (64 spp,
trace + resolve per frame) at 1920x1080 on the 512^3 synthetic CT volume, starting from a freshly reset voxel cache.
N > 1: spp split (BASELINE config 3, weak scaling): every rank holds the whole scene, renders its own 64 frames with its
own seeds and a token cap of 256/N, then the touched cache entries (one per pixel: all ranks trace the same camera and
therefore hit the same voxels) are summed with one NCCL all-reduce and every rank resolves the frame.  value = N * W*H*64 / max-over-ranks device time.

Keys beyond the base contract: `roofline` (dominant kernel k_trace, algorithmic bytes of SURVEY.md §8d over its
CUDA-event time), `cpu_baseline` (the CPU oracle on a bounded sample of the same workload), `e2e` (the same metric
through the C-ABI with host buffers: volume upload, env bind, flush incl. SDF build, 64 frames each read back).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

VOL_N = 512
W, H = 1920, 1080
SPP = 64
ENV_W, ENV_H = 2048, 1024
WORKLOAD = (f"{VOL_N}^3 short synthetic CT (synth_ct), {W}x{H}, {SPP} spp from a reset voxel cache, default TF rect(500,1200), "
            f"synthetic env {ENV_W}x{ENV_H}, camera (-400,400,-400) look (0.9,6.183)")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe).  The timed region of the default run is
    ~15 ms, shorter than one `nvidia-smi -lms` period, so the clocks are polled through NVML in a thread (every ~1 ms);
    `nvidia-smi --query-gpu` is the fallback when pynvml is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.nvml = None
        self.sm, self.reason_bits, self.stop_flag = [], 0, False
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES may renumber devices: resolve by the CUDA device's PCI bus id
            import torch
            bus = torch.cuda.get_device_properties(index).pci_bus_id if hasattr(torch.cuda.get_device_properties(index), "pci_bus_id") else None
            self.h = None
            if bus is not None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    h = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if pynvml.nvmlDeviceGetPciInfo(h).bus == bus:
                        self.h = h
                        break
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml = pynvml
        except Exception as e:
            log("NVML unavailable, falling back to nvidia-smi:", e)

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
                self.reason_bits |= int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                try:
                    self.reason_bits |= int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                except Exception:
                    pass
            time.sleep(0.001)

    def start(self):
        if self.nvml:
            self.stop_flag = False
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception as e:  # nvidia-smi missing: report it, do not fail the bench
            log("clock sampler unavailable:", e)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml:
            n = self.nvml
            self.stop_flag = True
            self.t.join(timeout=1.0)
            try:
                mx = float(n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM))
            except Exception:
                mx = None
            names = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            reasons = sorted(k for k, bit in names.items() if self.reason_bits & bit)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": mx, "reasons": reasons,
                    "samples": len(self.sm), "source": "NVML polled every 1 ms during the timed region"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


def make_scene():
    from cl_volume_renderer_b200 import synth
    t = time.time()
    vol = synth.synth_ct(VOL_N)
    env = synth.synth_env(ENV_W, ENV_H)
    pos, d = synth.default_camera(VOL_N)
    log(f"scene generated in {time.time() - t:.1f}s")
    return vol, env, pos, d


def alg_bytes_per_sample(c, trace_only):
    """B = 15*S + 12*Hn + 4*E + 18*Hp + 16*A + 4 bytes per sample.  For the trace kernel alone the
    8-byte resolve read per hit pixel belongs to k_resolve: 10*Hp instead of 18*Hp."""
    S = float(c["samples"])
    hp = 10 if trace_only else 18
    return (15 * c["steps"] + 12 * c["normals"] + 4 * c["env"] + hp * c["primary_hits"] + 16 * c["admitted"]) / S + 4


# ------------------------------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path.  The reference has no CPU implementation of its own and its OpenCL
    kernels cannot run here (no OpenCL runtime in the image), so this arm times the CPU oracle — the
    restatement of the reference kernels (oracle/oracle.cpp, OpenMP over pixels) — on the box's host cores.
    Each step = a bounded sample of the workload: ONE frame (1 spp) of the same 1920x1080 / 512^3 scene."""
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1; the reference arm uses every host core it is allowed to (libgomp reads the variable
    # when the library is loaded, which happens below)
    os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as o
    import ref_lib as R
    from cl_volume_renderer_b200 import synth
    vol, env, pos, d = make_scene()
    t = time.time()
    # SDF for the 512^3 scene: built by the oracle (bit-identical to the reference kernels' result — tests/
    # test_ref_pinning_cpu.py — but minutes faster than their 9-TF-evaluations-per-voxel base pass); untimed set-up
    sdf, iters = o.sdf_build(vol, synth.default_tf())
    sdf_s = time.time() - t
    log(f"oracle SDF build {VOL_N}^3: {sdf_s:.1f}s ({iters} iterations, {o.num_threads()} threads)")
    seeds = synth.glibc_rand(args.steps + args.warmup)
    if R.available():
        kind, cores = "reference", R.num_threads()
        r = R.Renderer(vol, env, synth.default_tf(), W, H, sdf)
        frame_fn = lambda s: r.render_frame(pos, d, s, threads=0)  # noqa: E731
        what = "the reference's own ray_marching.cl `render` kernel compiled for the host (oracle/_ref), OpenMP over rows"
    else:
        kind, cores = "port", o.num_threads()
        r = o.Renderer(vol, env, synth.default_tf(), W, H, sdf=sdf)
        frame_fn = lambda s: r.render_frame(pos, d, s)  # noqa: E731
        what = "CPU restatement of the reference kernels (oracle/oracle.cpp), OpenMP over rows"
    for k in range(args.warmup):
        frame_fn(seeds[k])
    t0 = time.perf_counter()
    for k in range(args.steps):
        frame_fn(seeds[args.warmup + k])
    dt = time.perf_counter() - t0
    value = W * H * args.steps / dt / 1e6
    sample = (f"{args.steps} x 1 spp frame of the workload; {what}; SDF built once beforehand on the CPU in {sdf_s:.1f}s, "
              f"untimed")
    print(json.dumps({
        "impl": "reference", "metric": "path_msamples_per_s", "value": value, "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": cores, "kind": kind, "sample": sample,
                         "sdf_build_ms_oracle": 1e3 * sdf_s},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


# ------------------------------------------------------------------------------------------------------------------------
class _DevArray:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from cl_volume_renderer_b200 import api, synth

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = api.Context(local_rank)
    ext = torch.cuda.ExternalStream(ctx.stream, device=local_rank)

    vol_np, env_np, pos, d = make_scene()
    # pinned host copies: the e2e leg uploads from these
    vol_pin = torch.empty(vol_np.shape, dtype=torch.int16, pin_memory=True)
    vol_pin.numpy()[...] = vol_np
    env_pin = torch.empty(env_np.shape, dtype=torch.uint8, pin_memory=True)
    env_pin.numpy()[...] = env_np
    tf_code = api.tf_format(synth.default_tf())
    all_seeds = synth.glibc_rand(SPP * max(world, 1))
    seeds = all_seeds[SPP * rank: SPP * (rank + 1)]

    vol = api.Volume(ctx, vol_pin.numpy())
    env = api.EnvMap(ctx, env_pin.numpy())
    r = api.Renderer(ctx, W, H)
    r.image_set(vol, env)
    r.next_event_code_set(tf_code)
    r.set_token_cap(max(256 // world, 1))
    r.flush_changes()
    ctx.synchronize()
    xchg_t = None
    if world > 1:
        # compact exchange buffer: one 8-byte cache entry per pixel (all ranks trace the same camera, hence the same voxels)
        xchg_t = torch.as_tensor(_DevArray(r.xchg_device_ptr, r.xchg_bytes // 4, "<i4"), device=f"cuda:{local_rank}")

    def step():
        r.reset_cache()
        r.render_frames(pos, d, seeds, readback=False)
        if world > 1:
            r.xchg_gather()
            with torch.cuda.stream(ext):
                dist.all_reduce(xchg_t)  # int32 view of the packed lanes: sums cannot carry across lanes (cap 256/N)
            r.xchg_scatter()
            r.resolve(readback=False)

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- SDF build time (second half of BASELINE's metric) ----
    sdf_ms = []
    for _ in range(3):
        ctx.synchronize()
        t0 = time.perf_counter()
        s = api.Sdf(ctx, vol, synth.default_tf())
        sdf_ms.append(1e3 * (time.perf_counter() - t0))
        levels = s.levels
        s.close()

    # ---- algorithmic counters of one step (same scene, same seeds) ----
    r.enable_counters(True)
    r.reset_cache()
    r.render_frames(pos, d, seeds, readback=False)
    counters = r.counters(reset=True)
    r.enable_counters(False)

    # ---- warm-up, then the timed region ----
    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = ctx.launches
    r.enable_timing(True)
    r.kernel_times(reset=True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for _ in range(args.steps):
        step()
    e1.record(ext)
    e1.synchronize()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    trace_ms, resolve_ms, nframes = r.kernel_times(reset=True)
    r.enable_timing(False)
    launches = ctx.launches - launches0
    t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    samples_per_step = W * H * SPP * world
    value = samples_per_step * args.steps / ms_max / 1e3  # Msamples/s

    # ---- A/B: the per-frame schedule (mode 1: every frame re-marches its primary rays, as 64 launches of the reference would) ----
    per_frame = None
    if world == 1:
        r.set_trace_mode(1)
        for _ in range(2):
            step()
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record(ext)
        for _ in range(3):
            step()
        p1.record(ext)
        p1.synchronize()
        pms = p0.elapsed_time(p1) / 3
        r.set_trace_mode(2)
        per_frame = {"value": W * H * SPP / pms / 1e3, "unit": "Msamples/s", "ms_per_step": pms,
                     "what": "same step with vr_renderer_set_trace_mode(1): k_trace re-marches the primary ray of every pixel in "
                             "each of the 64 frames (no reuse of the seed-independent part)"}

    # ---- saturated-cache rate: once a voxel holds 256 tokens its samples stop at the token check ----
    saturated = None
    if world == 1:
        r.reset_cache()
        for _ in range(5):
            r.render_frames(pos, d, seeds, readback=False)   # 320 spp without a reset
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(ext)
        for _ in range(3):
            r.render_frames(pos, d, seeds, readback=False)
        s1.record(ext)
        s1.synchronize()
        sms = s0.elapsed_time(s1) / 3
        saturated = {"value": W * H * SPP / sms / 1e3, "unit": "Msamples/s", "ms_per_step": sms,
                     "what": "64 spp after 320 spp without a cache reset (voxels under more than one pixel are at the 256 cap)"}

    # ---- secondary view: close-up camera (about a third of the pixels shaded instead of 8 %) ----
    cpos, cdir = synth.closeup_camera(VOL_N)
    closeup = None
    if world == 1:
        def cstep():
            r.reset_cache()
            r.render_frames(cpos, cdir, seeds, readback=False)
        r.enable_counters(True)
        cstep()
        cc = r.counters(reset=True)
        r.enable_counters(False)
        for _ in range(2):
            cstep()
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(ext)
        for _ in range(3):
            cstep()
        c1.record(ext)
        c1.synchronize()
        cms = c0.elapsed_time(c1) / 3
        cb = alg_bytes_per_sample(cc, trace_only=False)
        closeup = {"value": W * H * SPP / cms / 1e3, "unit": "Msamples/s", "ms_per_step": cms,
                   "camera": "synth.closeup_camera", "shaded_fraction": cc["primary_hits"] / float(cc["samples"]),
                   "steps_per_sample": cc["steps"] / float(cc["samples"]), "alg_bytes_per_sample": cb,
                   "alg_gbs": cb * W * H * SPP / (cms * 1e-3) / 1e9}

    # ---- e2e: the same metric through the C-ABI with host buffers ----
    # headless job: upload the scene from pinned host memory, flush (cache alloc/reset + SDF build), accumulate the step's
    # 64 samples per pixel with vr_render_frames, read the final frame back.  `interactive` = the reference UI's usage
    # instead: 64 x vr_render_frame, every frame read back (renderer.cpp:150).
    e2e_steps = 3

    def e2e_step(interactive):
        v2 = api.Volume(ctx, vol_pin.numpy())          # H2D 256 MiB + fetch_stats
        en2 = api.EnvMap(ctx, env_pin.numpy())         # H2D 8 MiB
        r2 = api.Renderer(ctx, W, H)
        r2.image_set(v2, en2)
        r2.next_event_code_set(tf_code)
        r2.set_token_cap(max(256 // world, 1))
        r2.flush_changes()                             # cache alloc + reset + SDF build
        hf = r2.host_frame()
        if interactive:
            r2.set_primary_reuse(2)                    # camera unchanged between the calls: primary records are kept
            for k in range(SPP):
                r2.render_frame(pos, d, seeds[k], out=hf)  # D2H W*H*4 per frame
        elif world == 1:
            r2.render_frames(pos, d, seeds, out=hf)        # one D2H of the final frame
        else:
            r2.render_frames(pos, d, seeds, readback=False)
            r2.xchg_gather()
            c2 = torch.as_tensor(_DevArray(r2.xchg_device_ptr, r2.xchg_bytes // 4, "<i4"), device=f"cuda:{local_rank}")
            with torch.cuda.stream(ext):
                dist.all_reduce(c2)
            r2.xchg_scatter()
            api._check(api.lib().vr_renderer_resolve(r2.h, hf.ctypes.data_as(api.C.c_void_p)))
        checksum = int(hf[::97, ::89].sum())
        r2.close(); en2.close(); v2.close()
        return checksum

    def time_e2e(interactive):
        e2e_step(interactive)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step(interactive)
        barrier()
        tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=f"cuda:{local_rank}")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return samples_per_step * e2e_steps / float(tt.item()) / 1e6

    # pipelined headless jobs: the volume of job k+1 is uploaded (vr_volume_upload_async, copy stream) while job k builds its
    # SDF, renders its 64 spp and reads its frame back.  Every job still pays its own H2D + D2H inside the timed region;
    # the copy simply no longer waits for the compute stream.
    def time_e2e_pipelined(nsteps):
        r2 = api.Renderer(ctx, W, H)                   # one long-lived renderer serves the stream of jobs
        r2.set_token_cap(max(256 // world, 1))
        hf = r2.host_frame()

        def job(v_ready, start_next):
            en2 = api.EnvMap(ctx, env_pin.numpy())         # first: H2D copies are served in issue order by one copy engine
            nxt = api.Volume(ctx, vol_pin.numpy(), async_upload=True) if start_next else None   # H2D 256 MiB + fetch_stats, async
            r2.image_set(v_ready, en2)
            r2.next_event_code_set(tf_code)
            r2.flush_changes()                             # frame reset + SDF build for the new volume
            if world == 1:
                r2.render_frames(pos, d, seeds, out=hf)
            else:
                r2.render_frames(pos, d, seeds, readback=False)
                r2.xchg_gather()
                c2 = torch.as_tensor(_DevArray(r2.xchg_device_ptr, r2.xchg_bytes // 4, "<i4"), device=f"cuda:{local_rank}")
                with torch.cuda.stream(ext):
                    dist.all_reduce(c2)
                r2.xchg_scatter()
                api._check(api.lib().vr_renderer_resolve(r2.h, hf.ctypes.data_as(api.C.c_void_p)))
            chk = int(hf[::97, ::89].sum())
            ctx.synchronize()
            en2.close(); v_ready.close()
            return nxt, chk
        v = api.Volume(ctx, vol_pin.numpy(), async_upload=True)
        for _ in range(3):                        # warm-up jobs (pool growth, the second SDF array); leave the next volume in flight
            v, _ = job(v, True)
        barrier()
        t0 = time.perf_counter()
        for k in range(nsteps):
            v, _ = job(v, True)                   # nsteps uploads are issued inside the timed region
        barrier()
        tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=f"cuda:{local_rank}")
        v.close(); r2.close()
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return samples_per_step * nsteps / float(tt.item()) / 1e6

    if os.environ.get("VR_E2E_BREAKDOWN"):
        for it in range(2):
            marks = []
            def m(name):
                ctx.synchronize(); marks.append((name, time.perf_counter()))
            m("start")
            v2 = api.Volume(ctx, vol_pin.numpy()); m("volume_upload+stats")
            en2 = api.EnvMap(ctx, env_pin.numpy()); m("env")
            r2 = api.Renderer(ctx, W, H); m("renderer_create")
            r2.image_set(v2, en2); r2.next_event_code_set(tf_code); r2.set_token_cap(max(256 // world, 1))
            r2.flush_changes(); m("flush")
            hf = r2.host_frame()
            r2.render_frames(pos, d, seeds, readback=False); m("render_frames")
            if world > 1:
                r2.xchg_gather(); m("gather")
                c2 = torch.as_tensor(_DevArray(r2.xchg_device_ptr, r2.xchg_bytes // 4, "<i4"), device=f"cuda:{local_rank}")
                m("as_tensor")
                with torch.cuda.stream(ext):
                    dist.all_reduce(c2)
                m("all_reduce")
                r2.xchg_scatter(); m("scatter")
            api._check(api.lib().vr_renderer_resolve(r2.h, hf.ctypes.data_as(api.C.c_void_p))); m("resolve+readback")
            r2.close(); m("renderer_close"); en2.close(); v2.close(); m("scene_close")
            log(f"[rank {rank}] e2e breakdown: " + ", ".join(f"{n} {1e3*(t-marks[i][1]):.2f}ms" for i, (n, t) in enumerate(marks[1:])))
    e2e_sequential = time_e2e(False)
    e2e_value = time_e2e_pipelined(2 * e2e_steps)
    e2e_interactive = time_e2e(True) if world == 1 else None

    # ---- CPU baseline (rank 0, N=1 only): the oracle on a bounded sample ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)))
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as o
        import ref_lib as R
        sdf_np = r.sdf_download()  # bit-identical to the oracle's (tests/test_parity_gpu.py); saves minutes of CPU SDF build
        if R.available():
            kind, cores = "reference", R.num_threads()
            ref = R.Renderer(vol_np, env_np, synth.default_tf(), W, H, sdf_np)
            frame_fn = lambda s: ref.render_frame(pos, d, s, threads=0)  # noqa: E731
            what = "the reference's own `render` kernel compiled for the host (oracle/_ref), OpenMP over rows"
        else:
            kind, cores = "port", o.num_threads()
            ref = o.Renderer(vol_np, env_np, synth.default_tf(), W, H, sdf=sdf_np)
            frame_fn = lambda s: ref.render_frame(pos, d, s)  # noqa: E731
            what = "oracle/oracle.cpp, OpenMP over rows"
        frame_fn(seeds[0])
        t0 = time.perf_counter()
        nfr = 0
        while nfr < 16 and (time.perf_counter() - t0) < 12.0:
            frame_fn(seeds[1 + nfr])
            nfr += 1
        dt = time.perf_counter() - t0
        cpu = {"value": W * H * nfr / dt / 1e6, "unit": "Msamples/s", "cores": cores, "kind": kind,
               "sample": f"{nfr} x 1 spp frame of the same 1920x1080/512^3 scene ({what}; SDF taken from the GPU build, which "
                         f"the parity tests pin bit-exact)"}

    # ---- the reference's own OpenCL kernels on THIS GPU (rank 0, N=1 only; part of the baseline leg, reported beside cpu_baseline):
    # when the driver ships an OpenCL runtime, oracle/_ref/libref_ocl.so JIT-compiles the reference's unmodified kernels with
    # the reference's options and runs the reference's host sequences (SDF build loop with its blocking counter transfers,
    # render_frame with its blocking frame pull) on the same B200, same scene, same seeds, from a reset cache.
    ref_gpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import ref_ocl_lib as RO
            if RO.available():
                RO.set_nearest(False)  # as shipped
                sc = RO.Scene(vol_np, env_np, tf_code, W, H)
                sc.render(pos, d, seeds[:2], readback=False)
                sc.reset()
                _, ms_pull = sc.render(pos, d, seeds[:SPP], pull_every_frame=True)
                sc.reset()
                _, ms_kernel = sc.render(pos, d, seeds[:SPP], pull_every_frame=False, readback=False)
                ref_gpu = {"device": RO.info(), "what": "the reference's unmodified OpenCL kernels (as shipped, -cl-mad-enable -cl-std=CL1.2) "
                                                        "and host sequences on this GPU: 64 x render_frame of the bench scene from a reset "
                                                        "cache (renderer.cpp:131-158); SDF build loop of signed_distance_field.cpp:7-35",
                           "value": W * H * SPP / ms_pull / 1e3, "unit": "Msamples/s", "ms_per_frame": ms_pull / SPP,
                           "kernel_only": {"value": W * H * SPP / ms_kernel / 1e3, "unit": "Msamples/s", "ms_per_frame": ms_kernel / SPP,
                                           "what": "same launches without the per-frame blocking frame pull"},
                           "sdf_build_ms": sc.sdf_ms, "sdf_iterations": sc.sdf_iterations, "jit_ms": sc.sdf_jit_ms + sc.render_jit_ms,
                           "note": "every sampler of the reference requests CLK_FILTER_LINEAR on integer images (undefined in OpenCL 1.2); "
                                   "NVIDIA's texture units interpolate, so as shipped the rays see other values than under the "
                                   "spec-defined NEAREST reading this repository implements (DESIGN.md 2.1); same amount of work"}
                sc.close()
            else:
                ref_gpu = {"unavailable": RO.error()}
        except Exception as e:  # the baseline must never take the bench line down
            ref_gpu = {"unavailable": f"{type(e).__name__}: {e}"}

    if rank == 0:
        hbm, peak_src = peaks()
        b_trace = alg_bytes_per_sample(counters, trace_only=True)
        n_launches = max(nframes // SPP, 1)             # one trace launch (k_trace + k_trace_pt) covers the 64 frames of a step
        trace_launch_ms = trace_ms / n_launches
        achieved = b_trace * W * H * SPP / (trace_launch_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic_k_trace.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        S = float(counters["samples"])
        out = {
            "metric": "path_msamples_per_s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "parallelism": (f"spp-split x{world}: 64 spp per rank, per-rank token cap 256/{world} (the reference's cap of 256 "
                                       f"per voxel holds for the summed cache); voxels under several pixels reach the per-rank cap "
                                       f"inside the step and their later samples stop at the token check, as in the reference's "
                                       f"steady state, so per-rank work shrinks with N") if world > 1 else "single GPU",
                       "l2": "inputs larger than L2 (cache 1 GiB + volume 256 MiB + SDF 128 MiB), no flush"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Msamples/s",
                    "h2d_bytes_per_step": int(vol_np.nbytes + env_np.nbytes), "d2h_bytes_per_step": int(W * H * 4),
                    "steps": 2 * e2e_steps,
                    "what": "per step (one headless job): volume + env map uploaded from pinned host memory, vr_renderer_flush "
                            "(cache alloc/reset + SDF build), vr_render_frames(64 seeds), final frame read back to the host; "
                            "jobs are pipelined: vr_volume_upload_async copies job k+1's volume on the copy stream while job k "
                            "computes (every job's H2D and D2H are inside the timed region); one renderer object serves all "
                            "jobs",
                    "sequential": {"value": e2e_sequential, "unit": "Msamples/s", "steps": e2e_steps,
                                   "what": "the same jobs one after the other with the blocking vr_volume_upload"},
                    "interactive": {"value": e2e_interactive, "unit": "Msamples/s", "d2h_bytes_per_step": int(W * H * 4 * SPP),
                                    "what": "same, but 64 x vr_render_frame with every frame read back (the reference UI's "
                                            "usage, renderer.cpp:131-158), vr_renderer_set_primary_reuse(2)"}},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "trace phase = k_primary (seed-independent part of the 64 samples of a pixel: ray, "
                                                   "box cut, primary march, env colour / hit voxel + normal; once per pixel and "
                                                   "step) + k_trace_pt (token admission + secondary paths per pixel and frame), "
                                                   "one launch pair per 64-frame step",
                         "achieved": achieved, "peak": hbm, "unit": "GB/s",
                         "frac": achieved / hbm, "traffic": traffic, "peak_source": peak_src,
                         "alg_bytes_per_sample": b_trace, "samples_per_launch": W * H * SPP,
                         "launch_ms": trace_launch_ms, "resolve_launch_ms": resolve_ms / n_launches,
                         "note": "achieved = the reference algorithm's bytes (SURVEY 8d: 15 B per march step, ...) for the 64 samples "
                                 "per pixel over the measured trace time; the kernels move fewer bytes: one SDF byte per step "
                                 "instead of 15 and the primary segment once per pixel instead of 64 times (DESIGN.md 4.1). "
                                 "Scattered 1-byte gathers: issue-bound, not HBM-bound (profiles/)",
                         "trace_share_of_step": trace_ms / ms if ms > 0 else None,
                         "per_sample": {"steps": counters["steps"] / S, "normals": counters["normals"] / S,
                                        "env": counters["env"] / S, "primary_hits": counters["primary_hits"] / S,
                                        "admitted": counters["admitted"] / S}},
            "cpu_baseline": cpu,
            "reference_on_gpu": ref_gpu,
            "closeup": closeup,
            "per_frame_schedule": per_frame,
            "saturated_cache": saturated,
            "sdf_build_ms": {"value": float(np.median(sdf_ms)), "levels": levels, "volume": f"{VOL_N}^3",
                             "note": "vr_sdf_build wall time incl. allocation, excl. upload (app/sdf_benchmark.cpp:15-20)"},
        }
        print(json.dumps(out), flush=True)
    r.close(); env.close(); vol.close(); ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        if args.warmup < 3:
            log(f"timing rules want at least 3 warm-up steps: running 3 instead of {args.warmup}")
            args.warmup = 3
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
