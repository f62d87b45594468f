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
C3_SPP = 1024
ENV_W, ENV_H = 2048, 1024
WORKLOAD = (f"{VOL_N}^3 short synthetic CT (synth_ct), {W}x{H}, {SPP} spp from a reset voxel cache, default TF rect(500,1200), "
            f"synthetic env {ENV_W}x{ENV_H}, camera (-400,400,-400) look (0.9,6.183)")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# The contract is ONE JSON line on stdout.  Libraries print there too (NCCL announces its version on stdout when NCCL_DEBUG asks for
# it), so file descriptor 1 is pointed at stderr for the whole run and the result line goes to the saved descriptor.
_REAL_STDOUT = None


def capture_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe).  The timed region of the default run is
    ~10 ms, shorter than one `nvidia-smi -lms` period, so the clocks are polled through NVML every ~1 ms — in a separate PROCESS,
    started before the warm-up: a polling thread inside the bench process competes with the launching thread for the interpreter
    lock and the driver, which a single-GPU step hides behind its queue but which cost a multi-GPU step 25 % (every rank waits
    for rank 0 in the exchange: 2.51 ms per step at 8 GPUs with the thread, 2.01 without).  `begin()` / `end()` bracket the timed
    region; only samples taken inside it are reported.  `nvidia-smi --query-gpu` is the fallback when pynvml is missing."""
    POLLER = r"""
import sys, time
import pynvml as n
n.nvmlInit()
bus = sys.argv[1]
h = None
for i in range(n.nvmlDeviceGetCount()):
    hh = n.nvmlDeviceGetHandleByIndex(i)
    b = n.nvmlDeviceGetPciInfo(hh).busId
    b = b.decode() if isinstance(b, bytes) else b
    if bus and bus.lower() in b.lower():
        h = hh
        break
if h is None:
    h = n.nvmlDeviceGetHandleByIndex(int(sys.argv[2]))
try:
    mx = n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)
except Exception:
    mx = -1
print("ready", mx, flush=True)
while True:
    try:
        sm = n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        print(repr(time.time()), sm, int(r), flush=True)
    except Exception:
        pass
    time.sleep(0.001)
"""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []
        self.mx = None
        self.t0 = self.t1 = None
        self.mode = None

    def start(self):
        """launch the poller and wait until it samples (call before the warm-up)"""
        try:
            import pynvml  # noqa: F401  (only to know the subprocess can import it)
            import torch
            bus = ""
            try:
                pr = torch.cuda.get_device_properties(self.index)
                bus = ":%02x:%02x." % (pr.pci_bus_id, pr.pci_device_id) if hasattr(pr, "pci_bus_id") else ""
            except Exception:
                pass
            self.proc = subprocess.Popen([sys.executable, "-c", self.POLLER, bus, str(self.index)], stdout=subprocess.PIPE, text=True)
            first = self.proc.stdout.readline().split()
            if not first or first[0] != "ready":
                raise RuntimeError("poller did not start")
            self.mx = float(first[1]) if float(first[1]) > 0 else None
            self.mode = "nvml"
        except Exception as e:
            log("NVML poller unavailable, falling back to nvidia-smi:", e)
            try:
                Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
                     "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
                self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={Q}", "--format=csv,noheader,nounits", "-i", str(self.index),
                                              "-lms", "100"], stdout=subprocess.PIPE, text=True)
                self.mode = "smi"
            except Exception as e2:  # nvidia-smi missing: report it, do not fail the bench
                log("clock sampler unavailable:", e2)
                self.proc = None
                return
        self.t = threading.Thread(target=self._read, daemon=True)  # drains the pipe; idle while nothing arrives
        self.t.start()

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def begin(self):
        self.t0 = time.time()

    def end(self):
        self.t1 = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampler unavailable"]}
        time.sleep(0.15 if self.mode == "smi" else 0.01)
        self.proc.terminate()
        names = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        if self.mode == "nvml":
            sm, bits, n_in = [], 0, 0
            every = []
            for ln in list(self.lines):
                f = ln.split()
                if len(f) != 3:
                    continue
                try:
                    t, c, r = float(f[0]), float(f[1]), int(f[2])
                except ValueError:
                    continue
                every.append(c)
                if self.t0 is not None and self.t1 is not None and self.t0 <= t <= self.t1:
                    sm.append(c); bits |= r; n_in += 1
            if not sm and every:  # a region shorter than one poll: the samples next to it
                sm = every[-3:]
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.mx, "reasons": sorted(k for k, b in names.items() if bits & b),
                    "samples": n_in, "source": "NVML polled every ~1 ms by a separate process; samples inside the timed region"}
        rows = [[c.strip() for c in ln.split(",")] for ln in self.lines]
        sm = [float(r[1]) for r in rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
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
    emit({
        "impl": "reference", "metric": "path_msamples_per_s", "value": value, "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": cores, "kind": kind, "sample": sample,
                         "sdf_build_ms_oracle": 1e3 * sdf_s},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


# ------------------------------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from cl_volume_renderer_b200 import api, synth

    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    ctx = api.Context(local_rank)
    if world > 1:
        # torch.distributed is plumbing only (gloo: hands the NCCL id to the ranks); every data-path collective goes through the
        # library's own NCCL communicator behind the C-ABI (vr_comm_init / vr_cache_allreduce / vr_frame_allgather / ...)
        dist.init_process_group("gloo")
        box = [api.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctx.comm_init(rank, world, box[0])
    ext = torch.cuda.ExternalStream(ctx.stream, device=local_rank)

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            ctx.comm_barrier()

    def rank_max(x):
        return float(ctx.comm_allreduce(np.array([x], dtype=np.float64), "max")[0]) if world > 1 else float(x)

    def rank_min_int(x):
        return int(ctx.comm_allreduce(np.array([x], dtype=np.int32), "min")[0]) if world > 1 else int(x)

    def timed(fn, steps, warmup):
        """CUDA events on the context's stream around `steps` calls, max over ranks; -> ms per call"""
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        for _ in range(steps):
            fn()
        e1.record(ext)
        e1.synchronize()
        barrier()
        return rank_max(e0.elapsed_time(e1)) / steps

    vol_np, env_np, pos, d = make_scene()
    vol_pin = torch.empty(vol_np.shape, dtype=torch.int16, pin_memory=True)   # the e2e legs upload from pinned host memory
    vol_pin.numpy()[...] = vol_np
    env_pin = torch.empty(env_np.shape, dtype=torch.uint8, pin_memory=True)
    env_pin.numpy()[...] = env_np
    tf_code = api.tf_format(synth.default_tf())
    all_seeds = synth.glibc_rand(max(SPP * world, C3_SPP))
    seeds = all_seeds[SPP * rank: SPP * (rank + 1)]
    cpos, cdir = synth.closeup_camera(VOL_N)

    vol = api.Volume(ctx, vol_pin.numpy())
    env = api.EnvMap(ctx, env_pin.numpy())
    hbm, peak_src = peaks()

    # ------------------------------------------------------------------------------------------------------------------
    def render_set(sampling, steps, warmup, full):
        """One line-set of the path tracer under a sampling mode: timed steps (value), counters + kernel times (roofline), and —
        when `full` — the close-up view, the per-frame schedule and the saturated-cache rate."""
        r = api.Renderer(ctx, W, H)
        r.set_sampling(sampling)
        r.image_set(vol, env)
        r.next_event_code_set(tf_code)
        r.flush_changes()
        ctx.synchronize()

        def step():
            r.reset_cache()
            r.render_frames(pos, d, seeds, readback=False)
            if world > 1:
                r.cache_allreduce()   # every rank accumulated a full 64-spp job (cap 256): wide lanes, resolve from the global sums

        r.enable_counters(True)
        r.reset_cache()
        r.render_frames(pos, d, seeds, readback=False)
        counters = r.counters(reset=True)
        r.enable_counters(False)
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()          # its own process, running before the warm-up
        for _ in range(warmup):
            step()
        barrier()
        launches0 = ctx.launches
        r.enable_timing(True)
        r.kernel_times(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler.begin()
        e0.record(ext)
        for _ in range(steps):
            step()
        e1.record(ext)
        e1.synchronize()
        sampler.end()
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        ms_local = e0.elapsed_time(e1)
        trace_ms, resolve_ms, nframes = r.kernel_times(reset=True)
        r.enable_timing(False)
        launches = ctx.launches - launches0
        ms = rank_max(ms_local)
        out = {"value": W * H * SPP * world * steps / ms / 1e3, "ms_per_step": ms / steps, "clocks": clocks, "gpu_launches": int(launches)}
        b_trace = alg_bytes_per_sample(counters, trace_only=True)
        n_launch = max(nframes // SPP, 1)
        launch_ms = trace_ms / n_launch
        achieved = b_trace * W * H * SPP / (launch_ms * 1e-3) / 1e9
        S = float(counters["samples"])
        out["roofline"] = {
            "bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "peak_source": peak_src,
            "alg_bytes_per_sample": b_trace, "samples_per_launch": W * H * SPP, "launch_ms": launch_ms,
            "resolve_launch_ms": resolve_ms / n_launch, "trace_share_of_step": trace_ms / ms_local if ms_local > 0 else None,
            "per_sample": {k: counters[k] / S for k in ("steps", "normals", "env", "primary_hits", "admitted")}}
        if full:
            r.set_trace_mode(1)
            pms = timed(step, 3, 2)
            r.set_trace_mode(2)
            out["per_frame_schedule"] = {
                "value": W * H * SPP / pms / 1e3, "unit": "Msamples/s", "ms_per_step": pms,
                "what": "same step with vr_renderer_set_trace_mode(1): k_trace re-marches the primary ray of every pixel in each of the "
                        "64 frames (no reuse of the seed-independent part), persistent warps run the secondary paths"}
            r.reset_cache()
            for _ in range(5):
                r.render_frames(pos, d, seeds, readback=False)   # 320 spp without a reset
            sms = timed(lambda: r.render_frames(pos, d, seeds, readback=False), 3, 0)
            out["saturated_cache"] = {"value": W * H * SPP / sms / 1e3, "unit": "Msamples/s", "ms_per_step": sms,
                                      "what": "64 spp after 320 spp without a cache reset (voxels under more than one pixel are at the 256 cap)"}

            def cstep():
                r.reset_cache()
                r.render_frames(cpos, cdir, seeds, readback=False)
            r.enable_counters(True)
            cstep()
            cc = r.counters(reset=True)
            r.enable_counters(False)
            cms = timed(cstep, 3, 2)
            cb = alg_bytes_per_sample(cc, trace_only=False)
            out["closeup"] = {"value": W * H * SPP / cms / 1e3, "unit": "Msamples/s", "ms_per_step": cms, "camera": "synth.closeup_camera",
                              "shaded_fraction": cc["primary_hits"] / float(cc["samples"]),
                              "steps_per_sample": cc["steps"] / float(cc["samples"]), "alg_bytes_per_sample": cb,
                              "alg_gbs": cb * W * H * SPP / (cms * 1e-3) / 1e9}
        return out, r

    # ---- e2e: the same metric through the C-ABI with host buffers --------------------------------------------------------
    def e2e_set(sampling, nsteps):
        """per step one headless job: volume + env map from pinned host memory, flush (cache reset + SDF build [+ textures and step
        field]), 64 spp, final frame read back to the host"""
        res = {}
        r2 = api.Renderer(ctx, W, H)                   # one long-lived renderer serves the stream of jobs
        r2.set_sampling(sampling)
        hf = r2.host_frame()
        if world == 1:
            def job(v_ready, start_next):
                en2 = api.EnvMap(ctx, env_pin.numpy())         # first: H2D copies are served in issue order by one copy engine
                nxt = api.Volume(ctx, vol_pin.numpy(), async_upload=True) if start_next else None   # H2D 256 MiB + fetch_stats, async
                r2.image_set(v_ready, en2)
                r2.next_event_code_set(tf_code)
                r2.flush_changes()
                r2.render_frames(pos, d, seeds, out=hf)
                chk = int(hf[::97, ::89].sum())
                ctx.synchronize()
                en2.close(); v_ready.close()
                return nxt, chk
            v = api.Volume(ctx, vol_pin.numpy(), async_upload=True)
            for _ in range(3):
                v, _ = job(v, True)
            barrier()
            t0 = time.perf_counter()
            for _ in range(nsteps):
                v, _ = job(v, True)
            barrier()
            dt = time.perf_counter() - t0
            v.close()
            res["value"] = W * H * SPP * nsteps / dt / 1e6
            res["h2d_bytes_per_step"] = int(vol_np.nbytes + env_np.nbytes)
            res["what"] = ("per step (one headless job): volume + env map uploaded from pinned host memory, vr_renderer_flush (cache "
                           "reset + SDF build), vr_render_frames(64 seeds), final frame read back; jobs are pipelined: "
                           "vr_volume_upload_async copies job k+1's volume on the copy stream while job k computes (every job's H2D and "
                           "D2H are inside the timed region)")
        else:
            z0, z1 = ctx.comm_slab(VOL_N)
            own = vol_pin.numpy()[z0:z1]
            r2.set_sharded_build(True)

            def job(v_ready, start_next):
                en2 = api.EnvMap(ctx, env_pin.numpy())
                # H2D of this rank's planes of job k+1, the rest over NVLink, fetch_stats: all on the copy stream / communicator
                nxt = api.Volume(ctx, own, sharded_dims=(VOL_N, VOL_N, VOL_N), async_upload=True) if start_next else None
                r2.image_set(v_ready, en2)
                r2.next_event_code_set(tf_code)
                r2.flush_changes()                                              # SDF build (z-slab sharded above 512^3)
                r2.render_frames(pos, d, seeds, readback=False)
                r2.cache_allreduce(readback=True, out=hf)
                chk = int(hf[::97, ::89].sum())
                en2.close(); v_ready.close()
                return nxt, chk
            v = api.Volume(ctx, own, sharded_dims=(VOL_N, VOL_N, VOL_N), async_upload=True)
            for _ in range(2):
                v, _ = job(v, True)
            barrier()
            t0 = time.perf_counter()
            for _ in range(nsteps):
                v, _ = job(v, True)
            barrier()
            dt = rank_max(time.perf_counter() - t0)
            v.close()
            res["value"] = W * H * SPP * world * nsteps / dt / 1e6
            res["h2d_bytes_per_step"] = int(own.nbytes + env_np.nbytes)
            res["what"] = ("per step (one headless job on N ranks): every rank uploads ITS z-slab of the volume from pinned host memory "
                           "(vr_volume_upload_sharded_async: the other planes arrive over NVLink; copy, gather and fetch_stats on the copy "
                           "stream and its own communicator while the previous job computes) and the env map, vr_renderer_flush builds the "
                           "SDF (on every rank: at 512^3 a level is one wave of thread blocks and z-slabs would only add halo swaps and a "
                           "gather — the library shards the build above that size, as the c4 / c5 legs do), every rank traces its own 64 seeds, "
                           "vr_cache_allreduce sums the touched cache entries and every rank reads the resolved frame back; every job's "
                           "H2D and D2H are inside the timed region")
        res.update({"unit": "Msamples/s", "d2h_bytes_per_step": int(W * H * 4), "steps": nsteps})
        r2.close()
        return res

    def e2e_sequential_and_interactive(sampling):
        def one(interactive):
            v2 = api.Volume(ctx, vol_pin.numpy())
            en2 = api.EnvMap(ctx, env_pin.numpy())
            r2 = api.Renderer(ctx, W, H)
            r2.set_sampling(sampling)
            r2.image_set(v2, en2)
            r2.next_event_code_set(tf_code)
            r2.flush_changes()
            hf = r2.host_frame()
            if interactive:
                r2.set_primary_reuse(2)                    # camera unchanged between the calls: primary records are kept
                for k in range(SPP):
                    r2.render_frame(pos, d, seeds[k], out=hf)  # the frame is pulled every call (renderer.cpp:150)
            else:
                r2.render_frames(pos, d, seeds, out=hf)
            chk = int(hf[::97, ::89].sum())
            r2.close(); en2.close(); v2.close()
            return chk
        out = {}
        for name, inter in (("sequential", False), ("interactive", True)):
            one(inter)
            barrier()
            t0 = time.perf_counter()
            for _ in range(3):
                one(inter)
            barrier()
            out[name] = W * H * SPP * 3 / (time.perf_counter() - t0) / 1e6
        return out

    def interactive_loop(rr):
        """64 x vr_render_frame of the resident scene from a reset cache, every frame pulled into the renderer-owned host frame —
        exactly what `reference_on_gpu.value` times for the reference (its render_frame loop, renderer.cpp:131-158; no upload, no
        SDF build inside the timed region)"""
        hf = rr.host_frame()
        rr.set_primary_reuse(2)
        best = 1e9
        for rep in range(3):
            rr.reset_cache()
            rr.render_frame(pos, d, all_seeds[SPP], out=hf)    # the camera's primary records + the first, complete pull
            ctx.synchronize()
            t0 = time.perf_counter()
            for k in range(SPP):
                rr.render_frame(pos, d, seeds[k], out=hf)
            ctx.synchronize()
            best = min(best, time.perf_counter() - t0)
        rr.set_primary_reuse(1)
        return {"value": W * H * SPP / best / 1e6, "unit": "Msamples/s", "ms_per_frame": 1e3 * best / SPP,
                "what": "64 x vr_render_frame(host frame) on the resident scene, camera at rest, vr_renderer_set_primary_reuse(2): trace of "
                        "the frame's secondary paths, resolve, incremental pull of the shaded pixels' bounding box; wall clock, best of 3"}

    # ---- SDF build time (second half of BASELINE's metric) ----
    sdf_ms = []
    for _ in range(3):
        ctx.synchronize()
        t0 = time.perf_counter()
        s = api.Sdf(ctx, vol, synth.default_tf())
        sdf_ms.append(1e3 * (time.perf_counter() - t0))
        levels = s.levels
        s.close()

    main, r = render_set(api.VR_SAMPLING_NEAREST, args.steps, args.warmup, full=(world == 1))
    lin, rl = render_set(api.VR_SAMPLING_HW_LINEAR, max(args.steps, 3), 3, full=(world == 1))
    if world == 1:
        main["interactive_loop"] = interactive_loop(r)
        lin["interactive_loop"] = interactive_loop(rl)
    rl.close()

    e2e_main = e2e_set(api.VR_SAMPLING_NEAREST, 6)
    e2e_lin = e2e_set(api.VR_SAMPLING_HW_LINEAR, 6)
    if world == 1:
        si = e2e_sequential_and_interactive(api.VR_SAMPLING_NEAREST)
        e2e_main["sequential"] = {"value": si["sequential"], "unit": "Msamples/s", "what": "the same jobs one after the other with the blocking vr_volume_upload"}
        e2e_main["interactive"] = {"value": si["interactive"], "unit": "Msamples/s", "d2h_bytes_per_step": int(W * H * 4 * SPP),
                                   "what": "same, but 64 x vr_render_frame with every frame read back (the reference UI's usage, "
                                           "renderer.cpp:131-158), vr_renderer_set_primary_reuse(2)"}
        sil = e2e_sequential_and_interactive(api.VR_SAMPLING_HW_LINEAR)
        e2e_lin["interactive"] = {"value": sil["interactive"], "unit": "Msamples/s", "d2h_bytes_per_step": int(W * H * 4 * SPP)}

    # ---- BASELINE config 3: 1024 spp of the scene, strong-split by spp over the ranks ------------------------------------------------
    def c3_strong():
        mine = all_seeds[rank:C3_SPP:world]                 # 1024 / N frames on this rank
        r.set_token_cap(max(256 // world, 1))               # the summed cache obeys the reference's cap of 256: packed lanes cannot carry

        def job():
            r.reset_cache()
            for k in range(0, len(mine), 64):
                r.render_frames(pos, d, mine[k:k + 64], readback=False)
            r.cache_allreduce()
        ms = timed(job, 2, 1)
        r.set_token_cap(256)
        return {"value": W * H * C3_SPP / ms / 1e3, "unit": "Msamples/s", "ms_per_job": ms, "spp_total": C3_SPP, "spp_per_rank": len(mine),
                "token_cap_per_rank": max(256 // world, 1), "scaling": "strong",
                "what": "BASELINE config 3: ONE 1024-spp job of the bench scene from a reset cache; rank r traces seeds r, r+N, ... with a "
                        "token cap of 256/N, then vr_cache_allreduce (compact: one entry per shaded pixel) + resolve on every rank; "
                        "time of the whole job, max over ranks.  A voxel under p pixels is offered 1024p/N samples per rank and admits "
                        "min(256, 1024p)/N of them: the total work is the same at every N"}

    # ---- BASELINE config 4: 1024^3 volume, 3840x2160, 256 spp, image-tile split ---------------------------------------------------
    big = {}

    def big_volume():
        """1024^3 int16 on the device: the 512^3 synthetic CT upsampled x2 (trilinear) by torch — generating it with numpy on the host
        costs minutes; -> (api.Volume, torch tensor that owns nothing the library needs after the copy)"""
        if "vol" not in big:
            import torch.nn.functional as F
            t = torch.from_numpy(vol_np).to(dev).float()[None, None]
            up = F.interpolate(t, scale_factor=2, mode="trilinear", align_corners=False)[0, 0].round_().to(torch.int16).contiguous()
            del t
            torch.cuda.synchronize()
            big["vol"] = api.Volume.from_device(ctx, up.data_ptr(), 2 * VOL_N, 2 * VOL_N, 2 * VOL_N)
            del up
            torch.cuda.empty_cache()
        return big["vol"]

    def c4_tiles():
        n4, W4, H4, SPP4, BLOCK = 2 * VOL_N, 3840, 2160, 256, 24
        v4 = big_volume()
        r4 = api.Renderer(ctx, W4, H4)
        r4.image_set(v4, env)
        r4.next_event_code_set(tf_code)
        r4.set_sharded_build(True)
        barrier()
        t0 = time.perf_counter()
        r4.flush_changes()
        barrier()
        flush_ms = rank_max(1e3 * (time.perf_counter() - t0))
        p4, d4 = synth.closeup_camera(n4)
        s4 = all_seeds[:SPP4]
        r4.set_row_blocks(BLOCK, rank, world)

        def job():
            r4.reset_cache()
            for k in range(0, SPP4, 64):
                r4.render_frames(p4, d4, s4[k:k + 64], readback=False)
            r4.frame_allgather()
        ms = timed(job, 2, 1)
        frame = r4.frame_allgather(readback=True)
        res = {"value": W4 * H4 * SPP4 / ms / 1e3, "unit": "Msamples/s", "ms_per_job": ms, "scaling": "strong", "volume": f"{n4}^3",
               "frame": f"{W4}x{H4}", "spp": SPP4, "rows_per_block": BLOCK, "flush_ms_incl_sharded_sdf_build": flush_ms,
               "shaded_fraction": float((frame[..., 3] == 1).mean()), "frame_checksum": int(frame.astype(np.uint64).sum()),
               "memory_per_rank_gib": round((2 + 1 + 2 + 8) * n4 ** 3 / 2 ** 30, 1),
               "what": "BASELINE config 4: 1024^3 volume (the 512^3 synthetic CT upsampled x2 on the device), close-up camera, 256 spp; "
                       "rows dealt out in blocks of 24 (vr_renderer_set_row_blocks), every rank traces its blocks into its own cache, "
                       "vr_frame_allgather completes the frame on every rank; time of the whole job incl. the all-gather, max over ranks"}
        r4.close()
        return res

    # ---- BASELINE config 5: SDF build + histogram + volume filter sweep, z-slab sharded -----------------------------------------
    def c5_sweep():
        rows = []
        tf = synth.default_tf()
        for n5 in (128, 256, 512, 1024):
            if n5 == VOL_N:
                v5, own = vol, False
            elif n5 == 2 * VOL_N:
                v5, own = big_volume(), False
            else:
                v5, own = api.Volume(ctx, synth.synth_ct(n5)), True
            st = v5.stats()
            rng = [float(x) for x in st]
            row = {"n": n5}
            # SDF
            ts, chk = [], None
            for rep in range(3):
                barrier()
                t0 = time.perf_counter()
                s5 = api.Sdf(ctx, v5, tf, sharded=True)
                ts.append(rank_max(1e3 * (time.perf_counter() - t0)))
                chk = s5.checksum()
                s5.close()
            row["sdf_build_ms"] = min(ts)
            ts = []
            for rep in range(3):
                barrier()
                t0 = time.perf_counter()
                s5 = api.Sdf(ctx, v5, tf, sharded="slab")
                ts.append(rank_max(1e3 * (time.perf_counter() - t0)))
                s5.close()
            row["sdf_build_ms_without_gather"] = min(ts)
            if world > 1:
                s1 = api.Sdf(ctx, v5, tf)      # the single-GPU build of the same volume on this rank
                row["sdf_identical_to_single_gpu_build"] = bool(rank_min_int(int(s1.checksum() == chk)))
                s1.close()
            # histogram (500x500 as the UI asks for, ui.cpp:148)
            ts = []
            for rep in range(3):
                barrier()
                t0 = time.perf_counter()
                bins = v5.histogram_sharded(500, 500, rng)
                ts.append(rank_max(1e3 * (time.perf_counter() - t0)))
            row["histogram_ms_incl_1MB_readback"] = min(ts)
            if world > 1:
                row["histogram_identical_to_single_gpu"] = bool(rank_min_int(int(np.array_equal(bins, v5.histogram(500, 500, rng)))))
            # bilateral filter on a copy (it replaces the volume)
            ts = []
            want = None
            if world > 1:
                c1 = api.Volume.from_device(ctx, _volume_device_ptr(v5), n5, n5, n5)
                c1.filter()
                want = c1.checksum()
                c1.close()
            for rep in range(2):
                cpy = api.Volume.from_device(ctx, _volume_device_ptr(v5), n5, n5, n5)
                barrier()
                t0 = time.perf_counter()
                cpy.filter_sharded()
                ctx.synchronize()
                ts.append(rank_max(1e3 * (time.perf_counter() - t0)))
                got = cpy.checksum()
                cpy.close()
            row["bilateral_ms"] = min(ts)
            if world > 1:
                row["bilateral_identical_to_single_gpu"] = bool(rank_min_int(int(got == want)))
            rows.append(row)
            if own:
                v5.close()
        return {"rows": rows, "what": "BASELINE config 5: vr_sdf_build_sharded (z-slabs + 16 halo planes, bit-volume halo swaps every 14 "
                                      "levels with ncclSend/ncclRecv, gather of the field so that every rank can render; "
                                      "`sdf_build_ms_without_gather` = vr_sdf_build_slab_only, the build itself), vr_histogram_sharded (slab counts + all-reduce), "
                                      "vr_volume_filter_sharded (slab + 2 halo planes, gather) — wall time per call incl. its "
                                      "synchronisation, min of 2-3 repetitions after the first, max over ranks; results compared with the "
                                      "single-GPU calls on every rank through device-side checksums"}

    def _volume_device_ptr(v):
        p = api.lib().vr_volume_device_ptr(v.h)
        assert p
        return p

    c3 = c3_strong()
    c4 = c4_tiles() if not args.skip_big else None
    c5 = c5_sweep() if not args.skip_big else None
    if "vol" in big:
        big["vol"].close()

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
                                   "NVIDIA's texture units interpolate: this run computes what vr_renderer_set_sampling("
                                   "VR_SAMPLING_HW_LINEAR) computes (the `hw_linear` line-set), not the NEAREST reading of the top-level line"}
                sc.close()
            else:
                ref_gpu = {"unavailable": RO.error()}
        except Exception as e:  # the baseline must never take the bench line down
            ref_gpu = {"unavailable": f"{type(e).__name__}: {e}"}

    if rank == 0:
        def prof(name):
            pth = os.path.join(ROOT, "profiles", name)
            return json.load(open(pth)) if os.path.exists(pth) else None
        traffic = (prof("traffic_k_trace.json") or {}).get("dram_bytes_per_launch")
        main["roofline"].update({
            "kernel": "trace phase = k_primary (seed-independent part of the 64 samples of a pixel: ray, box cut, primary march, env colour / "
                      "hit voxel + normal; once per pixel and step) + k_trace_pt (token admission + secondary paths per pixel and frame), "
                      "one launch pair per 64-frame step",
            "traffic": traffic,
            "note": "achieved = the reference algorithm's bytes (SURVEY 8d: 15 B per march step, ...) for the 64 samples per pixel over the "
                    "measured trace time; the kernels move fewer bytes (one SDF byte per step instead of 15, the primary segment once per "
                    "pixel instead of 64 times, DESIGN.md 4.1), so HBM is not what bounds them: see `issue_bound`",
            "issue_bound": prof("issue_k_trace_pt.json")})
        lin["roofline"].update({
            "kernel": "the same launch pair with LINEAR instantiations: a quiet step is one 2-byte gather from the step field, an event "
                      "test seven filtered texture fetches (vr_quiet.cu, DESIGN.md 4.1)",
            "traffic": (prof("traffic_k_trace_lin.json") or {}).get("dram_bytes_per_launch"),
            "issue_bound": prof("issue_k_trace_pt_lin.json")})
        lin.update({"metric": "path_msamples_per_s", "unit": "Msamples/s", "e2e": e2e_lin,
                    "what": "the same step with vr_renderer_set_sampling(VR_SAMPLING_HW_LINEAR): volume, gradient taps and environment map "
                            "read the way NVIDIA hardware serves the reference's CLK_FILTER_LINEAR samplers — what the reference's OpenCL "
                            "kernels compute on this GPU (`reference_on_gpu`)"})
        if ref_gpu and "kernel_only" in ref_gpu and "per_frame_schedule" in lin:
            lin["vs_reference_on_gpu"] = {"batched_over_kernel_only": lin["value"] / ref_gpu["kernel_only"]["value"],
                                          "per_frame_schedule_over_kernel_only": lin["per_frame_schedule"]["value"] / ref_gpu["kernel_only"]["value"],
                                          "interactive_loop_over_render_frame_loop": lin["interactive_loop"]["value"] / ref_gpu["value"]}
        out = {
            "metric": "path_msamples_per_s", "value": main["value"], "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sampling": "VR_SAMPLING_NEAREST (the filter OpenCL defines for integer images); the "
                                                         "interpolating reading NVIDIA hardware executes is the `hw_linear` line-set",
                       "parallelism": (f"spp-split x{world}, weak: every rank traces its own 64 seeds of the scene with the full token cap of 256 "
                                       f"(the same work per rank as the single-GPU step), then vr_cache_allreduce sums the touched cache "
                                       f"entries — one per shaded pixel, lanes as uint32 words — over NCCL behind the C-ABI and every rank "
                                       f"resolves the {64 * world}-spp frame; the reference's cap of 256 then holds per rank, not for the sum: "
                                       f"the 1024-spp job with the cap of 256 for the sum is `c3_strong`") if world > 1 else "single GPU",
                       "l2": "inputs larger than L2 (cache 1 GiB + volume 256 MiB + SDF 128 MiB), no flush"},
            "clocks": main["clocks"], "e2e": e2e_main, "gpu_launches": main["gpu_launches"], "roofline": main["roofline"],
            "cpu_baseline": cpu, "reference_on_gpu": ref_gpu,
            "closeup": main.get("closeup"), "per_frame_schedule": main.get("per_frame_schedule"), "saturated_cache": main.get("saturated_cache"),
            "interactive_loop": main.get("interactive_loop"),
            "hw_linear": lin, "c3_strong": c3, "c4_tiles": c4, "c5_sweep": c5,
            "sdf_build_ms": {"value": float(np.median(sdf_ms)), "levels": levels, "volume": f"{VOL_N}^3",
                             "note": "vr_sdf_build wall time incl. allocation, excl. upload (app/sdf_benchmark.cpp:15-20)"},
        }
        emit(out)
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
    ap.add_argument("--skip-big", action="store_true", help="skip the 1024^3 legs (c4_tiles, c5_sweep)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    capture_stdout()
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        if args.warmup < 3:
            log(f"timing rules want at least 3 warm-up steps: running 3 instead of {args.warmup}")
            args.warmup = 3
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
