"""Where a weak-scaling step of the spp split goes (run under torchrun on N GPUs): trace alone, exchange alone (ranks in step), both.
torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/xchg_probe.py"""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from cl_volume_renderer_b200 import api, synth
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
ctx = api.Context(lr)
if world > 1:
    dist.init_process_group("gloo")
    box = [api.Context.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx.comm_init(rank, world, box[0])
ext = torch.cuda.ExternalStream(ctx.stream, device=lr)
n, W, H, SPP = 512, 1920, 1080, 64
vol = api.Volume(ctx, synth.synth_ct(n)); env = api.EnvMap(ctx, synth.synth_env(2048, 1024))
r = api.Renderer(ctx, W, H); r.image_set(vol, env); r.set_tf(synth.default_tf()); r.flush_changes()
pos, d = synth.default_camera(n)
seeds = synth.glibc_rand(SPP * world)[SPP * rank: SPP * (rank + 1)]
def barrier():
    ctx.synchronize(); torch.cuda.synchronize()
    if world > 1: ctx.comm_barrier()
def timed(fn, steps, warmup=3):
    for _ in range(warmup): fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for _ in range(steps): fn()
    e1.record(ext); e1.synchronize(); barrier()
    ms = e0.elapsed_time(e1) / steps
    return float(ctx.comm_allreduce(np.array([ms], dtype=np.float64), "max")[0]) if world > 1 else ms
def trace():
    r.reset_cache(); r.render_frames(pos, d, seeds, readback=False)
def both():
    trace(); r.cache_allreduce()
def xchg():
    r.cache_allreduce()
out = {"world": world, "trace_ms": timed(trace, 10), "step_ms": timed(both, 10)}
trace(); ctx.synchronize()
out["exchange_alone_ms"] = timed(xchg, 20)
t0 = time.perf_counter(); 
for _ in range(20): xchg()
ctx.synchronize(); out["exchange_alone_wall_ms"] = 1e3 * (time.perf_counter() - t0) / 20
if rank == 0: print(json.dumps(out))
r.close(); env.close(); vol.close()
