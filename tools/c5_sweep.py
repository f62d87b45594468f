"""BASELINE config 5: SDF build + histogram + volume_filter sweep over volume sizes, z-slab sharded over the ranks of one node
with halo exchange (cl_volume_renderer_b200/parallel.py).  Launch:

    python tools/c5_sweep.py [sizes...]                                   (1 GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/c5_sweep.py 512 1024

Every rank generates the same synthetic volume (only its slab + halo is uploaded), times each phase on the device as the max
over ranks, checks the sharded SDF of the smallest size against the single-GPU build, and rank 0 prints one JSON line per size."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from cl_volume_renderer_b200 import api, parallel, synth
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dev = f"cuda:{lr}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    ctx = api.Context(lr)
    ext = torch.cuda.ExternalStream(ctx.stream, device=lr)
    sizes = [int(a) for a in sys.argv[1:]] or [128, 256, 512]
    tf = synth.default_tf()

    def tmax(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sync():
        ctx.synchronize(); torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for n in sizes:
        v = synth.synth_ct(n)
        slab = parallel.SlabVolume(ctx, v, rank, world, parallel.SDF_HALO)
        sync()

        def amm(x):
            t = torch.from_numpy(x.copy()).to(dev)
            if world > 1:
                lo = t[[0, 2]].clone(); hi = t[[1, 3]].clone()
                dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
                t = torch.stack([lo[0], hi[0], lo[1], hi[1]])
            return t.cpu().numpy()

        def asum(x):
            t = torch.from_numpy(x.astype(np.int64)).to(dev)
            if world > 1:
                dist.all_reduce(t)
            return t.cpu().numpy()

        st = parallel.stats(slab, amm)
        rng = [float(x) for x in st]
        t0 = time.perf_counter(); sync()
        t0 = time.perf_counter()
        bins = parallel.histogram(slab, 500, 500, rng, asum)
        sync(); hist_ms = tmax(1e3 * (time.perf_counter() - t0))
        assert int(bins.sum()) > 0

        def exchange(s):
            if world == 1:
                return
            t = parallel.bits_tensor(s, dev)
            down, up = s.boundary_planes()
            with torch.cuda.stream(ext):   # stream-ordered after the wave kernels, no host synchronisation
                parallel.exchange_planes(t, down, up, rank, dist)

        sdf_ms = []
        for rep in range(3):
            sync()
            t0 = time.perf_counter()
            s = parallel.SlabSdf(slab, tf)
            s.run(exchange)
            ctx.synchronize()
            sdf_ms.append(tmax(1e3 * (time.perf_counter() - t0)))
            if rep < 2:
                s.close()
        ok = None
        if n <= 256:   # parity of the sharded build against the single-GPU build of the whole volume
            mine = s.download()
            full = api.Volume(ctx, v); ref = api.Sdf(ctx, full, tf)
            ok = bool(np.array_equal(mine, ref.download()[slab.z0:slab.z1]))
            ref.close(); full.close()
            if world > 1:
                t = torch.tensor([int(ok)], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MIN); ok = bool(t.item())
        s.close()
        sync()
        t0 = time.perf_counter()
        parallel.bilateral(slab)
        sync(); filt_ms = tmax(1e3 * (time.perf_counter() - t0))
        slab.close()
        if rank == 0:
            print(json.dumps({"config": "c5 z-slab sweep", "n": n, "n_gpus": world, "stats": [int(x) for x in st],
                              "sdf_build_ms": min(sdf_ms), "sdf_build_ms_all": sdf_ms, "histogram_ms": hist_ms,
                              "bilateral_ms_incl_download": filt_ms, "sdf_parity_vs_single_gpu": ok,
                              "halo_planes": parallel.SDF_HALO, "levels_per_exchange": parallel.SDF_EXCHANGE_LEVELS}), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
