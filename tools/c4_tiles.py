"""BASELINE config 4: 1024^3 volume (2 GiB) with SDF empty-space skipping, 3840x2160, 256 spp, image-tile split over the ranks
of one node.  Every rank holds the whole scene (2 GiB volume + 1 GiB SDF + 8 GiB voxel cache) and traces interleaved row
blocks with its own cache (vr_renderer_set_rows); the blocks are all-gathered into the full frame (NCCL).

    python tools/c4_tiles.py [n] [W] [H] [spp]                                                   (1 GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port P tools/c4_tiles.py

Rank 0 prints one JSON line: path Msamples/s of the whole job (device time of the slowest rank incl. the all-gather), SDF
build ms, and a checksum of the stitched frame."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class _DevArray:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}


def main():
    import torch
    import torch.distributed as dist
    from cl_volume_renderer_b200 import api, synth
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 3840
    H = int(sys.argv[3]) if len(sys.argv) > 3 else 2160
    spp = int(sys.argv[4]) if len(sys.argv) > 4 else 256
    torch.cuda.set_device(lr)
    dev = f"cuda:{lr}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    ctx = api.Context(lr)
    ext = torch.cuda.ExternalStream(ctx.stream, device=lr)
    v = synth.synth_ct(n)
    vol = api.Volume(ctx, v)
    env = api.EnvMap(ctx, synth.synth_env(2048, 1024))
    r = api.Renderer(ctx, W, H)
    r.image_set(vol, env); r.set_tf(synth.default_tf())
    ctx.synchronize(); t0 = time.perf_counter()
    r.flush_changes()
    ctx.synchronize(); flush_ms = 1e3 * (time.perf_counter() - t0)
    pos, d = synth.closeup_camera(n)
    seeds = synth.glibc_rand(spp)
    # interleaved blocks of BLOCK rows: block b belongs to rank b % world (miss rows are cheap, shaded rows are not)
    BLOCK = 24
    nblocks = (H + BLOCK - 1) // BLOCK
    mine = [b for b in range(nblocks) if b % world == rank]
    frame_t = torch.as_tensor(_DevArray(r.frame_device_ptr, W * H, "<i4"), device=dev).view(H, W)

    def job():
        r.reset_cache()
        for b in mine:
            r.set_rows(b * BLOCK, min(H, (b + 1) * BLOCK))
            r.render_frames(pos, d, seeds, readback=False)
        if world > 1:
            with torch.cuda.stream(ext):   # every block is rank-owned: a SUM over ranks of frames that are zero elsewhere stitches them
                dist.all_reduce(frame_t)

    def sync():
        ctx.synchronize(); torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # rows outside a rank's blocks stay zero in its frame buffer (vr_renderer_create clears it, tracing only writes its rows)
    job(); sync()
    frame_t.zero_(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext); job(); e1.record(ext); e1.synchronize(); sync()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    frame = frame_t.cpu().numpy().view(np.uint8).reshape(H, W, 4)
    if rank == 0:
        print(json.dumps({"config": "c4 image-tile split", "volume": f"{n}^3", "frame": f"{W}x{H}", "spp": spp, "n_gpus": world,
                          "value": W * H * spp / ms / 1e3, "unit": "Msamples/s", "ms": ms, "flush_ms_incl_sdf_build": flush_ms,
                          "shaded_fraction": float((frame[..., 3] == 1).mean()), "env_fraction": float((frame[..., 3] == 200).mean()),
                          "frame_checksum": int(frame.astype(np.uint64).sum()), "rows_per_block": BLOCK,
                          "memory_per_rank_gib": round((2 * n ** 3 + n ** 3 + 8 * n ** 3) / 2 ** 30, 2)}), flush=True)
    r.close(); env.close(); vol.close(); ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
