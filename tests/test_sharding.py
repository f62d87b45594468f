"""z-slab sharding (BASELINE config 5): stats / histogram / bilateral filter / SDF of a volume split into
z-slabs with halo planes equal the unsharded results.

CPU: the slab plan, and the exchange schedule (K levels, swap K+2 planes) on a numpy model of the wave over gloo with
world_size 2.  GPU: three "ranks" inside one process drive the real kernels through the C-ABI."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle_lib as o  # noqa: E402
from cl_volume_renderer_b200 import parallel, synth  # noqa: E402


def test_plan_slabs_covers_the_volume():
    for nz, world, halo in [(64, 2, 8), (100, 3, 8), (96, 8, 8), (17, 1, 8)]:
        plan = parallel.plan_slabs(nz, world, halo)
        assert plan[0][0] == 0 and plan[-1][1] == nz and plan[0][2] == 0 and plan[-1][3] == 0
        for (a0, a1, _, ahi), (b0, b1, blo, _) in zip(plan, plan[1:]):
            assert a1 == b0 and ahi == min(halo, nz - a1) and blo == min(halo, b0)
    assert parallel.global_max_it((512, 512, 512)) == 127 and parallel.global_max_it((38, 35, 38)) == 19


# ---- numpy model of the wave: R_k = R_{k-1} | dilate(R_{k-1}) with clamped corner neighbours -------------------------
def _dilate(R):
    def sh(a, axis, d):
        idx = np.clip(np.arange(a.shape[axis]) + d, 0, a.shape[axis] - 1)
        return np.take(a, idx, axis=axis)
    out = np.zeros_like(R)
    for dz in (-1, 1):
        for dy in (-1, 1):
            for dx in (-1, 1):
                out |= sh(sh(sh(R, 0, dz), 1, dy), 2, dx)
    return out


def _levels_of(R0, nlevels):
    R, lev = R0.copy(), np.where(R0, 1, 0).astype(np.int16)
    for k in range(1, nlevels + 1):
        new = _dilate(R) & ~R
        lev[new] = k + 1
        R |= new
    return lev


def _model_rank(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    band = rng.random((40, 12, 20)) < 0.01          # the global band bits
    K, halo, nlev = parallel.SDF_EXCHANGE_LEVELS, parallel.SDF_HALO, 17
    z0, z1, lo, hi = parallel.plan_slabs(band.shape[0], world, halo)[rank]
    R = band[z0 - lo: z1 + hi].copy()
    lev = np.where(R, 1, 0).astype(np.int16)
    n_own = z1 - z0
    down = ((lo, lo), (0, lo)) if lo else None
    up = ((lo + n_own - hi, hi), (lo + n_own, hi)) if hi else None
    done = 0
    while done < nlev:
        for _ in range(min(K, nlev - done)):
            new = _dilate(R) & ~R
            lev[new] = done + 2
            R |= new
            done += 1
        if done < nlev:
            t = torch.from_numpy(R.view(np.uint8))
            parallel.exchange_planes(t, down, up, rank, dist)
    q.put((rank, lev[lo: lo + n_own].copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_exchange_schedule_is_exact_on_the_numpy_model_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + os.getpid() % 90
    procs = [ctx.Process(target=_model_rank, args=(k, 2, port, q)) for k in range(2)]
    [p.start() for p in procs]
    parts = dict(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    rng = np.random.default_rng(5)
    band = rng.random((40, 12, 20)) < 0.01
    want = _levels_of(band, 17)
    assert np.array_equal(np.concatenate([parts[0], parts[1]], axis=0), want)


# ---- GPU: the real kernels, three ranks in one process ------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("dims,world,tfname", [((70, 45, 96), 3, "default"), ((64, 64, 64), 2, "thr"), ((40, 33, 90), 3, "grad")])
def test_slab_sharded_kernels_equal_unsharded(vr_ctx, dims, world, tfname):
    from cl_volume_renderer_b200 import api
    v = synth.synth_ct(0, dims=dims)
    tf = {"default": synth.default_tf(), "thr": synth.threshold_tf(300),
          "grad": [{"min_v": 100.0, "max_v": 1400.0, "min_g": 50.0, "max_g": 900.0, "flags": 1, "rgba": (255, 0, 0, 128)}]}[tfname]
    slabs = [parallel.SlabVolume(vr_ctx, v, r, world, parallel.SDF_HALO) for r in range(world)]
    # fetch_stats: MIN/MAX over the partials
    parts = np.array([s.vol.stats() for s in slabs])
    got = [int(parts[:, 0].min()), int(parts[:, 1].max()), int(parts[:, 2].min()), int(parts[:, 3].max())]
    assert got == o.fetch_stats(v)
    # histogram: SUM of the partial bins
    rng = [float(x) for x in got]
    bins = sum(s.vol.histogram(60, 50, rng).astype(np.int64) for s in slabs)
    assert np.array_equal(bins, o.histogram(v, 60, 50, rng).astype(np.int64))
    # SDF: K levels, halo swap, ... against the unsharded build and the oracle
    sdfs = [parallel.SlabSdf(s, tf) for s in slabs]
    pending = list(sdfs)
    while not all(s.sdf.finished for s in sdfs):
        for s in sdfs:
            s.sdf.advance(parallel.SDF_EXCHANGE_LEVELS)
        if all(s.sdf.finished for s in sdfs):
            break
        vr_ctx.synchronize()
        parallel.local_exchange(sdfs, "cuda:0")
        for s in sdfs:
            s.sdf.mark_imported()
    got_sdf = np.concatenate([s.download() for s in sdfs], axis=0)
    assert np.array_equal(got_sdf, o.sdf_build(v, tf)[0])
    [s.close() for s in sdfs]
    # bilateral filter (2-plane halo): identical to the unsharded GPU filter
    full = api.Volume(vr_ctx, v)
    full.filter()
    want_f = full.download()
    got_f = np.concatenate([parallel.bilateral(s) for s in slabs], axis=0)
    assert np.array_equal(got_f, want_f)
    full.close(); [s.close() for s in slabs]
    del pending
