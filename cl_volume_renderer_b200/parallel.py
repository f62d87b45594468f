"""z-slab sharding of the volume kernels over the ranks of one node (BASELINE config 5).  The reference is
single-device; this is the host-side driver of the multi-GPU hooks of include/vr.h.  One process per GPU, torch.distributed
(NCCL over NVLink) carries the exchanges; `exchange` is injectable so that the same code runs with several "ranks" inside one
process (tests) or over gloo.

  plan_slabs      contiguous z-slabs + halo widths per rank
  stats / histogram   per-slab partials, combined with MIN/MAX and SUM all-reduces (4 ints; W*H uint32)
  bilateral       2-plane halo, no iteration, no exchange after the upload
  sdf             the wave goes stale from a slab's artificial ends by one plane per level (plus two at the start): run K
                  levels, swap the K+2 boundary planes of the bit volume with both neighbours, continue
"""
import numpy as np

from . import api

SDF_EXCHANGE_LEVELS = 14         # K: levels between two halo exchanges (9 swaps for the 125 levels of a >= 254^3 volume)
SDF_HALO = SDF_EXCHANGE_LEVELS + 2


def plan_slabs(nz, world, halo):
    """-> list of (z0, z1, lo, hi): rank r owns planes [z0, z1) and holds [z0 - lo, z1 + hi) (halo clipped at the faces)"""
    base, rem = divmod(nz, world)
    out, z = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        z0, z1 = z, z + n
        out.append((z0, z1, min(halo, z0), min(halo, nz - z1)))
        z = z1
    return out


def global_max_it(dims):
    """signed_distance_field.cpp:11 on the GLOBAL volume"""
    return min(max(dims) // 2, 127)


class SlabVolume:
    """A rank's part of the volume: api.Volume over planes [z0 - lo, z1 + hi) with the interior marked"""

    def __init__(self, ctx, volume_np, rank, world, halo):
        nz = volume_np.shape[0]
        self.dims_global = (volume_np.shape[2], volume_np.shape[1], nz)
        self.rank, self.world = rank, world
        self.z0, self.z1, self.lo, self.hi = plan_slabs(nz, world, halo)[rank]
        ext = np.ascontiguousarray(volume_np[self.z0 - self.lo: self.z1 + self.hi])
        self.vol = api.Volume(ctx, ext, interior=(self.lo, self.lo + (self.z1 - self.z0)))
        self.ctx = ctx

    def close(self):
        self.vol.close()


def stats(slab, allreduce_min_max):
    """fetch_stats of the global volume: allreduce_min_max(np.int32[4]) -> elementwise {min, max, min, max} over ranks"""
    return allreduce_min_max(np.array(slab.vol.stats(), dtype=np.int32))


def histogram(slab, width, height, rng, allreduce_sum):
    return allreduce_sum(slab.vol.histogram(width, height, rng))


def bilateral(slab):
    """5x5x5 bilateral filter of the rank's planes (needs halo >= 2): returns [z1-z0, ny, nx] int16"""
    assert (slab.lo >= 2 or slab.z0 == 0) and (slab.hi >= 2 or slab.z1 == slab.dims_global[2])
    slab.vol.filter()
    return slab.vol.download_planes(slab.lo, slab.z1 - slab.z0)


class SlabSdf:
    """level-by-level SDF of one rank's slab; `step(exchange)` runs K levels and one halo swap"""

    def __init__(self, slab, tf_specs):
        self.slab = slab
        self.n_own = slab.z1 - slab.z0
        # a rank forwards its `halo` lowest / highest OWN planes: with fewer own planes than that it would forward planes of its
        # other halo, which are stale after K levels, and the sharded field would be silently wrong
        if (slab.lo and self.n_own < slab.lo) or (slab.hi and self.n_own < slab.hi):
            raise ValueError(f"z-slab of {self.n_own} planes is thinner than the halo ({max(slab.lo, slab.hi)}): use fewer ranks "
                             f"(nz >= {SDF_HALO} * world) or the unsharded build")
        self.sdf = api.SdfSlab(slab.ctx, slab.vol, tf_specs, global_max_it(slab.dims_global))

    def boundary_planes(self):
        """(send_down, recv_down, send_up, recv_up) as (first plane, plane count) in slab coordinates; None at a face"""
        s = self.slab
        down = up = None
        if s.lo:   # my lowest `lo` interior planes go down; the neighbour's top planes arrive in my lower halo
            down = ((s.lo, s.lo), (0, s.lo))
        if s.hi:
            up = ((s.lo + self.n_own - s.hi, s.hi), (s.lo + self.n_own, s.hi))
        return down, up

    def run(self, exchange):
        """exchange(self): overwrite the halo planes of the current bit volume with the neighbours' boundary planes.
        STREAM CONTRACT: vr_sdf_slab_advance enqueues on the context's stream (api.Context.stream) and does not synchronise, so
        `exchange` must either run its copies / collectives on that same stream (torch.cuda.ExternalStream(ctx.stream), as
        tools/c5_sweep.py does) or call ctx.synchronize() first (as tests/test_sharding.py does); on any other stream the
        transfer races the wave kernels.  The C-ABI's vr_sdf_build_sharded does all of this inside the library."""
        while not self.sdf.finished:
            self.sdf.advance(SDF_EXCHANGE_LEVELS)
            if self.sdf.finished:
                break
            exchange(self)
            self.sdf.mark_imported()

    def download(self):
        return self.sdf.download(self.slab.lo, self.n_own)

    def close(self):
        self.sdf.close()


# ---- exchanges -------------------------------------------------------------------------------------------------------
class _DevArray:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 3}


def bits_tensor(slab_sdf, device):
    """torch view [nz_ext, plane_words] (int32) of the slab's CURRENT bit volume (changes buffer every level)"""
    import torch
    nz_ext = slab_sdf.sdf.dims[2]
    return torch.as_tensor(_DevArray(slab_sdf.sdf.bits_ptr, (nz_ext, slab_sdf.sdf.plane_words), "<i4"), device=device)


def exchange_planes(t, down, up, rank, dist):
    """t: tensor whose dim 0 is z (slab coordinates).  Sends the boundary interior planes to the z-neighbours and receives
    theirs into the halo planes, one batched P2P round (NCCL send/recv over NVLink, or gloo)."""
    ops = []
    if down is not None:
        (s0, n), (r0, m) = down
        ops += [dist.P2POp(dist.isend, t[s0:s0 + n], rank - 1), dist.P2POp(dist.irecv, t[r0:r0 + m], rank - 1)]
    if up is not None:
        (s0, n), (r0, m) = up
        ops += [dist.P2POp(dist.isend, t[s0:s0 + n], rank + 1), dist.P2POp(dist.irecv, t[r0:r0 + m], rank + 1)]
    for req in (dist.batch_isend_irecv(ops) if ops else []):
        req.wait()


def local_exchange(all_slab_sdfs, device):
    """all "ranks" live in this process on one device (tests): copy the planes directly"""
    import torch
    ts = [bits_tensor(s, device) for s in all_slab_sdfs]
    for r, s in enumerate(all_slab_sdfs):
        down, up = s.boundary_planes()
        if up is not None:   # my top interior planes -> the upper neighbour's lower halo, and its bottom planes -> my upper halo
            (s0, n), (r0, m) = up
            (ds0, dn), (dr0, dm) = all_slab_sdfs[r + 1].boundary_planes()[0]
            assert n == dm and m == dn
            ts[r + 1][dr0:dr0 + dm].copy_(ts[r][s0:s0 + n])
            ts[r][r0:r0 + m].copy_(ts[r + 1][ds0:ds0 + dn])
    torch.cuda.synchronize()
