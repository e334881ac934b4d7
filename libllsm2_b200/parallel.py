"""Multi-GPU layer: one process per GPU (torch.distributed, NCCL over NVLink / NVSwitch).

Two partitions of the hot path:

* utterance sharding -- utterances are independent (one llsm_chunk each), every rank runs the
  single-GPU path on its slice of the batch, no data-path collective (bench.py --gpus N);
* frame-range sharding -- one batch of long utterances, rank r synthesises frames
  [lo_r, hi_r) of every utterance; a frame reaches at most `halo` samples beyond its centre, so the
  partial sums only have to be completed in strips of `halo` samples around the shard boundaries:
  ONE all-gather of the boundary strips of (y_sin, y_noise), then a local add.
"""
import torch
import torch.distributed as dist

from ._lib import lib


def frame_shards(nfrm, world):
    """Contiguous, near-equal frame ranges [(lo, hi)] for `world` ranks."""
    edges = [round(r * nfrm / world) for r in range(world + 1)]
    return [(edges[r], edges[r + 1]) for r in range(world)]


def shard_sample_range(conf, lo, hi, nfrm, ny):
    """Output samples owned by the shard of frames [lo, hi): [position(lo), position(hi)), first shard
    from 0, last shard to ny."""
    L = lib()
    import ctypes as C
    sa = 0 if lo == 0 else L.llsm_b200_frame_position(lo, C.c_float(conf.thop), C.c_float(conf.fs))
    sb = ny if hi >= nfrm else L.llsm_b200_frame_position(hi, C.c_float(conf.thop), C.c_float(conf.fs))
    return sa, sb


def exchange_halos(partial, conf, rank, world, halo=None, group=None):
    """Complete the partial sums of a frame-range shard.

    partial: dict(y_sin, y_noise) of [B][ny] tensors holding this rank's partial sums (any device).
    Returns dict(y_sin, y_noise, y) restricted to the samples this rank owns ([B][sb - sa]) and (sa, sb).
    One all_gather of a [B][2 sides][2 components][halo] tensor per call.
    """
    import ctypes as C
    nfrm = conf.nfrm
    ny = partial["y_sin"].shape[1]
    if halo is None:
        halo = lib().llsm_b200_halo_length(C.byref(conf))
    shards = frame_shards(nfrm, world)
    lo, hi = shards[rank]
    sa, sb = shard_sample_range(conf, lo, hi, nfrm, ny)
    assert sb - sa >= halo, "shard shorter than the halo: use fewer ranks or longer utterances"
    B = partial["y_sin"].shape[0]
    dev = partial["y_sin"].device
    strips = torch.zeros((B, 2, 2, halo), dtype=torch.float32, device=dev)
    for c, k in enumerate(("y_sin", "y_noise")):
        p = partial[k]
        l0 = max(sa - halo, 0)
        strips[:, 0, c, halo - (sa - l0):] = p[:, l0:sa]                 # spill into the left neighbour
        r1 = min(sb + halo, ny)
        strips[:, 1, c, :r1 - sb] = p[:, sb:r1]                          # spill into the right neighbour
    gathered = [torch.empty_like(strips) for _ in range(world)]
    if world > 1:
        dist.all_gather(gathered, strips, group=group)
    else:
        gathered = [strips]
    out = {}
    for c, k in enumerate(("y_sin", "y_noise")):
        own = partial[k][:, sa:sb].clone()
        if rank > 0:                                                     # left neighbour's right strip
            own[:, :halo] += gathered[rank - 1][:, 1, c, :]
        if rank < world - 1:                                             # right neighbour's left strip
            own[:, -halo:] += gathered[rank + 1][:, 0, c, :]
        out[k] = own
    out["y"] = out["y_sin"] + out["y_noise"]
    return out, (sa, sb)
