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


def utterance_shards(nutt, world):
    """Contiguous, near-equal utterance ranges [(lo, hi)] for `world` ranks."""
    edges = [round(r * nutt / world) for r in range(world + 1)]
    return [(edges[r], edges[r + 1]) for r in range(world)]


def scatter_utterances(conf, frames, rank, world, src=0, group=None, device=None):
    """Utterance sharding of a batch that lives on one rank: `src` cuts the batch into contiguous utterance
    ranges, serialises each into one flat blob (llsm_b200_frames_pack: SURVEY.md 8(f) rank 4) and sends it to its
    rank -- two collectives in all (blob sizes, then the blobs), no per-array traffic. conf / frames are only read
    on `src` (numpy arrays keyed as api.FRAME_KEYS). Returns (conf_local, frames_local) on every rank, the arrays
    being views into the received blob."""
    import numpy as np
    from . import api, abi
    dev = device if device is not None else torch.device("cpu")
    blobs = None
    if rank == src:
        blobs = []
        for lo, hi in utterance_shards(conf.nutt, world):
            c = abi.make_conf(hi - lo, conf.nfrm, conf.maxnhar, conf.maxnhar_e, conf.npsd, conf.nchannel, conf.fs,
                              conf.thop, list(conf.chanfreq)[:conf.nchannel - 1], conf.lip_radius)
            part = {k: (np.ascontiguousarray(v[lo:hi]) if v is not None else None) for k, v in frames.items()}
            blobs.append(torch.from_numpy(api.frames_to_blob(c, part)))
    sizes = torch.zeros(world, dtype=torch.int64, device=dev)
    if rank == src:
        sizes = torch.tensor([b.numel() for b in blobs], dtype=torch.int64, device=dev)
    if world > 1:
        dist.broadcast(sizes, src=src, group=group)
    n = int(sizes.max().item())
    mine = torch.zeros(n, dtype=torch.uint8, device=dev)
    if world > 1:
        padded = None
        if rank == src:
            padded = [torch.cat([b, torch.zeros(n - b.numel(), dtype=torch.uint8)]).to(dev) for b in blobs]
        dist.scatter(mine, padded, src=src, group=group)
    else:
        mine = blobs[0].to(dev)
    blob = mine[:int(sizes[rank].item())].cpu().numpy()
    return api.blob_to_frames(blob)
