"""Plain torch restatement of the halo exchange of frame-range shards (llsm_b200_halo_exchange in the C library:
pack kernel -> ncclAllGather -> edge-add kernel), for the CPU-side world_size-2 gloo tests. Test infrastructure only."""
import torch
import torch.distributed as dist

from libllsm2_b200._lib import lib
from libllsm2_b200.parallel import frame_shards, shard_sample_range


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


