"""Multi-GPU layer: one process per GPU (torch.distributed, NCCL over NVLink / NVSwitch).

Two partitions of the hot path:

* utterance sharding -- utterances are independent (one llsm_chunk each), every rank runs the
  single-GPU path on its slice of the batch, no data-path collective (bench.py --gpus N);
* frame-range sharding -- one batch of long utterances, rank r synthesises frames
  [lo_r, hi_r) of every utterance; a frame reaches at most `halo` samples beyond its centre, so the
  partial sums only have to be completed in strips of `halo` samples around the shard boundaries:
  ONE all-gather of the boundary strips of (y_sin, y_noise), then a local add -- llsm_b200_halo_exchange of the C
  library (HaloExchange below); tests/dist_model.py keeps a plain torch restatement of it for the CPU (gloo) tests.
"""
import torch
import torch.distributed as dist

from ._lib import lib, check


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


class HaloExchange:
    """The exchange that completes frame-range shards, done by the C library (llsm_b200_halo_exchange: pack kernel ->
    ncclAllGather -> edge-add kernel on the context's stream). torch.distributed is only the courier of the
    ncclUniqueId: rank 0 draws it (llsm_b200_comm_unique_id), a broadcast hands it to the others, every rank then
    joins the library's own communicator (llsm_b200_comm_init)."""

    def __init__(self, ctx, conf, world, rank, stride, group=None, edges=None):
        import ctypes as C
        import numpy as np
        self.ctx, self.conf, self.world, self.rank, self.stride = ctx, conf, world, rank, stride
        L = lib()
        self.halo = L.llsm_b200_halo_length(C.byref(conf))
        self.bytes_per_rank = conf.nutt * 4 * self.halo * 4
        self.edges = np.asarray(edges if edges is not None else
                                [e for e, _ in frame_shards(conf.nfrm, world)] + [conf.nfrm], np.int32)
        ident = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_char * 128)()
            check(L.llsm_b200_comm_unique_id(buf))
            ident = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        if world > 1:
            dev = torch.device("cuda", ctx.device) if dist.get_backend(group) == "nccl" else torch.device("cpu")
            t = ident.to(dev)
            dist.broadcast(t, src=0, group=group)
            ident = t.cpu()
        raw = bytes(ident.numpy().tobytes())
        check(L.llsm_b200_comm_init(ctx._h, raw, int(rank), int(world)))

    def sample_range(self):
        import ctypes as C
        L = lib()
        return (L.llsm_b200_shard_position(C.byref(self.conf), int(self.edges[self.rank])),
                L.llsm_b200_shard_position(C.byref(self.conf), int(self.edges[self.rank + 1])))

    def exchange(self, out, lo=None, hi=None):
        """out: dict(y, y_sin, y_noise) of [B][stride] CUDA tensors holding this rank's partial sums; on return the
        samples this rank owns (sample_range()) are complete in all three."""
        import ctypes as C
        from . import abi
        o = abi.Output()
        o.y = out["y"].data_ptr() if out.get("y") is not None else None
        o.y_sin, o.y_noise, o.stride = out["y_sin"].data_ptr(), out["y_noise"].data_ptr(), out["y_sin"].shape[1]
        check(lib().llsm_b200_halo_exchange(self.ctx._h, C.byref(self.conf), C.byref(o),
                                            self.edges.ctypes.data_as(C.c_void_p)))
        return self.sample_range()


def bind_to_gpu_numa_node(local_rank):
    """Pin this process (and, by first touch, the pinned host buffers it allocates next) to the CPUs of the NUMA node
    the GPU hangs off, so that the ranks of one box do not all stage their host traffic through node 0. Best effort:
    returns the cpu list it bound to, or None when the topology cannot be read."""
    import os
    try:
        props = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return None
        cpus = []
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = set(os.sched_getaffinity(0))
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


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
