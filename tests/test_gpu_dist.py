"""Frame-range sharding on real GPUs: llsm_b200_synthesize_l0_shard + llsm_b200_halo_exchange (pack kernel ->
ncclAllGather -> edge-add kernel, all inside the C library) against the unsharded synthesis.
world = 1 runs on any box (it proves the library binds NCCL, builds its communicator and runs both kernels);
world = 2 needs two GPUs and skips otherwise (tools/dist_check.py is the same check under torchrun, any world)."""
import os
import socket

import numpy as np
import pytest

import support as S

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    import libllsm2_b200 as L
    from libllsm2_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    fr, conf = S.synth_frames(3, 60 * world, seed=41, nhar=64, maxnhar=64)
    dev = torch.device("cuda", rank)
    d = {k: (torch.from_numpy(v).to(dev) if v is not None else None) for k, v in fr.items()}
    ctx = L.Context(rank)
    lo, hi = parallel.frame_shards(conf.nfrm, world)[rank]
    part = L.synthesize_l0_shard(ctx, conf, d, lo, hi, white=None, seed=99)
    ex = parallel.HaloExchange(ctx, conf, world, rank, part["y"].shape[1])
    sa, sb = ex.exchange(part)
    full = L.synthesize_l0(ctx, conf, d, white=None, seed=99)
    torch.cuda.synchronize()
    err = {k: float((part[k][:, sa:sb] - full[k][:, sa:sb]).abs().max()) for k in ("y", "y_sin", "y_noise")}
    q.put((rank, sa, sb, err, float(full["y"].pow(2).mean().sqrt())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [1, 2])
def test_sharded_equals_unsharded(world):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=600) for _ in range(world)], key=lambda t: t[0])
    for p in ps:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res[0][1] == 0
    for a, b in zip(res[:-1], res[1:]):
        assert a[2] == b[1]                                    # the owned ranges tile the output
    for r in res:
        assert r[4] > 1e-3
        # identical up to the association of the float sums at the seams
        assert all(v < 1e-6 for v in r[3].values()), r
