"""N > 1 path on CPU: world_size-2 gloo run of the frame-range sharding (halo all-gather + add) with the
kernel emulator producing each rank's partial sums; the stitched result must equal the unsharded run."""
import ctypes as C
import os
import socket
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import support as S
from libllsm2_b200 import abi, parallel


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _emu_synth(conf, fr, white, lo=None, hi=None):
    emu = S.load_emu()
    ny = S.load_ref().ref_output_length(conf.nfrm, C.c_float(conf.thop), C.c_float(conf.fs)) if os.path.exists(
        os.path.join(S.ROOT, "oracle", "_ref", "libllsm2_ref.so")) else None
    from libllsm2_b200 import output_length
    ny = output_length(conf.nfrm, conf.thop, conf.fs)
    B = conf.nutt
    oy = np.zeros((B, ny), np.float32); oys = np.zeros_like(oy); oyn = np.zeros_like(oy)
    out = abi.Output(); out.y = oy.ctypes.data; out.y_sin = oys.ctypes.data; out.y_noise = oyn.ctypes.data; out.stride = ny
    so = abi.default_soptions(white.ctypes.data, 0)
    f = S.frames_struct(fr)
    if lo is None:
        assert emu.emu_synthesize_l0(C.byref(conf), C.byref(f), C.byref(so), C.byref(out)) == 0
    else:
        assert emu.emu_synthesize_l0_shard(C.byref(conf), C.byref(f), C.byref(so), C.byref(out), lo, hi) == 0
    return oy, oys, oyn


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    os.environ["LLSM_EMU_NOBUILD"] = "1"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fr, conf = S.synth_frames(2, 24, seed=31, nhar=40, maxnhar=40)
    rng = np.random.default_rng(5)
    white = rng.standard_normal((2, conf.nchannel, 20128 if False else min(20000, parallel_ny(conf)) + 128)).astype(np.float32)
    lo, hi = parallel.frame_shards(conf.nfrm, world)[rank]
    y, ys, yn = _emu_synth(conf, fr, white, lo, hi)
    # (1) the library's exchange: pack kernel -> all-gather -> edge-add kernel (kernels_halo.cuh on the CPU emulator,
    #     gloo standing in for ncclAllGather; same strip layout and call order as llsm_b200_halo_exchange)
    from libllsm2_b200._lib import lib
    L = lib()
    emu = S.load_emu()
    B, ny = ys.shape
    halo = L.llsm_b200_halo_length(C.byref(conf))
    edges = [e for e, _ in parallel.frame_shards(conf.nfrm, world)] + [conf.nfrm]
    spos = np.array([L.llsm_b200_shard_position(C.byref(conf), e) for e in edges], np.int32)
    ks, kn, ky = ys.copy(), yn.copy(), np.zeros_like(ys)
    strips = np.zeros((B, 2, 2, halo), np.float32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    emu.emu_halo_pack(B, ny, ny, halo, rank, world, vp(spos), vp(ks), vp(kn), vp(strips))
    gathered = [torch.empty(strips.shape, dtype=torch.float32) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(strips))
    g = np.ascontiguousarray(torch.stack(gathered).numpy())
    emu.emu_halo_add(B, ny, ny, halo, rank, world, vp(spos), vp(ks), vp(kn), vp(ky), vp(g))
    ksa, ksb = int(spos[rank]), int(spos[rank + 1])
    # (2) the plain torch restatement of the same exchange
    import dist_model
    part = {"y_sin": torch.from_numpy(ys), "y_noise": torch.from_numpy(yn)}
    out, (sa, sb) = dist_model.exchange_halos(part, conf, rank, world)
    assert (sa, sb) == (ksa, ksb)
    for a, b2 in ((ks, out["y_sin"]), (kn, out["y_noise"]), (ky, out["y"])):
        assert np.array_equal(a[:, sa:sb], b2.numpy())
    q.put((rank, sa, sb, ks[:, sa:sb].copy(), kn[:, sa:sb].copy(), ky[:, sa:sb].copy()))
    dist.barrier()
    dist.destroy_process_group()


def parallel_ny(conf):
    from libllsm2_b200 import output_length
    return output_length(conf.nfrm, conf.thop, conf.fs)


def test_frame_sharding_world2_gloo():
    S.load_emu()                      # build once, before forking
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    fr, conf = S.synth_frames(2, 24, seed=31, nhar=40, maxnhar=40)
    rng = np.random.default_rng(5)
    white = rng.standard_normal((2, conf.nchannel, min(20000, parallel_ny(conf)) + 128)).astype(np.float32)
    y, ys, yn = _emu_synth(conf, fr, white)
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == y.shape[1]      # ranges tile the output
    ys_s = np.concatenate([r[3] for r in res], 1); yn_s = np.concatenate([r[4] for r in res], 1)
    y_s = np.concatenate([r[5] for r in res], 1)
    # identical up to the association of float sums at the seam (1 ulp)
    assert np.abs(ys_s - ys).max() < 1e-6 and np.abs(yn_s - yn).max() < 1e-6 and np.abs(y_s - y).max() < 1e-6
    assert np.abs(ys).max() > 1e-3


def test_shard_ranges():
    assert parallel.frame_shards(400, 8)[0] == (0, 50) and parallel.frame_shards(400, 8)[-1] == (350, 400)
    assert parallel.frame_shards(7, 2) == [(0, 4), (4, 7)]


def _scatter_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fr, conf = (S.synth_frames(5, 12, seed=61) if rank == 0 else (None, None))
    c, f = parallel.scatter_utterances(conf, fr, rank, world)
    q.put((rank, c.nutt, {k: (None if v is None else np.array(v)) for k, v in f.items()}))
    dist.barrier()
    dist.destroy_process_group()


def test_utterance_scatter_world2():
    """Utterance sharding through the flat blob (llsm_b200_frames_pack / _unpack): rank 0 owns the batch, every
    rank ends up with its contiguous range of utterances, bit for bit."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_scatter_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    got = dict()
    for _ in range(2):
        r, n, f = q.get(timeout=240)
        got[r] = (n, f)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    fr, conf = S.synth_frames(5, 12, seed=61)
    shards = parallel.utterance_shards(5, 2)
    for r, (lo, hi) in enumerate(shards):
        n, f = got[r]
        assert n == hi - lo
        for k, v in fr.items():
            if v is None:
                assert f[k] is None
            else:
                assert np.array_equal(f[k], v[lo:hi]), (r, k)
