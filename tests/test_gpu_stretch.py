"""GPU parity of the frame interpolation / time-stretch operation (llsm_b200_frames_stretch through the C ABI) against
the reference's own interp_llsm_frame (test/demo-stretch.c:50-129, compiled into the reference build), plus
size-independent properties at the BASELINE config 2 batch size and the demo's whole device-resident pipeline."""
import numpy as np
import pytest
import support as S
from test_emu_stretch import _jitter

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import torch
    import libllsm2_b200 as L
    assert torch.cuda.is_available()
    c = L.Context(0)
    yield c
    c.close()


def _cuda(d):
    import torch
    return {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in d.items() if v is not None}


def _gpu_stretch(ctx, fr, conf, l1, base, ratio, res):
    import torch
    import libllsm2_b200 as L
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda() if a is not None else None  # noqa: E731
    cn, fo, lo = L.frames_stretch(ctx, conf, _cuda(fr), _cuda(l1), t(base), t(ratio), t(res))
    torch.cuda.synchronize()
    o = {k: v.cpu().numpy() for k, v in fo.items()}
    o.update({k: v.cpu().numpy() for k, v in lo.items()})
    return cn, o


@pytest.mark.parametrize("factor", [2.0, 0.6, 1.0])
def test_stretch_matches_reference(ctx, factor):
    import libllsm2_b200 as L
    fr, conf, l1 = S.stretch_case(B=3, F=40, seed=51)
    Fn = int(round(conf.nfrm * factor))
    base, ratio, res = L.stretch_map(conf.nfrm, Fn)
    res = _jitter(res, conf.nfrm, 6)
    ref = S.ref_stretch(fr, conf, l1, base, ratio, res)
    cn, o = _gpu_stretch(ctx, fr, conf, l1, base, ratio, res)
    assert cn.nfrm == Fn and cn.nutt == conf.nutt
    S.check_stretch(o, ref, exact=False)
    assert np.array_equal(o["ampl"], fr["ampl"][:, base])


def test_every_transition_at_edge_ratios(ctx):
    fr, conf, l1 = S.stretch_case(B=1, F=24, seed=21)
    F = conf.nfrm
    base = np.repeat(np.arange(F - 1, dtype=np.int32), 3)
    ratio = np.tile(np.array([0.0, 0.37, 1.0], np.float32), F - 1)
    ref = S.ref_stretch(fr, conf, l1, base, ratio, base)
    _, o = _gpu_stretch(ctx, fr, conf, l1, base, ratio, None)
    S.check_stretch(o, ref, exact=False)


def test_full_size_identity_and_midpoint(ctx):
    """1024 x 400 frames: a map with ratio 0 returns the source rows wherever neither quirk applies (both frames voiced
    or both unvoiced, magnitudes above the -80 dB floor); the midpoint of two voiced frames lies between them."""
    import torch
    import libllsm2_b200 as L
    from libllsm2_b200 import abi
    fr, conf, l1 = S.stretch_case(B=2, F=100, seed=61)
    rep = 512
    tile = lambda v: np.tile(v, (rep,) + (1,) * (v.ndim - 1))  # noqa: E731
    conf_b = abi.make_conf(2 * rep, conf.nfrm, conf.maxnhar, conf.maxnhar_e, conf.npsd, conf.nchannel, conf.fs, conf.thop,
                           list(conf.chanfreq)[:conf.nchannel - 1], conf.lip_radius)
    d = _cuda({k: tile(v) for k, v in fr.items() if v is not None})
    dl = _cuda({k: tile(v) for k, v in l1.items()})
    F = conf.nfrm
    base = torch.arange(F - 1, dtype=torch.int32).repeat(4).cuda()                      # 396 output frames
    ratio = torch.cat([torch.zeros(2 * (F - 1)), torch.full((2 * (F - 1),), 0.5)]).cuda()
    cn, fo, lo = L.frames_stretch(ctx, conf_b, d, dl, base, ratio, None)
    torch.cuda.synchronize()
    assert torch.equal(fo["psd"][:, :F - 1], d["psd"][:, :F - 1])
    assert torch.equal(fo["psdres"][:, :F - 1], d["psdres"][:, :F - 1])
    f0 = d["f0"]
    both = ((f0[:, :-1] > 0) & (f0[:, 1:] > 0))
    assert torch.equal(fo["f0"][:, :F - 1][both], f0[:, :-1][both])
    assert torch.equal(lo["vtmagn"][:, :F - 1][both], torch.clamp(dl["vtmagn"][:, :-1][both], min=-80.0))
    mid = fo["f0"][:, 2 * (F - 1):3 * (F - 1)][both]
    lo_f, hi_f = torch.minimum(f0[:, :-1], f0[:, 1:])[both], torch.maximum(f0[:, :-1], f0[:, 1:])[both]
    assert bool(((mid >= lo_f) & (mid <= hi_f)).all())
    assert torch.equal(fo["f0"][:2], fo["f0"][2 * (rep - 1):])                          # copies of an utterance agree


def test_demo_pipeline_stays_on_the_device(ctx):
    """test/demo-stretch.c:160-187 with device-resident arrays: tolayer1 -> phasepropagate(-1) -> stretch x2 ->
    tolayer0 -> phasepropagate(+1) -> synthesize. The stretched utterance is twice as long, finite, and carries about the
    same power as the unstretched layer 0 -> 1 -> 0 round trip (which on these synthetic frames is not the power of the
    input: the reference's envelope passes over the scattered harmonic amplitudes)."""
    import torch
    import libllsm2_b200 as L
    fr, conf = S.synth_frames(2, 120, seed=71, nhar=100, maxnhar=128, f0_lo=100, f0_hi=250)
    d = _cuda(fr)
    l1 = L.tolayer1(ctx, conf, d, 2048)
    back = dict(d)
    back.update(L.tolayer0(ctx, conf, d["f0"], l1))                   # the unstretched layer-1 round trip
    y0 = L.synthesize_l0(ctx, conf, back, seed=3)["y"]
    L.chunk_phasepropagate(ctx, conf, d, l1, sign=-1)
    base, ratio, res = L.stretch_map(conf.nfrm, 2 * conf.nfrm)
    t = lambda a: torch.from_numpy(a).cuda()  # noqa: E731
    cn, fo, lo = L.frames_stretch(ctx, conf, d, l1, t(base), t(ratio), t(_jitter(res, conf.nfrm, 7)))
    h = L.tolayer0(ctx, cn, fo["f0"], lo)
    fo.update(nhar=h["nhar"], ampl=h["ampl"], phse=h["phse"])
    L.chunk_phasepropagate(ctx, cn, fo, None, sign=1)
    y1 = L.synthesize_l0(ctx, cn, fo, seed=3)["y"]
    torch.cuda.synchronize()
    assert y1.shape[1] >= 2 * y0.shape[1] - 1000 and bool(torch.isfinite(y1).all())
    p0, p1 = float((y0 ** 2).mean()), float((y1 ** 2).mean())
    assert 0.5 < p1 / p0 < 2.0, (p0, p1)
