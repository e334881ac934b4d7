"""GPU parity of the layer-1 conversions (BASELINE config 4: L0 -> L1 on a batch of frames)."""
import numpy as np
import pytest
import support as S

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import libllsm2_b200 as L
    c = L.Context(0)
    yield c
    c.close()


def _dev(d):
    import torch
    return {k: (torch.from_numpy(np.ascontiguousarray(v)).cuda() if v is not None else None) for k, v in d.items()}


@pytest.mark.parametrize("kw,nfft", [(dict(seed=3, nhar=100, maxnhar=100), 2048),
                                     (dict(seed=4, nhar=256, maxnhar=256, f0_lo=60, f0_hi=86), 2048),
                                     (dict(seed=5, nhar=60, maxnhar=80, thop=128 / 44100.0, f0_lo=150, f0_hi=300), 1024)])
def test_tolayer1_and_back(ctx, kw, nfft):
    import torch
    import libllsm2_b200 as L
    fr, conf = S.synth_frames(3, 80, **kw)
    ref = S.ref_tolayer1(fr, conf, nfft)
    o = L.tolayer1(ctx, conf, _dev(fr), nfft)
    torch.cuda.synchronize()
    o = {k: v.cpu().numpy() for k, v in o.items()}
    S.check_layer1(o, ref, fr["f0"] > 0)
    ref0 = S.ref_tolayer0(fr["f0"], ref, conf)
    o0 = L.tolayer0(ctx, conf, torch.from_numpy(fr["f0"]).cuda(), _dev(ref))
    torch.cuda.synchronize()
    S.check_layer0_from_l1({k: v.cpu().numpy() for k, v in o0.items()}, ref0)


def test_config4_batch_4096_frames(ctx):
    """BASELINE config 4 size: 4096 frames through L0 -> L1 -> L0; spot-check rows against the oracle."""
    import torch
    import libllsm2_b200 as L
    fr, conf = S.synth_frames(8, 512, seed=11, nhar=128, maxnhar=128)
    o = L.tolayer1(ctx, conf, _dev(fr), 2048)
    torch.cuda.synchronize()
    sub = {k: (v[:1] if v is not None else None) for k, v in fr.items()}
    c1 = S.abi.make_conf(1, 512, 128, conf.maxnhar_e, conf.npsd, conf.nchannel, conf.fs, conf.thop)
    ref = S.ref_tolayer1(sub, c1, 2048)
    S.check_layer1({k: v[:1].cpu().numpy() for k, v in o.items()}, ref, sub["f0"] > 0)
