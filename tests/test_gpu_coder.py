"""GPU parity of the frame coder (llsm_b200_coder_encode / _decode) against the reference build (coder.c), and a
full-size batch (BASELINE config 4: 4096 frames) through encode -> decode."""
import numpy as np
import pytest
import support as S
from test_emu_coder import OS, OB, _frames

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import torch
    import libllsm2_b200 as L
    assert torch.cuda.is_available()
    c = L.Context(0)
    yield c
    c.close()


def _dev(d):
    import torch
    return {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in d.items() if v is not None}


def test_encode_and_decode_match_reference(ctx):
    import torch
    import libllsm2_b200 as L
    fr, conf, l1 = _frames(seed=7, B=2, F=12)
    nspec = l1["vtmagn"].shape[-1]
    ref = S.ref_coder_encode(fr["f0"], fr["psd"], l1, conf, OS, OB)
    d, dl = _dev(fr), _dev(l1)
    enc = L.coder_encode(ctx, conf, d["f0"], d["psd"], dl, OS, OB)
    torch.cuda.synchronize()
    S.check_coder_encode(enc.cpu().numpy(), ref, OS)
    for use_layer1 in (1, 0):
        rd = S.ref_coder_decode(ref, conf, nspec, OS, OB, use_layer1)
        o = L.coder_decode(ctx, conf, torch.from_numpy(ref).cuda(), nspec, OS, OB, bool(use_layer1))
        torch.cuda.synchronize()
        S.check_coder_decode({k: v.cpu().numpy() for k, v in o.items()}, rd, use_layer1)


def test_full_size_round_trip(ctx):
    """4096 frames (BASELINE config 4 size): encode, decode to layer 1, encode again -- the coder is (nearly)
    idempotent on its own output, a size-independent property."""
    import torch
    import libllsm2_b200 as L
    from libllsm2_b200 import abi
    fr, conf, l1 = _frames(seed=8, B=1, F=16)
    nspec = l1["vtmagn"].shape[-1]
    rep = 256
    tile = lambda v: np.ascontiguousarray(np.tile(v, (rep,) + (1,) * (v.ndim - 1)))
    cb = abi.make_conf(rep, conf.nfrm, conf.maxnhar, conf.maxnhar_e, conf.npsd, conf.nchannel, conf.fs, conf.thop,
                       list(conf.chanfreq)[:conf.nchannel - 1], conf.lip_radius)
    f0, psd = torch.from_numpy(tile(fr["f0"])).cuda(), torch.from_numpy(tile(fr["psd"])).cuda()
    dl = {"rd": torch.from_numpy(tile(l1["rd"])).cuda(), "vtmagn": torch.from_numpy(tile(l1["vtmagn"])).cuda()}
    e1 = L.coder_encode(ctx, cb, f0, psd, dl, OS, OB)
    o = L.coder_decode(ctx, cb, e1, nspec, OS, OB, True)
    vt = torch.nan_to_num(o["vtmagn"], neginf=-400.0)
    e2 = L.coder_encode(ctx, cb, o["f0"], o["psd"], {"rd": o["rd"], "vtmagn": vt}, OS, OB)
    torch.cuda.synchronize()
    assert torch.equal(e1[0], e1[rep - 1])                         # every copy of the utterance encodes alike
    assert torch.equal(e1[..., :3], e2[..., :3])                   # voicing, f0, Rd survive exactly
    v = e1[..., 0] > 0
    d = (e1[..., 3:3 + OS] - e2[..., 3:3 + OS]).abs()
    assert float(d[v][:, :24].median()) < 0.05                     # low-order mel-cepstrum: log-intensity units
