"""GPU parity of the chunk phase utilities (llsm_b200_chunk_phasepropagate / _phasesync_rps) against the
reference build (layer0.c:687-706, frame.c:152-178): bit-exact, and the analysis -> propagate -> synthesis use
of test/test-layer0-anasynth.c:62-63 stays device-resident."""
import numpy as np
import pytest
import support as S
from test_emu_phase import _source_phases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import torch
    import libllsm2_b200 as L
    assert torch.cuda.is_available()
    c = L.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("mode,arg", [(0, 1), (0, -1), (1, 0), (1, 1)])
def test_phase_ops_match_reference(ctx, mode, arg):
    import torch
    import libllsm2_b200 as L
    fr, conf = S.synth_frames(3, 120, seed=41)
    fr["nfrm_utt"] = np.asarray([120, 1, 77], np.int32)
    vs, nvs = _source_phases(fr, conf, 6)
    ref = S.ref_phase_op(fr, conf, mode, arg, vs, nvs)
    d = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in fr.items() if v is not None}
    l1 = {"vsphse": torch.from_numpy(vs).cuda(), "nvs": torch.from_numpy(nvs).cuda()}
    if mode == 0:
        L.chunk_phasepropagate(ctx, conf, d, l1, sign=arg)
    else:
        L.chunk_phasesync_rps(ctx, conf, d, l1, layer1_based=arg)
    torch.cuda.synchronize()
    assert np.array_equal(d["phse"].cpu().numpy(), ref["phse"])
    assert np.array_equal(d["ephse"].cpu().numpy(), ref["ephse"])
    assert np.array_equal(l1["vsphse"].cpu().numpy(), ref["vsphse"])


def test_full_size_batch_round_trip(ctx):
    """BASELINE config 2 size (1024 x 400 frames): propagate forward and back on the device."""
    import torch
    import libllsm2_b200 as L
    fr, conf = S.synth_frames(4, 400, seed=42)
    rep = 256
    d = {k: torch.from_numpy(np.ascontiguousarray(np.tile(v, (rep,) + (1,) * (v.ndim - 1)))).cuda() for k, v in fr.items() if v is not None}
    from libllsm2_b200 import abi
    cb = abi.make_conf(4 * rep, conf.nfrm, conf.maxnhar, conf.maxnhar_e, conf.npsd, conf.nchannel, conf.fs, conf.thop,
                       list(conf.chanfreq)[:conf.nchannel - 1], conf.lip_radius)
    p0 = d["phse"].clone()
    L.chunk_phasepropagate(ctx, cb, d, None, sign=1)
    assert not torch.equal(p0, d["phse"])
    assert torch.equal(d["phse"][:4], d["phse"][4 * (rep - 1):])        # every copy of an utterance gets the same phases
    L.chunk_phasepropagate(ctx, cb, d, None, sign=-1)
    torch.cuda.synchronize()
    err = torch.remainder(d["phse"] - p0 + np.pi, 2 * np.pi) - np.pi
    assert float(err[..., :8].abs().max()) < 1e-3
