"""GPU parity of layer-1 (pulse-by-pulse) synthesis, BASELINE config 3 harmonic count included."""
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


@pytest.mark.parametrize("kw,pattern", [(dict(seed=3, nhar=100, maxnhar=100), "blocks"),
                                        (dict(seed=4, nhar=256, maxnhar=256, f0_lo=60, f0_hi=86), "all"),
                                        (dict(seed=5, nhar=100, maxnhar=100), "none")])
def test_pbp_synthesis(ctx, kw, pattern):
    import torch
    import libllsm2_b200 as L
    B, F = 2, 120
    fr, conf = S.synth_frames(B, F, **kw)
    pbp = np.zeros((B, F), np.int32)
    if pattern == "blocks":
        pbp[:, (np.arange(F) % 40) > 20] = 1          # toggled on and off like test/test-layer1-anasynth.c:35-38
    elif pattern == "all":
        pbp[:] = 1
    ref, l1 = S.ref_synthesize_l1(fr, conf, pbp, seed=9)
    white = S.ref_white_noise(conf, seed=9)
    fr2 = dict(fr); fr2["nhar"] = None; fr2["ampl"] = None; fr2["phse"] = None
    out = L.synthesize_l1(ctx, conf, _dev(fr2), _dev(l1), pbpsyn=torch.from_numpy(pbp).cuda(),
                          white=torch.from_numpy(white).cuda())
    torch.cuda.synchronize()
    for r, k in zip(ref, ("y", "y_sin", "y_noise")):
        e = S.rms(out[k].cpu().numpy() - r)
        assert e < 1e-4, (k, e)


def test_pbp_ragged_batch_with_an_empty_utterance(ctx):
    """Utterances of 30, 0 and 11 frames in one batch against per-utterance reference runs; silence past each end."""
    import torch
    import libllsm2_b200 as L
    fr, conf, pbp, l1, refs, white, ny = S.pbp_ragged_case()
    out = L.synthesize_l1(ctx, conf, _dev(fr), _dev(l1), pbpsyn=torch.from_numpy(pbp).cuda(),
                          white=torch.from_numpy(white).cuda())
    torch.cuda.synchronize()
    S.check_pbp_ragged(tuple(out[k].cpu().numpy()[:, :ny] for k in ("y", "y_sin", "y_noise")), refs, 1e-4)
