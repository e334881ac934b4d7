"""GPU parity of the streaming synthesizer (llsm_b200_rt_* through the C ABI) against
llsm_rtsynth_buffer_* of the oracle build. Bar: RMS < 1e-4."""
import numpy as np
import pytest
import support as S

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def ctx():
    import torch
    import libllsm2_b200 as L
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    c = L.Context(0)
    yield c
    c.close()


def _slice(fr, lo, hi):
    return {k: (np.ascontiguousarray(v[:, lo:hi]) if v is not None and k != "nfrm_utt" else None) for k, v in fr.items()}


def _stream(ctx, fr, conf, white, block, device, clear_at=-1, options=None):
    import torch
    import libllsm2_b200 as L
    rt = L.RtSynth(ctx, conf, white=torch.from_numpy(white).cuda() if device else white, options=options)
    lat = rt.latency
    ps, aps = [], []
    F = conf.nfrm
    i = 0
    while i < F:
        k = min(block, F - i)
        if clear_at >= 0 and i < clear_at < i + k:
            k = clear_at - i
        if i == clear_at:
            rt.clear()
        part = _slice(fr, i, i + k)
        if device:
            part = {kk: (torch.from_numpy(v).cuda() if v is not None else None) for kk, v in part.items()}
        p, ap = rt.feed(part, k)
        if device:
            torch.cuda.synchronize(); p, ap = p.cpu().numpy(), ap.cpu().numpy()
        ps.append(p.copy()); aps.append(ap.copy())
        i += k
    rt.close()
    return np.concatenate(ps, 1), np.concatenate(aps, 1), lat


def _check(ref, got, tol=TOL):
    P, A, lat = ref
    p, ap, l = got
    assert l == lat and p.shape == P.shape
    e = (S.rms(p - P), S.rms(ap - A))
    assert np.isfinite(p).all() and np.isfinite(ap).all()
    assert e[0] < tol and e[1] < tol, e
    return e


@pytest.mark.parametrize("device,block", [(True, 1), (False, 7), (True, 30)])
def test_rt_c2_shape(ctx, device, block):
    fr, conf = S.synth_frames(3, 30)
    ref = S.ref_rtsynth(fr, conf, seed=5)
    white = S.ref_rt_white(conf, seed=5)
    e = _check(ref, _stream(ctx, fr, conf, white, block, device))
    assert max(e) < 1e-6


def test_rt_noninteger_hop_unvoiced_gap_and_clear(ctx):
    fr, conf = S.synth_frames(2, 40, thop=100.5 / 44100.0)
    fr["f0"][:, 8:14] = 0; fr["nhar"][:, 8:14] = 0; fr["enhar"][:, 8:14] = 0
    white = S.ref_rt_white(conf, seed=2)
    _check(S.ref_rtsynth(fr, conf, seed=2), _stream(ctx, fr, conf, white, 5, True))
    _check(S.ref_rtsynth(fr, conf, seed=2, clear_at=17), _stream(ctx, fr, conf, white, 5, False, clear_at=17))


def test_rt_many_harmonics_iczt_switch(ctx):
    fr, conf = S.synth_frames(1, 16, thop=128 / 44100.0, nhar=400, maxnhar=400, nhar_e=5, npsd=128, f0_lo=50, f0_hi=90)
    white = S.ref_rt_white(conf, seed=3)
    _check(S.ref_rtsynth(fr, conf, seed=3), _stream(ctx, fr, conf, white, 4, True))
    _check(S.ref_rtsynth(fr, conf, seed=3, use_iczt=0), _stream(ctx, fr, conf, white, 4, True, options={"use_iczt": 0}))


def test_rt_device_generator_statistics(ctx):
    """No host templates: Philox N(0,1) on the device; the aperiodic part keeps the level of the reference's."""
    import libllsm2_b200 as L
    fr, conf = S.synth_frames(2, 60)
    P, A, lat = S.ref_rtsynth(fr, conf, seed=4)
    rt = L.RtSynth(ctx, conf, seed=1234)
    p, ap = rt.feed(_slice(fr, 0, 60), 60)
    assert S.rms(p - P) < 1e-6
    n0 = 20 * 441
    ratio = S.rms(ap[:, n0:]) / S.rms(A[:, n0:])
    assert 0.8 < ratio < 1.25, ratio


def _l1_case(ctx, B, F, pbp, device, host_tracker=False, block=6, remove_hm=True, seed=6):
    import torch
    import libllsm2_b200 as L
    fr, conf = S.synth_frames(B, F, seed=3, nhar=100, maxnhar=100)
    ref = S.ref_rtsynth(fr, conf, seed=seed, use_l1=1, pbpsyn=pbp, remove_hm=1 if remove_hm else 0)
    _, l1 = S.ref_synthesize_l1(fr, conf, pbp, seed=9)
    white = S.ref_rt_white(conf, seed=seed)
    rt = L.RtSynth(ctx, conf, white=torch.from_numpy(white).cuda() if device else white, nspec=l1["vtmagn"].shape[-1],
                   host_tracker=host_tracker)
    lat = rt.latency
    ps, aps = [], []
    for i in range(0, F, block):
        k = min(block, F - i)
        part = _slice(fr, i, i + k)
        if remove_hm:
            part["nhar"] = part["ampl"] = part["phse"] = None
        lp = {kk: np.ascontiguousarray(v[:, i:i + k]) for kk, v in l1.items()}
        pp = np.ascontiguousarray(pbp[:, i:i + k])
        if device:
            part = {kk: (torch.from_numpy(v).cuda() if v is not None else None) for kk, v in part.items()}
            lp = {kk: torch.from_numpy(v).cuda() for kk, v in lp.items()}
            pp = torch.from_numpy(pp).cuda()
        p, ap = rt.feed(part, k, layer1=lp, pbpsyn=pp)
        if device:
            torch.cuda.synchronize(); p, ap = p.cpu().numpy(), ap.cpu().numpy()
        ps.append(p.copy()); aps.append(ap.copy())
    rt.close()
    return ref, (np.concatenate(ps, 1), np.concatenate(aps, 1), lat)


@pytest.mark.parametrize("device,host_tracker", [(True, False), (False, False), (False, True)])
def test_rt_layer1_pulse_by_pulse(ctx, device, host_tracker):
    B, F = 3, 48
    pbp = np.zeros((B, F), np.int32)
    pbp[0, 10:22] = 1; pbp[0, 30:41] = 1; pbp[1, 5:40] = 1
    ref, got = _l1_case(ctx, B, F, pbp, device, host_tracker)
    e = _check(ref, got)
    assert max(e) < 1e-5, e


def test_rt_layer1_all_pbp_with_stored_hm(ctx):
    B, F = 1, 30
    pbp = np.ones((B, F), np.int32)
    ref, got = _l1_case(ctx, B, F, pbp, True, remove_hm=False, block=30)
    assert max(_check(ref, got)) < 1e-5
