"""GPU parity of the analysis path (llsm_b200_analyze_l0) against the oracle's llsm_analyze, and the
analysis -> synthesis round trip."""
import numpy as np
import pytest
import support as S
from test_emu_analysis import check_analysis

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import torch
    import libllsm2_b200 as L
    assert torch.cuda.is_available()
    c = L.Context(0)
    yield c
    c.close()


def _analyze_gpu(ctx, conf, x, f0, method=1):
    import torch
    import libllsm2_b200 as L
    out = L.analyze_l0(ctx, conf, torch.from_numpy(x).cuda(), torch.from_numpy(f0).cuda(), want_residual=True,
                       options={"hm_method": method})
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}


def _case(ctx, B, F, method=1, **kw):
    fr, conf = S.synth_frames(B, F, **kw)
    y, ys, yn = S.ref_synthesize(fr, conf, seed=7)
    ref = S.ref_analyze(y, fr["f0"], conf, hm_method=method)
    o = _analyze_gpu(ctx, conf, np.ascontiguousarray(y), fr["f0"], method)
    check_analysis(o, ref, conf)
    return fr, conf, y, ref, o


def test_analysis_c2_shape(ctx):
    _case(ctx, 2, 200, seed=3, nhar=100, maxnhar=100)


def test_analysis_peak_picking_method(ctx):
    """LLSM_AOPTION_HMPP, the method test/test-layer0-anasynth.c uses by default (:34-37)."""
    _case(ctx, 2, 150, method=0, seed=13, nhar=100, maxnhar=100)
    _case(ctx, 1, 120, method=0, seed=14, thop=128 / 44100.0, nhar=200, maxnhar=400, nhar_e=5, npsd=128, f0_lo=70, f0_hi=200)


def test_analysis_c1_shape(ctx):
    _case(ctx, 1, 150, seed=4, thop=128 / 44100.0, nhar=200, maxnhar=400, nhar_e=5, npsd=128, f0_lo=70, f0_hi=200)


def test_analysis_all_unvoiced(ctx):
    fr, conf = S.synth_frames(1, 60, seed=5)
    fr["f0"][:] = 0; fr["nhar"][:] = 0; fr["enhar"][:] = 0
    y, ys, yn = S.ref_synthesize(fr, conf, seed=7)
    ref = S.ref_analyze(y, fr["f0"], conf)
    o = _analyze_gpu(ctx, conf, np.ascontiguousarray(y), fr["f0"])
    assert np.all(o["nhar"] == 0) and np.all(o["ampl"] == 0)
    assert np.abs(o["psd"] - ref["psd"]).max() < 0.05


def test_host_entry_matches_device(ctx):
    import libllsm2_b200 as L
    fr, conf = S.synth_frames(2, 60, seed=6, nhar=80, maxnhar=80)
    y, _, _ = S.ref_synthesize(fr, conf, seed=7)
    a = _analyze_gpu(ctx, conf, np.ascontiguousarray(y), fr["f0"])
    b = L.analyze_l0_host(ctx, conf, y, fr["f0"], want_residual=True)
    for k in a:
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("slices", ["1", "3"])
def test_host_pipeline_is_slicing_invariant(ctx, slices, monkeypatch):
    import libllsm2_b200 as L
    fr, conf = S.synth_frames(5, 40, seed=16, nhar=80, maxnhar=80)
    y, _, _ = S.ref_synthesize(fr, conf, seed=7)
    a = _analyze_gpu(ctx, conf, np.ascontiguousarray(y), fr["f0"])
    monkeypatch.setenv("LLSM_B200_HOST_SLICES", slices)
    b = L.analyze_l0_host(ctx, conf, y, fr["f0"], want_residual=True)
    for k in a:
        assert np.array_equal(a[k], b[k]), k


def test_analysis_synthesis_roundtrip(ctx):
    """analyse on the GPU, resynthesise on the GPU: the harmonic part must reproduce the oracle's own
    analysis->synthesis round trip (size-independent property: x ~ x_sin + x_res)."""
    import torch
    import libllsm2_b200 as L
    fr, conf, y, ref, o = _case(ctx, 2, 120, seed=9, nhar=100, maxnhar=100)
    nx = y.shape[1]
    d = {k: torch.from_numpy(np.ascontiguousarray(o[k])).cuda() for k in ("f0", "nhar", "ampl", "phse")}
    d["nfrm_utt"] = None
    xs = L.synthesize_harmonics(ctx, conf, d, nx, with_options=False).cpu().numpy()
    assert S.rms((y - xs) - o["x_res"]) < 1e-6


def test_analysis_c2_128_harmonics(ctx):
    """BASELINE configs[1] harmonic count: 128 harmonics stay below Nyquist for f0 <= 170 Hz."""
    fr, conf, y, ref, o = _case(ctx, 2, 120, seed=31, nhar=128, maxnhar=128, f0_lo=90, f0_hi=170)
    assert ref["nhar"].max() == 128


@pytest.mark.parametrize("phase_ops", [0, 3])
def test_anasynth_host_chain(ctx, phase_ops):
    """llsm_b200_anasynth_host (wave in -> analyse -> [phasesync_rps, phasepropagate] -> synthesise -> wave out, the
    chunk never leaves the device) against the same chain on the reference build (test/test-layer0-anasynth.c:40-66),
    identical white-noise templates; and against the separate host entries (bit-identical)."""
    import ctypes as C
    import libllsm2_b200 as L
    fr, conf = S.synth_frames(3, 70, seed=33, nhar=100, maxnhar=100)
    x, _, _ = S.ref_synthesize(fr, conf, seed=7)
    x = np.ascontiguousarray(x)
    white = S.ref_white_noise(conf, seed=21)
    ny = L.output_length(conf.nfrm, conf.thop, conf.fs)
    out = {k: np.zeros((conf.nutt, ny), np.float32) for k in ("y", "y_sin", "y_noise")}
    f0r = np.zeros_like(fr["f0"])
    L.anasynth_host(ctx, conf, x, fr["f0"], white=white, phase_ops=phase_ops, out=out, f0_refined=f0r)
    # reference: analyse, phase operations, synthesise
    ref = S.ref_analyze(x, fr["f0"], conf)
    assert np.abs(f0r - ref["f0"]).max() < 1e-3
    if phase_ops:
        r1 = S.ref_phase_op(ref, conf, 1, 0)
        ref["phse"], ref["ephse"] = r1["phse"], r1["ephse"]
        r2 = S.ref_phase_op(ref, conf, 0, 1)
        ref["phse"], ref["ephse"] = r2["phse"], r2["ephse"]
    ref["nfrm_utt"] = None
    y, ys, yn = S.ref_synthesize(ref, conf, seed=21)
    for got, want, name in ((out["y"], y, "y"), (out["y_sin"], ys, "y_sin"), (out["y_noise"], yn, "y_noise")):
        assert S.rms(want) > 1e-4
        assert S.rms(got - want) < 1e-4, (name, S.rms(got - want))
    # the same chain through the two separate host entries
    a = L.analyze_l0_host(ctx, conf, x, fr["f0"])
    if phase_ops:
        import torch
        d = {k: torch.from_numpy(v).cuda() for k, v in a.items()}
        L.chunk_phasesync_rps(ctx, conf, d, layer1_based=0)
        L.chunk_phasepropagate(ctx, conf, d, sign=1)
        torch.cuda.synchronize()
        a = {k: v.cpu().numpy() for k, v in d.items()}
    a["nfrm_utt"] = None
    b = L.synthesize_l0_host(ctx, conf, a, white=white)
    for k in ("y", "y_sin", "y_noise"):
        assert np.array_equal(b[k], out[k]), k
    # y only, slicing invariance
    y_only = L.anasynth_host(ctx, conf, x, fr["f0"], white=white, phase_ops=phase_ops)
    assert np.array_equal(y_only["y"], out["y"])


@pytest.mark.parametrize("slice_list", ["1,3,1", "2,1", "7"])
def test_anasynth_host_is_slicing_invariant(ctx, slice_list, monkeypatch):
    """The chained host entry cuts the batch into unequal slices (small at both ends: llsm_b200_host_slice_plan); the
    waveforms must not depend on the cut (utterances are independent; the device noise generator is indexed by the
    absolute utterance number)."""
    import libllsm2_b200 as L
    fr, conf = S.synth_frames(7, 40, seed=34, nhar=60, maxnhar=64)
    x, _, _ = S.ref_synthesize(fr, conf, seed=7)
    x = np.ascontiguousarray(x)
    monkeypatch.setenv("LLSM_B200_HOST_SLICES", "1")
    a = L.anasynth_host(ctx, conf, x, fr["f0"], seed=5)["y"].copy()
    monkeypatch.delenv("LLSM_B200_HOST_SLICES")
    monkeypatch.setenv("LLSM_B200_HOST_SLICE_LIST", slice_list)
    b = L.anasynth_host(ctx, conf, x, fr["f0"], seed=5)["y"]
    assert S.rms(a) > 1e-3
    assert np.array_equal(a, b)


def test_analysis_long_utterance_cluster_of_eight(ctx):
    """7.5 s utterance: the sub-band filter's sequence (330 k samples) is spread over a cluster of eight CTAs
    (kernels_iir_smem.cuh: carry between the CTAs through distributed shared memory); odd frame count, so the noise
    spectra's last warp works on a lone frame."""
    _case(ctx, 1, 1501, seed=21, nhar=60, maxnhar=64)


def test_analysis_low_f0_long_windows(ctx):
    """f0 50-78 Hz: every voiced frame's window exceeds the staged kernels' capacity (four periods of 80 Hz), so the
    main and the envelope pass go through the device frame list to the general kernels, and the noise spectra's Hann
    window (three periods) exceeds 2048 samples: the time-aliased path of the warp kernel."""
    _case(ctx, 2, 60, seed=22, nhar=100, maxnhar=128, f0_lo=50, f0_hi=78)


def test_analysis_48k(ctx):
    """48 kHz: other window lengths, same transform sizes (2048 / 1024: the warp kernel)."""
    _case(ctx, 1, 80, seed=23, nhar=100, maxnhar=128, fs=48000.0)


def test_analysis_16k_block_fft_path(ctx):
    """16 kHz: transform sizes 512 / 512, served by the block-FFT noise-spectra kernel; sequences short enough for a
    single-CTA filter."""
    _case(ctx, 2, 80, seed=24, nhar=40, maxnhar=40, fs=16000.0, f0_lo=100, f0_hi=200, nch=3)


def test_analysis_bench_length_cluster_of_two(ctx):
    """400 frames (2 s, BASELINE configs[1]'s utterance length): 88 200 samples per sub-band sequence = a cluster of two
    CTAs in the shared-memory filter, the configuration bench.py times."""
    _case(ctx, 1, 400, seed=25, nhar=128, maxnhar=128)
