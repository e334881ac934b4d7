"""Analysis kernels on the CPU thread emulator against the oracle's llsm_analyze (HMCZT, the
reference's default method). Input = an oracle-synthesised waveform, so the analysis sees
speech-like material."""
import ctypes as C
import numpy as np
import support as S
from libllsm2_b200 import abi


def check_analysis(o, ref, conf):
    """Parity bars for the analysis outputs (FP_TYPE=float):
       f0 / nhar / enhar identical; amplitudes 1e-6 absolute; amplitude-weighted phase 1e-6;
       residual waveform RMS < 1e-6; PSD within 0.05 dB (Kalman smoothing of float log-spectra);
       envelope means 1e-4 relative."""
    assert np.array_equal(o["nhar"], ref["nhar"])
    assert np.array_equal(o["enhar"], ref["enhar"])
    assert np.abs(o["f0"] - ref["f0"]).max() < 1e-3
    assert np.abs(o["ampl"] - ref["ampl"]).max() < 1e-6
    assert np.abs(S.phase_err(o["phse"], ref["phse"]) * ref["ampl"]).max() < 1e-6
    assert S.rms(o["x_res"] - ref["x_res"]) < 1e-6
    assert np.abs(o["psd"] - ref["psd"]).max() < 0.05
    assert np.abs(o["psdres"] - ref["psdres"]).max() < 0.1
    assert S.rms(o["psd"] - ref["psd"]) < 5e-3
    assert (np.abs(o["edc"] - ref["edc"]) / np.abs(ref["edc"])).max() < 1e-4
    assert np.abs(o["eampl"] - ref["eampl"]).max() < 1e-6 * max(1.0, float(ref["eampl"].max()))
    assert np.abs(S.phase_err(o["ephse"], ref["ephse"]) * ref["eampl"]).max() < 1e-5 * float(ref["eampl"].max())


import pytest


@pytest.mark.parametrize("method", [1, 0])     # LLSM_AOPTION_HMCZT (reference default), LLSM_AOPTION_HMPP
def test_analysis_c2_shape_small(method):
    fr, conf = S.synth_frames(1, 24, seed=3, nhar=100, maxnhar=100)
    y, ys, yn = S.ref_synthesize(fr, conf, seed=7)
    nx = y.shape[1]
    ref = S.ref_analyze(y, fr["f0"], conf, hm_method=method)
    emu = S.load_emu()
    o = S.alloc_analysis_out(conf, nx, fr["f0"])
    ao = abi.AOptions(); ao.f0_refine = 1; ao.hm_method = method; ao.rel_winsize = 4.0
    fo = S.frames_out_struct(o)
    rc = emu.emu_analyze_l0(C.byref(conf), C.byref(ao), y.ctypes.data_as(C.c_void_p), nx, nx,
                            C.byref(fo), o["x_res"].ctypes.data_as(C.c_void_p))
    assert rc == 0
    check_analysis(o, ref, conf)


def test_analysis_two_channels():
    """nchannel = 2 (one band edge): the sub-band envelope analysis with fewer channels than the default four."""
    fr, conf = S.synth_frames(1, 20, seed=4, nhar=80, maxnhar=80, nch=2, nhar_e=3)
    y, ys, yn = S.ref_synthesize(fr, conf, seed=5)
    nx = y.shape[1]
    ref = S.ref_analyze(y, fr["f0"], conf, hm_method=1)
    emu = S.load_emu()
    o = S.alloc_analysis_out(conf, nx, fr["f0"])
    ao = abi.AOptions(); ao.f0_refine = 1; ao.hm_method = 1; ao.rel_winsize = 4.0
    fo = S.frames_out_struct(o)
    rc = emu.emu_analyze_l0(C.byref(conf), C.byref(ao), y.ctypes.data_as(C.c_void_p), nx, nx,
                            C.byref(fo), o["x_res"].ctypes.data_as(C.c_void_p))
    assert rc == 0
    check_analysis(o, ref, conf)


def _emu_case(B, F, **kw):
    fr, conf = S.synth_frames(B, F, **kw)
    y, ys, yn = S.ref_synthesize(fr, conf, seed=7)
    nx = y.shape[1]
    ref = S.ref_analyze(y, fr["f0"], conf, hm_method=1)
    emu = S.load_emu()
    o = S.alloc_analysis_out(conf, nx, fr["f0"])
    ao = abi.AOptions(); ao.f0_refine = 1; ao.hm_method = 1; ao.rel_winsize = 4.0
    fo = S.frames_out_struct(o)
    rc = emu.emu_analyze_l0(C.byref(conf), C.byref(ao), y.ctypes.data_as(C.c_void_p), nx, nx,
                            C.byref(fo), o["x_res"].ctypes.data_as(C.c_void_p))
    assert rc == 0
    check_analysis(o, ref, conf)


def test_analysis_low_f0_frame_list():
    """f0 50-78 Hz: windows beyond the staged kernels' capacity (four periods of 80 Hz) -- both harmonic passes hand
    every voiced frame to the general kernels through the device frame list; the noise spectra's three-period Hann
    window exceeds 2048 samples (time-aliased path of the warp kernel)."""
    _emu_case(1, 21, seed=22, nhar=100, maxnhar=128, f0_lo=50, f0_hi=78)


def test_analysis_16k_three_channels():
    """16 kHz, three noise channels: transform sizes 512 / 512 (block-FFT noise spectra), odd frame count."""
    _emu_case(1, 25, seed=24, nhar=40, maxnhar=40, fs=16000.0, f0_lo=100, f0_hi=200, nch=3)


def test_f0_refinement_matches_the_oracle_to_the_last_bit():
    """llsm_refine_f0 (dsputils.c:72-94) with the taps m / -m of the detector's window paired (refine_f0_kernel): the
    refined track agrees with the oracle's in the last float digit on nearly every frame -- the first and last frames
    (windows cut by the utterance's ends: one side of every pair is zero padding) and high / low f0 included."""
    for kw in (dict(seed=3, nhar=100, maxnhar=100), dict(seed=8, nhar=40, maxnhar=64, f0_lo=300, f0_hi=520),
               dict(seed=9, nhar=100, maxnhar=128, f0_lo=82, f0_hi=112)):
        fr, conf = S.synth_frames(1, 30, **kw)
        y, ys, yn = S.ref_synthesize(fr, conf, seed=7)
        nx = y.shape[1]
        ref = S.ref_analyze(y, fr["f0"], conf, hm_method=1)
        emu = S.load_emu()
        o = S.alloc_analysis_out(conf, nx, fr["f0"])
        ao = abi.AOptions(); ao.f0_refine = 1; ao.hm_method = 1; ao.rel_winsize = 4.0
        fo = S.frames_out_struct(o)
        rc = emu.emu_analyze_l0(C.byref(conf), C.byref(ao), y.ctypes.data_as(C.c_void_p), nx, nx,
                                C.byref(fo), o["x_res"].ctypes.data_as(C.c_void_p))
        assert rc == 0
        v = ref["f0"] > 0
        assert v.sum() >= 10
        assert not np.array_equal(ref["f0"], fr["f0"])               # the refinement moved the track
        d = np.abs(o["f0"][v] - ref["f0"][v])
        ulp = np.spacing(ref["f0"][v])
        assert (d <= 2 * ulp).all(), (kw, float((d / ulp).max()))
        assert (d == 0).mean() >= 0.9, (kw, float((d == 0).mean()))
