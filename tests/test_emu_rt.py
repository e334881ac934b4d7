"""Streaming synthesis kernels on the CPU thread emulator against llsm_rtsynth_buffer_* of the
oracle build (llsmrt.c). Bar: RMS < 1e-4; observed ~1e-8."""
import ctypes as C
import numpy as np
import pytest
import support as S
from libllsm2_b200 import abi


def _run(B, F, seed=5, mutate=None, use_iczt=1, clear_at=-1, **kw):
    fr, conf = S.synth_frames(B, F, **kw)
    if mutate:
        mutate(fr)
    P, A, lat = S.ref_rtsynth(fr, conf, seed=seed, use_iczt=use_iczt, clear_at=clear_at)
    white = S.ref_rt_white(conf, seed=seed)
    so = abi.default_soptions(white.ctypes.data, 0)
    so.use_iczt = use_iczt
    emu = S.load_emu()
    cap = P.shape[1] + 64
    op = np.full((B, cap), np.nan, np.float32); oap = op.copy()
    n = C.c_int(0); l = C.c_int(0)
    f = S.frames_struct(fr)
    rc = emu.emu_rtsynth(C.byref(conf), C.byref(f), C.byref(so), op.ctypes.data_as(C.c_void_p),
                         oap.ctypes.data_as(C.c_void_p), cap, C.byref(n), C.byref(l), clear_at)
    assert rc == 0
    assert n.value == P.shape[1] and l.value == lat
    return (P, A), (op[:, :n.value], oap[:, :n.value])


def _check(ref, got, tol=1e-6):
    for r, g, name in zip(ref, got, ("p", "ap")):
        assert np.isfinite(g).all(), name
        assert S.rms(g - r) < tol, (name, S.rms(g - r))
        assert S.rms(r) > 1e-4, name


def test_rt_c2_shape():
    ref, got = _run(2, 30)
    _check(ref, got)


def test_rt_noninteger_hop_and_unvoiced_gaps():
    def mut(fr):
        fr["f0"][:, 8:14] = 0; fr["nhar"][:, 8:14] = 0; fr["enhar"][:, 8:14] = 0
    ref, got = _run(1, 40, thop=100.5 / 44100.0, mutate=mut)
    _check(ref, got)


def test_rt_many_harmonics_iczt_switch():
    ref, got = _run(1, 16, thop=128 / 44100.0, nhar=400, maxnhar=400, nhar_e=5, npsd=128,
                    f0_lo=50, f0_hi=90)
    _check(ref, got)
    ref, got = _run(1, 16, thop=128 / 44100.0, nhar=400, maxnhar=400, nhar_e=5, npsd=128,
                    f0_lo=50, f0_hi=90, use_iczt=0)
    _check(ref, got)


def test_rt_clear_midstream():
    """llsm_rtsynth_buffer_clear keeps the modulation rings and the previous noise model (llsmrt.c:578-602)."""
    ref, got = _run(1, 24, thop=100.5 / 44100.0, clear_at=11)
    _check(ref, got)


def _run_l1(B, F, pbp, seed=6, remove_hm=1, host_tracker=0, block=5, **kw):
    fr, conf = S.synth_frames(B, F, seed=3, nhar=100, maxnhar=100, **kw)
    P, A, lat = S.ref_rtsynth(fr, conf, seed=seed, use_l1=1, pbpsyn=pbp, remove_hm=remove_hm)
    _, l1 = S.ref_synthesize_l1(fr, conf, pbp, seed=9)
    white = S.ref_rt_white(conf, seed=seed)
    so = abi.default_soptions(white.ctypes.data, 0)
    fr2 = dict(fr)
    if remove_hm:
        fr2["nhar"] = None; fr2["ampl"] = None; fr2["phse"] = None
    f = S.frames_struct(fr2)
    s = abi.Layer1(); s.rd = l1["rd"].ctypes.data; s.vtmagn = l1["vtmagn"].ctypes.data
    s.vsphse = l1["vsphse"].ctypes.data; s.nvs = l1["nvs"].ctypes.data; s.nspec = l1["vtmagn"].shape[-1]
    cap = P.shape[1] + 64
    op = np.full((B, cap), np.nan, np.float32); oap = op.copy()
    n = C.c_int(0); l = C.c_int(0)
    emu = S.load_emu()
    rc = emu.emu_rtsynth_l1(C.byref(conf), C.byref(f), C.byref(s), pbp.ctypes.data_as(C.c_void_p), C.byref(so),
                            op.ctypes.data_as(C.c_void_p), oap.ctypes.data_as(C.c_void_p), cap, C.byref(n), C.byref(l),
                            host_tracker, block)
    assert rc == 0
    assert n.value == P.shape[1] and l.value == lat
    return (P, A), (op[:, :n.value], oap[:, :n.value])


@pytest.mark.parametrize("mode,host_tracker", [("switching", 0), ("switching", 1), ("all", 0), ("none", 0)])
def test_rt_layer1_pulse_by_pulse(mode, host_tracker):
    """use_l1 streaming: onset two periods early, windowed hand-over from the pulse buffer, trapezoid
    catch-up at termination (llsmrt.c:305-419); HM derived from layer 1 when absent."""
    B, F = (2 if mode == "switching" else 1), 48            # one utterance is enough for the uniform modes
    pbp = np.zeros((B, F), np.int32)
    if mode == "switching":
        pbp[0, 10:22] = 1; pbp[0, 30:41] = 1; pbp[1, 5:40] = 1
    elif mode == "all":
        pbp[:] = 1
    ref, got = _run_l1(B, F, pbp, host_tracker=host_tracker)
    _check(ref, got, tol=1e-5)


def test_rt_layer1_with_harmonic_model_present():
    B, F = 1, 30
    pbp = np.zeros((B, F), np.int32); pbp[0, 12:20] = 1
    ref, got = _run_l1(B, F, pbp, remove_hm=0, block=30)
    _check(ref, got, tol=1e-5)


@pytest.mark.parametrize("nch,nhar_e", [(2, 3), (1, 2)])
def test_rt_other_channel_counts(nch, nhar_e):
    ref, got = _run(1, 20, nch=nch, nhar_e=nhar_e)
    _check(ref, got)
