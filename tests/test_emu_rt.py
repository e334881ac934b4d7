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
