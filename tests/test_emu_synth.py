"""Kernel logic on the CPU: the CUDA kernel sources, compiled for the thread emulator
(tests/emu), against the oracle build of the reference. Bar: RMS error < 1e-4 (FP_TYPE=float);
observed ~1e-8. The GPU parity tests proper are tests/test_gpu_*.py."""
import ctypes as C
import numpy as np
import pytest
import support as S
from libllsm2_b200 import abi


def _run(B, F, seed=7, nfrm_utt=None, mutate=None, chanfreq=None, **kw):
    fr, conf = S.synth_frames(B, F, **kw)
    for i, f in enumerate(chanfreq or ()):
        conf.chanfreq[i] = f
    if mutate:
        mutate(fr)
    if nfrm_utt is not None:
        fr["nfrm_utt"] = np.asarray(nfrm_utt, np.int32)
    y, ys, yn = S.ref_synthesize(fr, conf, seed=seed)
    emu = S.load_emu()
    ny = ys.shape[1]
    white = S.ref_white_noise(conf, seed=seed, nfrm_utt=fr["nfrm_utt"])
    so = abi.default_soptions(white.ctypes.data, 0)
    oy = np.full((B, ny), np.nan, np.float32); oys = oy.copy(); oyn = oy.copy()
    out = abi.Output(); out.y = oy.ctypes.data; out.y_sin = oys.ctypes.data
    out.y_noise = oyn.ctypes.data; out.stride = ny
    f = S.frames_struct(fr)
    assert emu.emu_synthesize_l0(C.byref(conf), C.byref(f), C.byref(so), C.byref(out)) == 0
    return (y, ys, yn), (oy, oys, oyn)


def _check(ref, got, tol=1e-4):
    for r, g, name in zip(ref, got, ("y", "y_sin", "y_noise")):
        assert np.isfinite(g).all(), name
        assert S.rms(g - r) < tol, (name, S.rms(g - r))


def test_c2_shape_small():
    ref, got = _run(2, 24, seed=3)
    _check(ref, got, 1e-6)


def test_c1_shape_small():
    ref, got = _run(1, 40, thop=128 / 44100.0, nhar=400, maxnhar=400, nhar_e=5, npsd=128,
                    f0_lo=70, f0_hi=200)
    _check(ref, got, 1e-6)


def test_all_unvoiced_noninteger_hop():
    def mut(fr):
        fr["f0"][:] = 0; fr["nhar"][:] = 0; fr["enhar"][:] = 0
    ref, got = _run(1, 30, thop=100.5 / 44100.0, mutate=mut)
    assert np.all(got[1] == 0)
    _check(ref, got, 1e-6)


def test_ragged_batch():
    ref, got = _run(3, 20, nfrm_utt=[20, 7, 13])
    _check(ref, got, 1e-6)


def test_silent_noise_frames_are_skipped():
    def mut(fr):
        fr["psd"][:, 5:12, :] = -120.0
    ref, got = _run(1, 24, mutate=mut)
    _check(ref, got, 1e-6)


@pytest.mark.parametrize("nch,nhar_e", [(2, 3), (3, 3), (1, 2)])
def test_other_channel_counts(nch, nhar_e):
    """Fewer than four noise channels: the excitation kernel's generic coefficient layout (padded frame slots)."""
    ref, got = _run(1, 20, seed=5, nch=nch, nhar_e=nhar_e)
    _check(ref, got, 1e-6)


def test_six_channels():
    """More than four noise channels: the eight-channel instance of the excitation kernel."""
    ref, got = _run(1, 20, seed=8, nch=6, nhar_e=3, chanfreq=(1000.0, 2000.0, 4000.0, 8000.0, 12000.0))
    _check(ref, got, 1e-6)


def test_empty_utterance_in_a_ragged_batch():
    ref, got = _run(3, 20, nfrm_utt=[20, 0, 13])
    _check(ref, got, 1e-6)
    assert all(np.all(g[1] == 0) for g in got)
