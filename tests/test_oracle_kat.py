"""Pin the oracle: the reference's own dependency-free known-answer tests, re-run against the
build of the unmodified reference sources + ciglet shim (oracle/_ref). Tolerances are the
reference's (test/test-dsputils.c:86-127 chirp KAT, :135-166 glottal-fit KAT)."""
import ctypes as C
import math
import numpy as np
import pytest
import support as S

fp = C.POINTER(C.c_float)


def _chirp(method):
    L = S.load_ref()
    nx, fs, thop = 100000, 20000.0, 0.005
    nfrm = int(math.floor(np.float32(nx) / np.float32(fs) / np.float32(thop)))
    center = np.round(np.arange(nfrm) * np.float32(thop) * np.float32(fs)).astype(int)
    rate = center.astype(np.float32) / nx
    f0 = (100 + 100 * rate).astype(np.float32)
    i = np.arange(nx)
    ph = np.cumsum((100 + 100 * i / nx) / fs * 2 * 3.1415927)
    x = ((i / nx) * np.sin(ph) + 0.5 * np.sin(2 * ph) + 0.25 * np.sin(3 * ph)).astype(np.float32)
    nhar = (C.c_int * nfrm)(); ampl = (fp * nfrm)(); phse = (fp * nfrm)()
    L.llsm_harmonic_analysis(x.ctypes.data_as(fp), nx, C.c_float(fs), f0.ctypes.data_as(fp), nfrm,
                             C.c_float(thop), C.c_float(4.0), 3, method, nhar, ampl, phse)
    A = np.array([[ampl[t][k] for k in range(3)] for t in range(nfrm)])
    P = np.array([phse[t][0] for t in range(nfrm)])
    return nfrm, rate, f0, thop, A, P


@pytest.mark.parametrize("method", [0, 1])   # LLSM_AOPTION_HMPP, LLSM_AOPTION_HMCZT
def test_chirp_harmonic_analysis(method):
    nfrm, rate, f0, thop, A, P = _chirp(method)
    s = slice(5, nfrm - 5)
    for k, truth in enumerate([rate, 0.5, 0.25]):
        e = np.zeros(nfrm); e[s] = (A[:, k] - truth)[s]
        assert abs(e.mean()) < 0.01 and e.std() < 0.01
    pe = np.zeros(nfrm - 1)
    for i in range(5, nfrm - 5):
        d = P[i] - (P[i - 1] + f0[i] * 2 * 3.1415927 * thop)
        pe[i - 1] = (d + math.pi) % (2 * math.pi) - math.pi
    assert abs(pe.mean()) < 0.1 and pe.std() < 0.1


def test_glottal_fitting():
    L = S.load_ref()

    class LF(C.Structure):
        _fields_ = [(n, C.c_float) for n in ("T0", "te", "tp", "ta", "Ee")]
    L.llsm_create_cached_glottal_model.restype = C.c_void_p
    L.llsm_spectral_glottal_fitting.restype = C.c_float
    L.linspace.restype = fp
    L.lfmodel_from_rd.restype = LF
    L.lfmodel_spectrum.restype = fp
    pl = L.linspace(C.c_float(0.02), C.c_float(3.0), 64)
    cgm = C.c_void_p(L.llsm_create_cached_glottal_model(pl, 64, 20))
    freq = (200.0 * (np.arange(20) + 1)).astype(np.float32)
    for k in range(0, 500, 7):
        rd = 0.3 + (2.5 - 0.3) / 500 * k
        lf = L.lfmodel_from_rd(C.c_float(rd), C.c_float(1 / 200.0), C.c_float(1.0))
        a = L.lfmodel_spectrum(lf, freq.ctypes.data_as(fp), 20, None)
        arr = np.array([a[j] / (j + 1) * 3.7 for j in range(20)], np.float32)
        est = L.llsm_spectral_glottal_fitting(arr.ctypes.data_as(fp), 20, cgm)
        assert abs(est - rd) < 0.02


def test_iczt_matches_sinusoid_bank():
    """test/test-harmonic.c:40-48: ICZT and recurrent generators agree (100 harmonics, 1024 samples)."""
    L = S.load_ref()
    L.llsm_synthesize_harmonic_frame.restype = fp
    L.llsm_synthesize_harmonic_frame_iczt.restype = fp
    rng = np.random.default_rng(0)
    a = rng.uniform(0, 1, 100).astype(np.float32); p = rng.uniform(-3, 3, 100).astype(np.float32)
    y0 = L.llsm_synthesize_harmonic_frame(a.ctypes.data_as(fp), p.ctypes.data_as(fp), 100, C.c_float(0.004), 1024)
    y1 = L.llsm_synthesize_harmonic_frame_iczt(a.ctypes.data_as(fp), p.ctypes.data_as(fp), 100, C.c_float(0.004), 1024)
    y0 = np.array([y0[i] for i in range(1024)]); y1 = np.array([y1[i] for i in range(1024)])
    snr = 10 * np.log10((y0 ** 2).sum() / ((y0 - y1) ** 2).sum())
    assert snr > 80


def test_edgecase_all_unvoiced_noninteger_hop():
    """test/test-layer0-edgecase.c: all-unvoiced F0, hop of 100.5 samples: y_sin == 0, no crash."""
    fr, conf = S.synth_frames(1, 50, thop=100.5 / 44100.0, seed=1)
    fr["f0"][:] = 0; fr["nhar"][:] = 0; fr["enhar"][:] = 0
    y, ys, yn = S.ref_synthesize(fr, conf)
    assert np.all(ys == 0) and np.isfinite(yn).all()
