"""The oracle's ciglet shim against independent implementations (numpy / scipy).

ciglet is an un-vendored dependency of the reference: the oracle restates the primitives the hot
path calls. This file pins the restatement of each transform-like primitive to an independent implementation of the
same published definition, so that an error in the shim cannot hide behind "CUDA path == oracle". FP_TYPE is float
(the parity build works in double inside the shim), so the bars are a few float roundings."""
import ctypes as C
import numpy as np
import pytest
import support as S

F32 = np.float32


def _lib():
    lib = S.load_ref()
    fp = C.POINTER(C.c_float)
    for name in ("hanning", "blackman"):
        getattr(lib, name).restype = fp
    lib.interp1.restype = fp
    lib.interp1.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    lib.filtfilt.restype = fp
    lib.filtfilt.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    lib.minphase.restype = fp
    lib.minphase.argtypes = [C.c_void_p, C.c_int]
    lib.kalmanf1d.restype = fp
    lib.kalmanf1d.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.kalmans1d.restype = fp
    lib.kalmans1d.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    lib.gensins.restype = fp
    lib.gensins.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int]
    lib.moving_avg.restype = fp
    lib.moving_avg.argtypes = [C.c_void_p, C.c_int, C.c_float]
    for name in ("fft", "ifft"):
        getattr(lib, name).argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_void_p]
    for name in ("czt", "iczt"):
        getattr(lib, name).argtypes = [C.c_void_p] * 4 + [C.c_float, C.c_int]
    lib.ddct.argtypes = [C.c_int, C.c_int, C.c_void_p]
    return lib


def _take(ptr, n):
    """Copy n floats out of a malloc'ed result (leaked on purpose: a few KB per test)."""
    return np.ctypeslib.as_array(ptr, shape=(n,)).copy()


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("n", [8, 64, 1024, 2048])
def test_fft_pair_matches_numpy(n):
    lib, rng = _lib(), np.random.default_rng(n)
    xr, xi = rng.normal(size=n).astype(F32), rng.normal(size=n).astype(F32)
    yr, yi = np.zeros(n, F32), np.zeros(n, F32)
    lib.fft(_p(xr), _p(xi), _p(yr), _p(yi), n, None)
    ref = np.fft.fft(xr.astype(np.float64) + 1j * xi)
    scale = np.abs(ref).max()
    assert np.abs(yr + 1j * yi - ref).max() < 2e-6 * scale
    zr, zi = np.zeros(n, F32), np.zeros(n, F32)
    lib.ifft(_p(yr), _p(yi), _p(zr), _p(zi), n, None)              # numpy's convention: 1/n on the inverse
    assert np.abs(zr - xr).max() < 1e-5 and np.abs(zi - xi).max() < 1e-5


@pytest.mark.parametrize("n,w0", [(100, 0.031), (441, 2 * np.pi * 120 / 44100), (7, 1.3)])
def test_czt_pair_matches_the_definition(n, w0):
    lib, rng = _lib(), np.random.default_rng(n)
    xr, xi = rng.normal(size=n).astype(F32), rng.normal(size=n).astype(F32)
    w0 = float(F32(w0))
    k = np.arange(n)
    x = xr.astype(np.float64) + 1j * xi
    yr, yi = np.zeros(n, F32), np.zeros(n, F32)
    lib.czt(_p(xr), _p(xi), _p(yr), _p(yi), C.c_float(w0), n)
    ref = np.exp(-1j * w0 * np.outer(k, k)) @ x                       # Y[k] = sum x[m] exp(-i w0 k m)
    assert np.abs(yr + 1j * yi - ref).max() < 3e-6 * np.abs(ref).max()
    lib.iczt(_p(xr), _p(xi), _p(yr), _p(yi), C.c_float(w0), n)
    ref = np.exp(1j * w0 * np.outer(k, k)) @ x / n                    # y[t] = (1 / n) sum X[k] exp(+i w0 k t)
    assert np.abs(yr + 1j * yi - ref).max() < 3e-6 * max(np.abs(ref).max(), 1.0)


@pytest.mark.parametrize("n", [64, 1024])
def test_ddct_matches_scipy(n):
    """Ooura's ddct by its published definition against scipy's DCT-II / DCT-III (unnormalised):
    isgn = -1 is half of scipy's type 2; isgn = +1 is (type 3 + a[0]) / 2."""
    from scipy.fft import dct
    lib, rng = _lib(), np.random.default_rng(n)
    a = rng.normal(size=n).astype(F32)
    c = a.copy(); lib.ddct(n, -1, _p(c))
    ref = 0.5 * dct(a.astype(np.float64), type=2)
    assert np.abs(c - ref).max() < 2e-6 * np.abs(ref).max()
    c = a.copy(); lib.ddct(n, 1, _p(c))
    ref = 0.5 * (dct(a.astype(np.float64), type=3) + a[0])
    assert np.abs(c - ref).max() < 2e-6 * np.abs(ref).max()
    # the pair as coder.c uses it (coder.c:149-152,189-192): a[0] *= 0.5; ddct(+1); a *= 2 / n inverts ddct(-1)
    c = a.copy(); lib.ddct(n, -1, _p(c)); c[0] *= 0.5; lib.ddct(n, 1, _p(c)); c *= 2.0 / n
    assert np.abs(c - a).max() < 1e-5


def test_windows_are_the_periodic_forms():
    from scipy.signal import get_window
    lib = _lib()
    for n in (441, 882, 1024):
        assert np.abs(_take(lib.hanning(n), n) - get_window("hann", n, fftbins=True)).max() < 1e-7
        assert np.abs(_take(lib.blackman(n), n) - get_window("blackman", n, fftbins=True)).max() < 1e-7


def test_interp1_is_clamped_linear_interpolation():
    lib, rng = _lib(), np.random.default_rng(1)
    xi = np.sort(rng.uniform(0, 10, 20)).astype(F32); yi = rng.normal(size=20).astype(F32)
    xq = rng.uniform(-2, 12, 200).astype(F32)
    got = _take(lib.interp1(_p(xi), _p(yi), 20, _p(xq), 200), 200)
    assert np.abs(got - np.interp(xq.astype(np.float64), xi, yi)).max() < 1e-5


def test_filtfilt_is_two_zero_state_passes():
    """Forward pass, reversal, forward pass, reversal with zero initial state and no padding: scipy's lfilter twice."""
    from scipy.signal import lfilter, cheby1
    lib, rng = _lib(), np.random.default_rng(2)
    b, a = cheby1(4, 1.0, 0.2)
    b32, a32 = b.astype(F32), a.astype(F32)
    x = rng.normal(size=3000).astype(F32)
    got = _take(lib.filtfilt(_p(b32), 5, _p(a32), 5, _p(x), 3000), 3000)
    y = lfilter(b32.astype(np.float64), a32.astype(np.float64), x.astype(np.float64))
    ref = lfilter(b32.astype(np.float64), a32.astype(np.float64), y[::-1])[::-1]
    assert np.abs(got - ref).max() < 1e-5 * np.abs(ref).max()


def test_minphase_is_the_folded_cepstrum_phase():
    lib, rng = _lib(), np.random.default_rng(3)
    nfft, ns = 512, 257
    # a known minimum-phase system: H(z) = 1 / (1 - 0.9 z^-1 + 0.5 z^-2), all poles inside the unit circle
    w = np.pi * np.arange(ns) / (ns - 1)
    H = 1.0 / (1 - 0.9 * np.exp(-1j * w) + 0.5 * np.exp(-2j * w))
    lm = np.log(np.abs(H)).astype(F32)
    got = _take(lib.minphase(_p(lm), nfft), ns)
    assert np.abs(got - np.angle(H)).max() < 2e-4                     # cepstral aliasing at nfft = 512


def test_kalman_filter_and_smoother_match_a_plain_implementation():
    lib, rng = _lib(), np.random.default_rng(4)
    n = 300
    z = np.cumsum(rng.normal(0, 0.3, n)).astype(F32) + rng.normal(0, 1.0, n).astype(F32)
    Q = rng.uniform(0.01, 0.2, n).astype(F32); R = rng.uniform(0.5, 2.0, n).astype(F32)
    P = np.zeros(n, F32)
    y = _take(lib.kalmanf1d(_p(z), _p(Q), _p(R), n, _p(P), None), n)
    xs, Ps = np.zeros(n), np.zeros(n)
    xs[0], Ps[0] = z[0], R[0]                                         # the shim's documented initial state
    for t in range(1, n):
        Pp = Ps[t - 1] + Q[t]
        K = Pp / (Pp + R[t])
        xs[t] = xs[t - 1] + K * (z[t] - xs[t - 1]); Ps[t] = (1 - K) * Pp
    assert np.abs(y - xs).max() < 1e-5 and np.abs(P - Ps).max() < 1e-6
    s = _take(lib.kalmans1d(_p(y), _p(P), _p(Q), n), n)
    sm = np.zeros(n); sm[-1] = y[-1]
    for t in range(n - 2, -1, -1):                                    # Rauch-Tung-Striebel, random-walk model
        Pp = float(P[t]) + Q[t + 1]
        sm[t] = y[t] + float(P[t]) / Pp * (sm[t + 1] - y[t])
    assert np.abs(s - sm).max() < 1e-5


def test_gensins_time_origin_is_the_centre():
    lib = _lib()
    f = np.array([100.0, 250.0], F32); a = np.array([0.5, 0.25], F32); ph = np.array([0.3, -1.0], F32)
    n, fs = 400, 8000.0
    got = _take(lib.gensins(_p(f), _p(a), _p(ph), 2, C.c_float(fs), n), n)
    t = np.arange(n) - n // 2
    ref = sum(a[k] * np.cos(2 * np.pi * float(f[k]) / fs * t + float(ph[k])) for k in range(2))
    assert np.abs(got - ref).max() < 1e-6


def test_moving_avg_half_order():
    lib, rng = _lib(), np.random.default_rng(5)
    x = rng.normal(size=50).astype(F32)
    got = _take(lib.moving_avg(_p(x), 50, C.c_float(3.0)), 50)
    ref = np.array([x[max(0, i - 3):min(50, i + 4)].astype(np.float64).mean() for i in range(50)])
    assert np.abs(got - ref).max() < 1e-6


class _LF(C.Structure):
    _fields_ = [("T0", C.c_float), ("te", C.c_float), ("tp", C.c_float), ("ta", C.c_float), ("Ee", C.c_float)]


@pytest.mark.parametrize("rd", [0.5, 1.0, 1.8, 2.6])
def test_lf_spectrum_matches_numerical_transform_of_the_waveform(rd):
    """The closed-form LF spectrum of the shim against a brute-force Fourier integral of the textbook LF
    flow-derivative waveform (Fant, Liljencrants & Lin 1985), its two implicit parameters solved independently with
    scipy: E(t) = E0 exp(alpha t) sin(wg t) up to te, then -(Ee / (eps ta)) (exp(-eps (t - te)) - exp(-eps (tc - te)))."""
    from scipy.optimize import brentq
    lib = _lib()
    lib.lfmodel_from_rd.restype = _LF
    lib.lfmodel_from_rd.argtypes = [C.c_float, C.c_float, C.c_float]
    lib.lfmodel_spectrum.restype = C.POINTER(C.c_float)
    lib.lfmodel_spectrum.argtypes = [_LF, C.c_void_p, C.c_int, C.c_void_p]
    T0 = 1.0 / 120.0
    m = lib.lfmodel_from_rd(rd, T0, 1.0)
    te, tp, ta, Ee = float(m.te), float(m.tp), float(m.ta), float(m.Ee)          # relative to T0, tc = 1
    wg = np.pi / tp
    eps = brentq(lambda e: e * ta - 1.0 + np.exp(-e * (1.0 - te)), 1e-3, 1e6) if ta < 1.0 - te else None
    assert eps is not None
    N = 1 << 18
    t = (np.arange(N) + 0.5) / N                                                 # midpoint rule on [0, 1)

    def waveform(alpha):
        E0 = -Ee / (np.exp(alpha * te) * np.sin(wg * te))
        op = E0 * np.exp(alpha * t) * np.sin(wg * t)
        rp = -(Ee / (eps * ta)) * (np.exp(-eps * (t - te)) - np.exp(-eps * (1.0 - te)))
        return np.where(t <= te, op, rp)

    alpha = brentq(lambda a: waveform(a).sum() / N, -200.0, 400.0, xtol=1e-10)   # zero net flow over the period
    e = waveform(alpha)
    freq = (np.arange(1, 41) * 120.0).astype(F32)
    ph = np.zeros(40, F32)
    mag = _take(lib.lfmodel_spectrum(m, _p(freq), 40, _p(ph)), 40)
    w = 2 * np.pi * freq.astype(np.float64) * T0
    ref = T0 * (e[None, :] * np.exp(-1j * w[:, None] * t[None, :])).sum(1) / N
    assert np.abs(mag - np.abs(ref)).max() < 2e-4 * np.abs(ref).max()
    big = np.abs(ref) > 1e-3 * np.abs(ref).max()
    assert np.abs(S.phase_err(ph, np.angle(ref)))[big].max() < 2e-3


def test_if_detector_refines_f0_of_a_harmonic_signal():
    """llsm_refine_f0 (dsputils.c:72-94) through the shim's instantaneous-frequency detector: a steady harmonic signal
    at 123.4 Hz analysed with a 3 % detuned f0 track (3.4 Hz off) comes back within 0.4 Hz. The estimator is exact for
    one complex exponential; what is left here is the neighbouring harmonics leaking through the four-period Hann
    window (about -45 dB at one harmonic spacing, times the 123 Hz distance), a property of the method."""
    lib = S.load_ref()
    lib.llsm_refine_f0.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_int, C.c_float]
    fs, thop, nfrm, true = 16000.0, 0.005, 60, 123.4
    nx = int(fs * thop * (nfrm + 2))
    t = np.arange(nx) / fs
    x = sum(np.cos(2 * np.pi * true * k * t + 0.3 * k) / k for k in range(1, 6)).astype(F32)
    f0 = np.full(nfrm, 120.0, F32)
    lib.llsm_refine_f0(_p(x), nx, C.c_float(fs), _p(f0), nfrm, C.c_float(thop))
    assert np.abs(f0[8:-8] - true).max() < 0.4, np.abs(f0[8:-8] - true).max()
    assert np.abs(f0[8:-8] - true).max() < 0.12 * 3.4


@pytest.mark.parametrize("window,nwin", [("hanning", 1024), ("blackman", 882), ("hanning", 3000)])
def test_stft_reads_amplitude_and_phase_of_a_sinusoid_at_the_frame_centre(window, nwin):
    """Known answer: A cos(2 pi f t + phi) with f on a bin gives |X| 2 / sum(w) = A and arg X = the phase at the frame
    centre (zero-phase window placement), also for a window longer than the transform (time aliasing)."""
    lib = S.load_ref()
    nfft, fs, k0, A, phi = 2048, 16000.0, 37, 0.7, 0.9
    f = k0 * fs / nfft
    nx = 8000
    t = np.arange(nx)
    x = (A * np.cos(2 * np.pi * f / fs * t + phi)).astype(F32)
    centers = np.array([2500, 4001, 5555], np.int32)
    nw = np.full(3, nwin, np.int32)
    ns = nfft // 2 + 1
    mag, ph = np.zeros((3, ns), F32), np.zeros((3, ns), F32)
    rows = lambda a: (C.c_void_p * 3)(*[a[i].ctypes.data for i in range(3)])  # noqa: E731
    norm = np.zeros(3, F32)
    lib.cig_stft_forward.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_int,
                                     C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.cig_stft_forward(_p(x), nx, _p(centers), _p(nw), 3, nfft, window.encode(), 0, 0, _p(norm), None, rows(mag), rows(ph))
    for i, c in enumerate(centers):
        assert abs(mag[i, k0] * 2.0 / norm[i] - A) < 2e-4
        want = 2 * np.pi * f / fs * float(c) + phi
        assert abs(S.phase_err(ph[i, k0], want)) < 2e-4
        assert int(np.argmax(mag[i])) == k0


def test_spec2env_removes_ripple_at_the_pitch_period_and_keeps_the_envelope():
    """The cepstral lifter is a sinc with its first zero at quefrency 1 / f0: harmonic ripple of period f0 in the log
    spectrum vanishes, a slow envelope passes (times the lifter's gain at that quefrency)."""
    lib = S.load_ref()
    lib.cig_spec2env.restype = C.POINTER(C.c_float)
    lib.cig_spec2env.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_void_p]
    nfft, ns = 2048, 1025
    f0 = 1.0 / 64.0                                   # cycles per sample: ripple period 32 bins, quefrency 64
    k = np.arange(ns)
    slow_q = 4                                        # envelope component at quefrency 4
    logS = -3.0 + 0.8 * np.cos(2 * np.pi * slow_q * k / nfft) + 0.5 * np.cos(2 * np.pi * k / 32.0)
    Sp = np.exp(logS).astype(F32)
    env = _take(lib.cig_spec2env(_p(Sp), nfft, C.c_float(f0), 0, None), ns)
    x = f0 * slow_q
    gain = np.sin(np.pi * x) / (np.pi * x) * (1.18 - 0.18 * np.cos(2 * np.pi * x))
    want = -3.0 + 0.8 * gain * np.cos(2 * np.pi * slow_q * k / nfft)
    assert np.abs(env - want).max() < 2e-5


def test_interp1u_has_an_exclusive_end_point_and_blank_filling_holds_the_ends():
    lib = S.load_ref()
    fp = C.POINTER(C.c_float)
    lib.interp1u.restype = fp
    lib.interp1u.argtypes = [C.c_float, C.c_float, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    yi = np.array([0.0, 1.0, 4.0, 9.0], F32)                         # knots at 0, 1, 2, 3 when x1 = 4 (last knot + one step)
    xq = np.array([-1.0, 0.5, 2.5, 3.0, 3.9], F32)
    got = _take(lib.interp1u(0.0, 4.0, _p(yi), 4, _p(xq), 5), 5)
    assert np.allclose(got, [0.0, 0.5, 6.5, 9.0, 9.0], atol=1e-6)
    lib.interp_in_blank.restype = fp
    lib.interp_in_blank.argtypes = [C.c_void_p, C.c_int, C.c_float]
    x = np.array([0, 0, 2, 0, 0, 5, 0], F32)
    got = _take(lib.interp_in_blank(_p(x), 7, 0.0), 7)
    assert np.allclose(got, [2, 2, 2, 3, 4, 5, 5], atol=1e-6)
