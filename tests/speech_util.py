"""The reference's audio fixtures and its statistical acceptance metrics, for both libraries.

* fixtures(): tests/golden/speech_fixtures.npz (tools/gen_speech_fixtures.py): test/arctic_a0001.wav and
  test/are-you-ready.wav as float32 in [-1, 1) (ciglet's wavread scaling) with the harness F0 track.
* anasynth(): the call sequence of test/test-layer0-anasynth.c:29-63 through the llsm.h symbols of a library
  (the reference build or the drop-in): llsm_analyze -> llsm_synthesize -> llsm_chunk_phasesync_rps(0) +
  llsm_chunk_phasepropagate(+1) -> llsm_synthesize.
* verify_data_distribution / verify_spectral_distribution: numpy restatement of test/verify-utils.h:8-171
  (Perez-Cruz KL estimate on sorted samples, dithered; STFT correlation) -- the reference's own pass / fail bars.
"""
import ctypes as C
import os

import numpy as np

import compat_util as U

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libc = C.CDLL(None)
_fx = None


def fixtures():
    global _fx
    if _fx is None:
        d = np.load(os.path.join(ROOT, "tests", "golden", "speech_fixtures.npz"))
        _fx = {}
        for key in ("arctic", "ready"):
            _fx[key] = dict(x=(d[key + "_x"].astype(np.float32) / np.float32(32768.0)), fs=float(d[key + "_fs"]),
                            f0=d[key + "_f0"].astype(np.float32), nhop=int(d["nhop"]))
    return _fx


def chunk_flat(L, ck, nfrm, maxnhar, maxnhar_e, npsd, nchannel):
    """Every layer-0 member of a chunk as flat arrays (ragged rows zero-padded)."""
    o = dict(f0=np.zeros(nfrm, np.float32), nhar=np.zeros(nfrm, np.int32), ampl=np.zeros((nfrm, maxnhar), np.float32),
             phse=np.zeros((nfrm, maxnhar), np.float32), psd=np.zeros((nfrm, npsd), np.float32),
             psdres=np.zeros((nfrm, npsd), np.float32), edc=np.zeros((nfrm, nchannel), np.float32),
             enhar=np.zeros((nfrm, nchannel), np.int32), eampl=np.zeros((nfrm, nchannel, maxnhar_e), np.float32),
             ephse=np.zeros((nfrm, nchannel, maxnhar_e), np.float32))
    for i in range(nfrm):
        fr = ck.contents.frames[i]
        o["f0"][i] = C.cast(L.llsm_container_get(fr, U.F0), U.fp)[0]
        hm = L.llsm_container_get(fr, U.HMI)
        if hm:
            h = C.cast(hm, C.POINTER(U.HM)).contents
            o["nhar"][i] = h.nhar
            if h.nhar:
                o["ampl"][i, :h.nhar] = np.ctypeslib.as_array(h.ampl, (h.nhar,))
                o["phse"][i, :h.nhar] = np.ctypeslib.as_array(h.phse, (h.nhar,))
        nm = C.cast(L.llsm_container_get(fr, U.NMI), C.POINTER(U.NM)).contents
        o["psd"][i] = np.ctypeslib.as_array(nm.psd, (npsd,))
        o["edc"][i] = np.ctypeslib.as_array(nm.edc, (nchannel,))
        for c in range(nchannel):
            e = nm.eenv[c].contents
            o["enhar"][i, c] = e.nhar
            if e.nhar:
                o["eampl"][i, c, :e.nhar] = np.ctypeslib.as_array(e.ampl, (e.nhar,))
                o["ephse"][i, c, :e.nhar] = np.ctypeslib.as_array(e.phse, (e.nhar,))
        r = L.llsm_container_get(fr, U.PSDRES)
        if r:
            o["psdres"][i] = np.ctypeslib.as_array(C.cast(r, U.fp), (npsd,))
    return o


def arctic_options(L, fs, nhop, method):
    """llsm_aoptions of test/test-layer0-anasynth.c:29-37."""
    ao = L.llsm_create_aoptions()
    ao.contents.thop = np.float32(nhop) / np.float32(fs)
    ao.contents.npsd = 128
    ao.contents.maxnhar = 400
    ao.contents.maxnhar_e = 5
    ao.contents.hm_method = 1 if method == "czt" else 0
    return ao


def anasynth(L, x, fs, f0, nhop, method, seeds=(7, 8)):
    """test/test-layer0-anasynth.c:29-63 on library L. Returns dict(chunk=flat members, f0=refined track the
    caller's array was overwritten with, out1=(y, y_sin, y_noise), chunk2=members after the phase operations,
    out2=...)."""
    x = np.ascontiguousarray(x, np.float32)
    f0c = np.array(f0, np.float32, copy=True)
    nfrm = len(f0c)
    ao = arctic_options(L, fs, nhop, method)
    so = L.llsm_create_soptions(C.c_float(fs))
    ck = L.llsm_analyze(ao, x.ctypes.data_as(U.fp), len(x), C.c_float(fs), f0c.ctypes.data_as(U.fp), nfrm, None)
    assert ck, "llsm_analyze returned NULL"
    res = dict(f0=f0c, chunk=chunk_flat(L, ck, nfrm, 400, 5, 128, 4))
    libc.srand(seeds[0])
    o = L.llsm_synthesize(so, ck)
    assert o, "llsm_synthesize returned NULL"
    res["out1"] = U.output_arrays(o)
    L.llsm_delete_output(o)
    L.llsm_chunk_phasesync_rps(ck, 0)
    L.llsm_chunk_phasepropagate(ck, 1)
    res["chunk2"] = chunk_flat(L, ck, nfrm, 400, 5, 128, 4)
    libc.srand(seeds[1])
    o = L.llsm_synthesize(so, ck)
    assert o, "llsm_synthesize returned NULL"
    res["out2"] = U.output_arrays(o)
    L.llsm_delete_output(o)
    L.llsm_delete_chunk(ck); L.llsm_delete_aoptions(ao); L.llsm_delete_soptions(so)
    return res


# ---------------------------------------------------------------- test/verify-utils.h
def empirical_kld(x, y):
    """verify-utils.h:8-30 (Perez-Cruz 2008): KL estimate between the samples x and their approximation y."""
    xs = np.sort(np.asarray(x, np.float64))
    ys = np.sort(np.asarray(y, np.float64))
    nx, ny = len(xs), len(ys)
    ys[0] = xs[0]
    ys[-1] = xs[-1]
    # yi = first index with ys[yi] >= xs[i], capped at ny - 1 and floored at 1 (the C loop only moves forward)
    yi = np.minimum(np.searchsorted(ys, xs, side="left"), ny - 1)
    yi = np.maximum(np.maximum.accumulate(yi), 1)
    xi = np.maximum(np.arange(nx), 1)
    dx = np.maximum(xs[xi] - xs[xi - 1], 1e-10)
    dy = np.maximum(ys[yi] - ys[yi - 1], 1e-10)
    return float(np.sum(np.log(ny * dy / nx / dx)) / nx - 1.0)


def _dither(v, rng):
    return v + rng.normal(0, 1.0, len(v)) * 1e-4


def verify_data_distribution(x, y, seed=0):
    """verify-utils.h:76-110: KLD of the waveform, its first and its second difference (bar: each < 0.05)."""
    rng = np.random.default_rng(seed)
    x = np.asarray(x, np.float64); y = np.asarray(y, np.float64)
    k0 = empirical_kld(_dither(x, rng), _dither(y, rng))
    # the C code leaves element 0 (and the last one) of the dithered copies from the previous stage
    k1 = empirical_kld(_dither(np.diff(x), rng), _dither(np.diff(y), rng))
    k2 = empirical_kld(_dither(np.diff(x, 2), rng), _dither(np.diff(y, 2), rng))
    return k0, k1, k2


def _stft_mag(x, nhop=512, hop_fc=4):
    nfft = nhop * hop_fc
    nfrm = len(x) // nhop
    win = 0.42 - 0.5 * np.cos(2 * np.pi * np.arange(nfft) / nfft) + 0.08 * np.cos(4 * np.pi * np.arange(nfft) / nfft)
    pad = np.concatenate([np.zeros(nfft), np.asarray(x, np.float64), np.zeros(nfft)])
    out = np.zeros((nfrm, nfft // 2 + 1))
    for i in range(nfrm):
        c = i * nhop + nfft
        out[i] = np.abs(np.fft.rfft(pad[c - nfft // 2:c + nfft // 2] * win)) * 2.0 / win.sum()
    return out


def verify_spectral_distribution(x, y, seed=0):
    """verify-utils.h:121-171: correlation of the STFT magnitudes (bar > 0.95), KLD of their distribution and of
    their frame-to-frame difference (bar < 0.05 each)."""
    rng = np.random.default_rng(seed)
    X, Y = _stft_mag(x), _stft_mag(y)
    m = min(len(X), len(Y))
    cc = float(np.corrcoef(X[:m].ravel(), Y[:m].ravel())[0, 1])
    k0 = empirical_kld(_dither(X.ravel(), rng), _dither(Y.ravel(), rng))
    k1 = empirical_kld(_dither(np.diff(X, axis=0).ravel(), rng), _dither(np.diff(Y, axis=0).ravel(), rng))
    return cc, k0, k1
