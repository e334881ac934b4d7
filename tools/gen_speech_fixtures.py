#!/usr/bin/env python
"""Generates tests/golden/speech_fixtures.npz: the two audio fixtures the reference's own tests run on
(test/arctic_a0001.wav: test/test-layer0-anasynth.c:17, BASELINE configs[0]; test/are-you-ready.wav:
test/test-pbpeffects.c:90) as int16 samples, plus a deterministic harness F0 track for each.

The reference tracks F0 with libpyin (hop 128, 50-500 Hz: test/test-layer0-anasynth.c:19-27), which is not
part of the reference tree and absent here. llsm_analyze takes F0 as an INPUT array, so the
tracker is outside the path: the harness below (normalised autocorrelation, parabolic refinement, median
smoothing, short-run removal) only has to hand the SAME array to the reference build and to the CUDA
library. /root/reference does not exist on the GPU box, hence the committed copy.

  python tools/gen_speech_fixtures.py          (needs /root/reference)
"""
import os
import sys
import wave

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/test"
NHOP = 128


def read_wav(path):
    with wave.open(path, "rb") as w:
        assert w.getnchannels() == 1 and w.getsampwidth() == 2, (w.getnchannels(), w.getsampwidth())
        fs = w.getframerate()
        x = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").copy()
    return x, fs


def track_f0(x, fs, nhop=NHOP, fmin=50.0, fmax=500.0):
    """Deterministic harness F0 (Hz, 0 = unvoiced), one value per hop, centre of frame i at i * nhop."""
    xf = x.astype(np.float64) / 32768.0
    nfrm = len(xf) // nhop
    nwin = int(2 ** np.ceil(np.log2(fs * 0.04)))            # 2048 at 44.1 kHz
    lag_lo, lag_hi = int(fs / fmax), int(np.ceil(fs / fmin))
    pad = np.concatenate([np.zeros(nwin), xf, np.zeros(nwin)])
    win = np.hanning(nwin)
    # autocorrelation of the window itself, to normalise the taper away
    wr = np.fft.irfft(np.abs(np.fft.rfft(win, 2 * nwin)) ** 2)[:nwin]
    wr /= wr[0]
    f0 = np.zeros(nfrm)
    score = np.zeros(nfrm)
    rms_all = np.sqrt(np.mean(xf ** 2))
    for i in range(nfrm):
        c = i * nhop + nwin
        seg = pad[c - nwin // 2:c + nwin // 2]
        e = np.sqrt(np.mean(seg ** 2))
        if e < 0.05 * rms_all:
            continue
        s = (seg - seg.mean()) * win
        r = np.fft.irfft(np.abs(np.fft.rfft(s, 2 * nwin)) ** 2)[:nwin]
        if r[0] <= 0:
            continue
        r = r / r[0] / np.maximum(wr, 1e-3)
        k = lag_lo + int(np.argmax(r[lag_lo:lag_hi + 1]))
        # prefer the shortest lag whose peak is within 10 % of the best one (octave errors)
        best = r[k]
        for kk in range(lag_lo + 1, k):
            if r[kk] > 0.9 * best and r[kk] >= r[kk - 1] and r[kk] >= r[kk + 1]:
                k = kk
                break
        a, b, cc = r[k - 1], r[k], r[k + 1]
        d = a - 2 * b + cc
        off = 0.5 * (a - cc) / d if d != 0 else 0.0
        off = float(np.clip(off, -0.5, 0.5))
        f0[i] = fs / (k + off)
        score[i] = r[k]
    voiced = score > 0.55
    f0 = np.where(voiced, f0, 0.0)
    # median smoothing over voiced neighbours, then drop voiced runs shorter than 5 frames
    sm = f0.copy()
    for i in range(nfrm):
        if f0[i] > 0:
            w = f0[max(0, i - 2):i + 3]
            w = w[w > 0]
            sm[i] = np.median(w)
    f0 = sm
    i = 0
    while i < nfrm:
        if f0[i] > 0:
            j = i
            while j < nfrm and f0[j] > 0:
                j += 1
            if j - i < 5:
                f0[i:j] = 0
            i = j
        else:
            i += 1
    return np.clip(f0, 0, fmax).astype(np.float32)


def main():
    out = {}
    for key, name in (("arctic", "arctic_a0001.wav"), ("ready", "are-you-ready.wav")):
        x, fs = read_wav(os.path.join(REF, name))
        f0 = track_f0(x, fs)
        out[key + "_x"] = x
        out[key + "_fs"] = np.int32(fs)
        out[key + "_f0"] = f0
        v = f0 > 0
        print("%s: fs %d, %d samples (%.2f s), %d frames, %.0f %% voiced, f0 %.0f..%.0f Hz (median %.0f)"
              % (name, fs, len(x), len(x) / fs, len(f0), 100 * v.mean(), f0[v].min(), f0[v].max(), np.median(f0[v])))
    dst = os.path.join(ROOT, "tests", "golden", "speech_fixtures.npz")
    np.savez_compressed(dst, nhop=np.int32(NHOP), **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    sys.exit(main())
