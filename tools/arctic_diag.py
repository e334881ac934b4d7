"""Diagnostic (GPU box): where do the analysis outputs of the CUDA library and of the reference build differ on
arctic_a0001.wav? Prints per-member error statistics and the frames that carry the largest noise-PSD differences."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import support as S, compat_util as U, speech_util as SU
from libllsm2_b200._lib import lib

libs = (U.bind(lib()), U.bind(S.load_ref()))
fx = SU.fixtures()["arctic"]
for method in ("pp", "czt"):
    a, b = [SU.anasynth(L, fx["x"], fx["fs"], fx["f0"], fx["nhop"], method) for L in libs]
    ca, cb = a["chunk"], b["chunk"]
    d = np.abs(ca["psd"] - cb["psd"])
    fr_max = d.max(1)
    print("== %s: psd |diff| max %.4f dB, 99.9 %% %.5f, 99 %% %.5f, median %.6f; frames > 0.05 dB: %d of %d"
          % (method, d.max(), np.quantile(d, 0.999), np.quantile(d, 0.99), np.median(d), int((fr_max > 0.05).sum()), len(fr_max)))
    worst = np.argsort(-fr_max)[:12]
    xe = fx["x"]
    for i in sorted(worst):
        c = i * fx["nhop"]
        seg = xe[max(0, c - 256):c + 256]
        print("   frame %4d f0 %6.1f  max diff %.4f dB at bin %3d  psd there %.1f dB  frame rms %.2e"
              % (i, cb["f0"][i], fr_max[i], int(d[i].argmax()), cb["psd"][i, int(d[i].argmax())], float(np.sqrt(np.mean(seg ** 2)))))
    print("   f0 max diff %.2e, nhar equal %s, ampl max diff %.3e (scale %.3f), psdres max %.4f, edc rel %.2e"
          % (np.abs(ca["f0"] - cb["f0"]).max(), np.array_equal(ca["nhar"], cb["nhar"]), np.abs(ca["ampl"] - cb["ampl"]).max(),
             np.abs(cb["ampl"]).max(), np.abs(ca["psdres"] - cb["psdres"]).max(),
             (np.abs(ca["edc"] - cb["edc"]) / np.abs(cb["edc"]).max()).max()))
    for key in ("out1", "out2"):
        print("   %s waveform RMS error: y %.2e  y_sin %.2e  y_noise %.2e   (rms y %.3f)" % (
            key, S.rms(a[key][0] - b[key][0]), S.rms(a[key][1] - b[key][1]), S.rms(a[key][2] - b[key][2]), S.rms(b[key][0])))
