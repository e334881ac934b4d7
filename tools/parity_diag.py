"""Parity figures of the analysis path on the GPU against the oracle, printed (not asserted): the numbers quoted in
DESIGN.md section 7. Usage (GPU box): python tools/parity_diag.py"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import support as S
import libllsm2_b200 as L

ctx = L.Context(0)
for name, kw, method in (("c2 128 harmonics, CZT", dict(seed=31, nhar=128, maxnhar=128, f0_lo=90, f0_hi=170), 1),
                         ("c2 100 harmonics, CZT", dict(seed=3, nhar=100, maxnhar=100), 1),
                         ("c2 100 harmonics, peak picking", dict(seed=13, nhar=100, maxnhar=100), 0)):
    fr, conf = S.synth_frames(2, 150, **kw)
    y, ys, yn = S.ref_synthesize(fr, conf, seed=7)
    ref = S.ref_analyze(y, fr["f0"], conf, hm_method=method)
    out = L.analyze_l0(ctx, conf, torch.from_numpy(np.ascontiguousarray(y)).cuda(), torch.from_numpy(fr["f0"]).cuda(),
                       want_residual=True, options={"hm_method": method})
    torch.cuda.synchronize()
    o = {k: v.cpu().numpy() for k, v in out.items()}
    v = ref["f0"] > 0
    print("%s: f0 identical %.4f (max %.2e Hz), nhar equal %s, ampl %.2e, phase x ampl %.2e, x_res rms %.2e (signal rms %.3f, "
          "residual rms %.2e), psd max %.4f dB rms %.5f, psdres max %.4f, edc rel %.2e, eampl %.2e, ephse x eampl %.2e" % (
        name, float((o["f0"][v] == ref["f0"][v]).mean()), float(np.abs(o["f0"] - ref["f0"]).max()),
        bool(np.array_equal(o["nhar"], ref["nhar"])), float(np.abs(o["ampl"] - ref["ampl"]).max()),
        float(np.abs(S.phase_err(o["phse"], ref["phse"]) * ref["ampl"]).max()), S.rms(o["x_res"] - ref["x_res"]), S.rms(y),
        S.rms(ref["x_res"]), float(np.abs(o["psd"] - ref["psd"]).max()), S.rms(o["psd"] - ref["psd"]),
        float(np.abs(o["psdres"] - ref["psdres"]).max()), float((np.abs(o["edc"] - ref["edc"]) / np.abs(ref["edc"])).max()),
        float(np.abs(o["eampl"] - ref["eampl"]).max()), float(np.abs(S.phase_err(o["ephse"], ref["ephse"]) * ref["eampl"]).max())))
ctx.close()
