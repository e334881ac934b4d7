#!/usr/bin/env python
"""Generate the golden fixtures from the reference build (oracle/_ref = unmodified reference
sources + ciglet shim). Run in the development container, where /root/reference exists:

    python tests/golden/make_golden.py

Each .npz holds seeded inputs and the reference's outputs for one small case; tests/test_golden.py
checks the oracle build against them on every box (so a drifting shim is caught) and the GPU path
against them on the GPU box.
"""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import support as S

CASES = {
    "synth_c2": dict(B=2, F=24, kw=dict(seed=101)),
    "synth_c1": dict(B=1, F=30, kw=dict(seed=102, thop=128 / 44100.0, nhar=300, maxnhar=300, nhar_e=5, npsd=128, f0_lo=75, f0_hi=200)),
    "synth_oddhop": dict(B=1, F=20, kw=dict(seed=103, thop=100.5 / 44100.0, nhar=50, f0_lo=200, f0_hi=330)),
}

for name, c in CASES.items():
    fr, conf = S.synth_frames(c["B"], c["F"], **c["kw"])
    y, ys, yn = S.ref_synthesize(fr, conf, seed=5)
    white = S.ref_white_noise(conf, seed=5)
    meta = np.array([conf.nutt, conf.nfrm, conf.maxnhar, conf.maxnhar_e, conf.npsd, conf.nchannel], np.int32)
    fmeta = np.array([conf.fs, conf.thop], np.float32)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), meta=meta, fmeta=fmeta, white=white,
                        y=y, y_sin=ys, y_noise=yn, **{"in_" + k: v for k, v in fr.items() if v is not None})
    print(name, y.shape, os.path.getsize(os.path.join(HERE, name + ".npz")) // 1024, "KiB")

# analysis fixture: waveform = reference synthesis of synth_c2-like parameters
fr, conf = S.synth_frames(1, 40, seed=104, nhar=100, maxnhar=100)
y, _, _ = S.ref_synthesize(fr, conf, seed=6)
ref = S.ref_analyze(y, fr["f0"], conf)
np.savez_compressed(os.path.join(HERE, "analysis_c2.npz"),
                    meta=np.array([1, 40, 100, conf.maxnhar_e, conf.npsd, conf.nchannel], np.int32),
                    fmeta=np.array([conf.fs, conf.thop], np.float32), x=y, f0_in=fr["f0"],
                    **{"out_" + k: v for k, v in ref.items()})
print("analysis_c2", os.path.getsize(os.path.join(HERE, "analysis_c2.npz")) // 1024, "KiB")
