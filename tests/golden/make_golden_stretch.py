#!/usr/bin/env python
"""Golden fixture of the frame interpolation (test/demo-stretch.c:16-129 through oracle/_ref): layer-1 frames with every
voicing transition, a frame map with jittered PSDRES indices, and interp_llsm_frame's outputs. Run where
/root/reference exists:

    python tests/golden/make_golden_stretch.py
"""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import support as S
from libllsm2_b200 import stretch_map

fr, conf, l1 = S.stretch_case(B=1, F=16, seed=201, nfft=1024)
base, ratio, res = stretch_map(conf.nfrm, 26)
res = np.clip(res + np.random.default_rng(7).integers(-2, 3, res.shape), 0, conf.nfrm - 1).astype(np.int32)
ref = S.ref_stretch(fr, conf, l1, base, ratio, res)
path = os.path.join(HERE, "stretch_l1.npz")
np.savez_compressed(path,
                    meta=np.array([conf.nutt, conf.nfrm, conf.maxnhar, conf.maxnhar_e, conf.npsd, conf.nchannel], np.int32),
                    fmeta=np.array([conf.fs, conf.thop], np.float32), base=base, ratio=ratio, residx=res,
                    **{"in_" + k: v for k, v in fr.items() if v is not None},
                    **{"l1_" + k: v for k, v in l1.items()}, **{"out_" + k: v for k, v in ref.items()})
print("stretch_l1", os.path.getsize(path) // 1024, "KiB")
