"""Helper of tests/test_emu_variants.py: one analysis on the CPU thread emulation, outputs saved to argv[1]. The kernel
selection switches (LLSM_*_VARIANT) are read once per process, hence a process per variant."""
import ctypes as C
import os
import sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import support as S
from libllsm2_b200 import abi

kw = dict(seed=3, nhar=100, maxnhar=128)
if sys.argv[2] == "low":
    kw.update(f0_lo=50, f0_hi=78)
fr, conf = S.synth_frames(2, 25, **kw)
y, ys, yn = S.ref_synthesize(fr, conf, seed=7)
nx = y.shape[1]
emu = S.load_emu()
o = S.alloc_analysis_out(conf, nx, fr["f0"])
ao = abi.AOptions(); ao.f0_refine = 1; ao.hm_method = 1; ao.rel_winsize = 4.0
fo = S.frames_out_struct(o)
rc = emu.emu_analyze_l0(C.byref(conf), C.byref(ao), y.ctypes.data_as(C.c_void_p), nx, nx, C.byref(fo),
                        o["x_res"].ctypes.data_as(C.c_void_p))
assert rc == 0
np.savez(sys.argv[1], **{k: o[k] for k in ("f0", "nhar", "ampl", "phse", "psd", "psdres", "edc", "enhar", "eampl", "ephse", "x_res")})
