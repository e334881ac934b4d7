"""Layer-1 conversion kernels on the CPU thread emulator against the oracle (layer1.c)."""
import ctypes as C
import numpy as np
import support as S
from libllsm2_b200 import abi


def _l1struct(d):
    s = abi.Layer1()
    s.rd, s.vtmagn, s.vsphse, s.nvs = (d["rd"].ctypes.data, d["vtmagn"].ctypes.data, d["vsphse"].ctypes.data,
                                       d["nvs"].ctypes.data)
    s.nspec = d["vtmagn"].shape[-1]
    return s


def test_tolayer1_and_back():
    fr, conf = S.synth_frames(1, 24, seed=3, nhar=100, maxnhar=100)
    nfft = 2048
    ref = S.ref_tolayer1(fr, conf, nfft)
    emu = S.load_emu()
    o = {k: np.zeros_like(v) for k, v in ref.items()}
    f = S.frames_struct(fr)
    s = _l1struct(o)
    assert emu.emu_tolayer1(C.byref(conf), C.byref(f), nfft, C.byref(s)) == 0
    S.check_layer1(o, ref, fr["f0"] > 0)
    ref0 = S.ref_tolayer0(fr["f0"], ref, conf)
    o0 = {k: np.zeros_like(v) for k, v in ref0.items()}
    s = _l1struct(ref)
    assert emu.emu_tolayer0(C.byref(conf), fr["f0"].ctypes.data_as(C.c_void_p), C.byref(s),
                            o0["nhar"].ctypes.data_as(C.c_void_p), o0["ampl"].ctypes.data_as(C.c_void_p),
                            o0["phse"].ctypes.data_as(C.c_void_p)) == 0
    S.check_layer0_from_l1(o0, ref0)
