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


def test_tolayer1_ragged_batch_with_an_empty_utterance():
    """Utterances of 24, 0 and 9 frames: every utterance is converted as if it were alone (the Rd track is smoothed
    along its own frames only)."""
    nfu = np.asarray([24, 0, 9], np.int32)
    fr, conf = S.synth_frames(3, 24, seed=23, nhar=100, maxnhar=100)
    nfft = 2048
    o = dict(rd=np.zeros((3, 24), np.float32), vtmagn=np.zeros((3, 24, nfft // 2 + 1), np.float32),
             vsphse=np.zeros((3, 24, conf.maxnhar), np.float32), nvs=np.zeros((3, 24), np.int32))
    fr2 = dict(fr); fr2["nfrm_utt"] = nfu
    f = S.frames_struct(fr2)
    s = _l1struct(o)
    assert S.load_emu().emu_tolayer1(C.byref(conf), C.byref(f), nfft, C.byref(s)) == 0
    for b, n in enumerate(nfu):
        if n == 0:
            continue
        one = {k: (np.ascontiguousarray(v[b:b + 1, :n]) if v is not None else None) for k, v in fr.items()}
        c1 = abi.make_conf(1, int(n), conf.maxnhar, conf.maxnhar_e, conf.npsd, conf.nchannel, conf.fs, conf.thop)
        ref = S.ref_tolayer1(one, c1, nfft)
        S.check_layer1({k: v[b:b + 1, :n] for k, v in o.items()}, ref, one["f0"] > 0)
