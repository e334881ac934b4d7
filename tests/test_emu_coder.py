"""Frame coder (llsm_coder_encode / llsm_coder_decode_layer0 / _layer1, coder.c) of the CPU thread emulation of the
kernels against the reference build: encode real layer-1 frames, decode them both ways, and the round trip."""
import ctypes as C
import numpy as np
import pytest
import support as S
from libllsm2_b200 import abi

OS, OB = 64, 5          # test/test-coder.c:31


def _l1struct(d, nspec):
    s = abi.Layer1()
    for k in ("rd", "vtmagn", "vsphse", "nvs"):
        if k in d:
            setattr(s, k, d[k].ctypes.data)
    s.nspec = nspec
    return s


def _frames(seed=3, B=1, F=10):
    fr, conf = S.synth_frames(B, F, seed=seed, nhar=100, maxnhar=256, f0_lo=100, f0_hi=210)
    l1 = S.ref_tolayer1(fr, conf, 2048)
    return fr, conf, l1


def emu_encode(fr, conf, l1):
    emu = S.load_emu()
    nspec = l1["vtmagn"].shape[-1]
    enc = np.zeros((conf.nutt, conf.nfrm, OS + OB + 3), np.float32)
    s = _l1struct(l1, nspec)
    assert emu.emu_coder_encode(C.byref(conf), fr["f0"].ctypes.data_as(C.c_void_p), fr["psd"].ctypes.data_as(C.c_void_p),
                                C.byref(s), OS, OB, enc.ctypes.data_as(C.c_void_p)) == 0
    return enc


def emu_decode(enc, conf, nspec, use_layer1):
    emu = S.load_emu()
    B, F = conf.nutt, conf.nfrm
    o = dict(f0=np.zeros((B, F), np.float32), rd=np.zeros((B, F), np.float32), psd=np.zeros((B, F, conf.npsd), np.float32),
             nhar=np.zeros((B, F), np.int32), ampl=np.zeros((B, F, conf.maxnhar), np.float32),
             phse=np.zeros((B, F, conf.maxnhar), np.float32), vtmagn=np.zeros((B, F, nspec), np.float32),
             vsphse=np.zeros((B, F, conf.maxnhar), np.float32))
    fo = abi.FramesOut()
    fo.f0, fo.nhar, fo.ampl, fo.phse, fo.psd = (o["f0"].ctypes.data, o["nhar"].ctypes.data, o["ampl"].ctypes.data,
                                                o["phse"].ctypes.data, o["psd"].ctypes.data)
    s = _l1struct(o, nspec)
    assert emu.emu_coder_decode(C.byref(conf), enc.ctypes.data_as(C.c_void_p), OS, OB, int(use_layer1), C.byref(fo),
                                C.byref(s)) == 0
    return o


def test_encode_matches_reference():
    fr, conf, l1 = _frames()
    ref = S.ref_coder_encode(fr["f0"], fr["psd"], l1, conf, OS, OB)
    assert ref[..., 0].min() == 0 and ref[..., 0].max() == 1            # voiced and unvoiced frames both present
    S.check_coder_encode(emu_encode(fr, conf, l1), ref, OS)


@pytest.mark.parametrize("use_layer1", [1, 0])
def test_decode_matches_reference(use_layer1):
    fr, conf, l1 = _frames(seed=4, F=8)
    enc = S.ref_coder_encode(fr["f0"], fr["psd"], l1, conf, OS, OB)
    nspec = l1["vtmagn"].shape[-1]
    ref = S.ref_coder_decode(enc, conf, nspec, OS, OB, use_layer1)
    S.check_coder_decode(emu_decode(enc, conf, nspec, use_layer1), ref, use_layer1)


def test_round_trip_keeps_the_envelope():
    """encode -> decode_layer1 reproduces a smooth version of the vocal-tract envelope (64 mel-cepstral numbers):
    within a few dB over the speech band, as in the reference (test/test-coder.c listens to exactly this)."""
    fr, conf, l1 = _frames(seed=5, F=6)
    nspec = l1["vtmagn"].shape[-1]
    o = emu_decode(emu_encode(fr, conf, l1), conf, nspec, 1)
    v = fr["f0"] > 0
    band = slice(int(300 / 22050 * nspec), int(6000 / 22050 * nspec))
    d = np.abs(o["vtmagn"][v][:, band] - l1["vtmagn"][v][:, band])
    assert np.median(d) < 3.0, np.median(d)
    assert np.array_equal(o["f0"], fr["f0"])
