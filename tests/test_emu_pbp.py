"""Pulse-by-pulse (use_l1) synthesis kernels on the CPU thread emulator against the oracle."""
import ctypes as C
import numpy as np
import pytest
import support as S
from libllsm2_b200 import abi


@pytest.mark.parametrize("mode", ["middle", "all", "none"])
def test_pbp_synthesis(mode):
    B, F = 1, 30
    fr, conf = S.synth_frames(B, F, seed=3, nhar=100, maxnhar=100)
    pbp = np.zeros((B, F), np.int32)
    if mode == "middle":
        pbp[:, F // 3: 2 * F // 3] = 1
    elif mode == "all":
        pbp[:] = 1
    ref, l1 = S.ref_synthesize_l1(fr, conf, pbp, seed=9)
    ny = ref[0].shape[1]
    white = S.ref_white_noise(conf, seed=9)
    so = abi.default_soptions(white.ctypes.data, 0)
    oy = np.zeros((B, ny), np.float32); oys = np.zeros_like(oy); oyn = np.zeros_like(oy)
    out = abi.Output(); out.y = oy.ctypes.data; out.y_sin = oys.ctypes.data; out.y_noise = oyn.ctypes.data; out.stride = ny
    fr2 = dict(fr); fr2["nhar"] = None; fr2["ampl"] = None; fr2["phse"] = None     # HM removed: derived from layer 1
    f = S.frames_struct(fr2)
    s = abi.Layer1(); s.rd = l1["rd"].ctypes.data; s.vtmagn = l1["vtmagn"].ctypes.data
    s.vsphse = l1["vsphse"].ctypes.data; s.nvs = l1["nvs"].ctypes.data; s.nspec = l1["vtmagn"].shape[-1]
    emu = S.load_emu()
    assert emu.emu_synthesize_l1(C.byref(conf), C.byref(f), C.byref(s), pbp.ctypes.data_as(C.c_void_p),
                                 C.byref(so), C.byref(out)) == 0
    for r, g, name in zip(ref, (oy, oys, oyn), ("y", "y_sin", "y_noise")):
        assert S.rms(g - r) < 1e-5, (name, S.rms(g - r))


def test_pbp_ragged_batch_with_an_empty_utterance():
    """Utterances of 30, 0 and 11 frames in one batch: each row equals the reference run on that utterance alone
    (an empty utterance is silence)."""
    fr, conf, pbp, l1, refs, white, ny = S.pbp_ragged_case()
    B = conf.nutt
    so = abi.default_soptions(white.ctypes.data, 0)
    oy = np.full((B, ny), np.nan, np.float32); oys = oy.copy(); oyn = oy.copy()
    out = abi.Output(); out.y = oy.ctypes.data; out.y_sin = oys.ctypes.data; out.y_noise = oyn.ctypes.data; out.stride = ny
    f = S.frames_struct(fr)
    s = abi.Layer1(); s.rd = l1["rd"].ctypes.data; s.vtmagn = l1["vtmagn"].ctypes.data
    s.vsphse = l1["vsphse"].ctypes.data; s.nvs = l1["nvs"].ctypes.data; s.nspec = 1025
    assert S.load_emu().emu_synthesize_l1(C.byref(conf), C.byref(f), C.byref(s), pbp.ctypes.data_as(C.c_void_p),
                                          C.byref(so), C.byref(out)) == 0
    S.check_pbp_ragged((oy, oys, oyn), refs, 1e-5)
