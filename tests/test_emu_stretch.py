"""Frame interpolation / time-stretch (test/demo-stretch.c:6-129,169-185): the CPU thread
emulation of the kernel against the reference's own interp_llsm_frame, and the host map helper."""
import ctypes as C
import numpy as np
import pytest
import support as S
from libllsm2_b200 import abi
from libllsm2_b200._lib import lib


def _l1struct(d, nspec):
    s = abi.Layer1()
    for k in ("rd", "vtmagn", "vsphse", "nvs"):
        setattr(s, k, d[k].ctypes.data)
    s.nspec = nspec
    return s


def emu_stretch(fr, conf, l1, base, ratio, residx, per_utt=False):
    emu = S.load_emu()
    B, n = conf.nutt, conf.nchannel
    Fn = base.shape[-1]
    nspec = l1["vtmagn"].shape[-1]
    o = dict(f0=np.zeros((B, Fn), np.float32), rd=np.zeros((B, Fn), np.float32), vtmagn=np.zeros((B, Fn, nspec), np.float32),
             vsphse=np.zeros((B, Fn, conf.maxnhar), np.float32), nvs=np.zeros((B, Fn), np.int32),
             psd=np.zeros((B, Fn, conf.npsd), np.float32), psdres=np.zeros((B, Fn, conf.npsd), np.float32),
             edc=np.zeros((B, Fn, n), np.float32), enhar=np.zeros((B, Fn, n), np.int32),
             eampl=np.zeros((B, Fn, n, conf.maxnhar_e), np.float32), ephse=np.zeros((B, Fn, n, conf.maxnhar_e), np.float32),
             nhar=np.zeros((B, Fn), np.int32), ampl=np.zeros((B, Fn, conf.maxnhar), np.float32),
             phse=np.zeros((B, Fn, conf.maxnhar), np.float32))
    src, dst = S.frames_struct(fr), S.frames_out_struct(o)
    sl, dl = _l1struct(l1, nspec), _l1struct(o, nspec)
    rc = emu.emu_frames_stretch(C.byref(conf), C.byref(src), C.byref(sl), Fn, base.ctypes.data_as(C.c_void_p),
                                ratio.ctypes.data_as(C.c_void_p),
                                residx.ctypes.data_as(C.c_void_p) if residx is not None else None,
                                1 if per_utt else 0, C.byref(dst), C.byref(dl))
    assert rc == 0
    return o


def _jitter(res, F, seed):
    rng = np.random.default_rng(seed)
    return np.clip(res + rng.integers(-2, 3, res.shape), 0, F - 1).astype(np.int32)          # test/demo-stretch.c:173-174


def test_map_follows_the_demo_arithmetic():
    L = lib()
    for F, Fn in ((24, 48), (400, 613), (7, 3), (1000, 1000)):
        base, ratio, res = (np.zeros(Fn, np.int32), np.zeros(Fn, np.float32), np.zeros(Fn, np.int32))
        assert L.llsm_b200_stretch_map(F, Fn, base.ctypes.data, ratio.ctypes.data, res.ctypes.data) == 0
        mapped = (np.arange(Fn, dtype=np.float32) * np.float32(F)) / np.float32(Fn)           # test/demo-stretch.c:170
        b = mapped.astype(np.int32)
        assert np.array_equal(res, b)
        assert np.array_equal(ratio, mapped - b.astype(np.float32))
        assert np.array_equal(base, np.minimum(b, F - 2))
    assert L.llsm_b200_stretch_map(1, 4, base.ctypes.data, ratio.ctypes.data, None) != 0      # needs two frames


@pytest.mark.parametrize("factor", [2.0, 0.6])
def test_stretch_matches_reference(factor):
    fr, conf, l1 = S.stretch_case(B=2, F=24)
    Fn = int(round(conf.nfrm * factor))
    from libllsm2_b200 import stretch_map
    base, ratio, res = stretch_map(conf.nfrm, Fn)
    res = _jitter(res, conf.nfrm, 5)
    ref = S.ref_stretch(fr, conf, l1, base, ratio, res)
    v = (ref["f0"] > 0)
    assert v.any() and (~v).any()
    o = emu_stretch(fr, conf, l1, base, ratio, res)
    S.check_stretch(o, ref, exact=True)
    # the layer-0 harmonics ride along from frame `base`
    assert np.array_equal(o["nhar"], fr["nhar"][:, base])
    assert np.array_equal(o["ampl"], fr["ampl"][:, base])


def test_all_voicing_cases_and_edge_ratios():
    """ratio 0 reproduces frame `base` (apart from the documented quirks: the -80 dB floor, the fade of a voiced frame
    next to an unvoiced one is 0 dB at ratio 0), ratio 1 does not reproduce frame base + 1 when the left frame has more
    harmonics; both are checked against the reference rather than assumed."""
    fr, conf, l1 = S.stretch_case(B=1, F=24, seed=21)
    F = conf.nfrm
    base = np.repeat(np.arange(F - 1, dtype=np.int32), 3)
    ratio = np.tile(np.array([0.0, 0.37, 1.0], np.float32), F - 1)
    ref = S.ref_stretch(fr, conf, l1, base, ratio, base)
    o = emu_stretch(fr, conf, l1, base, ratio, None)                                    # residx NULL = base
    S.check_stretch(o, ref, exact=True)
    f0 = fr["f0"][0]
    kinds = {(bool(f0[b] > 0), bool(f0[b + 1] > 0)) for b in range(F - 1)}
    assert kinds == {(True, True), (True, False), (False, True), (False, False)}
    at0 = o["f0"][0, 0::3]
    both = (f0[:-1] > 0) & (f0[1:] > 0)
    assert np.array_equal(at0[both], f0[:-1][both])


def test_per_utterance_maps():
    fr, conf, l1 = S.stretch_case(B=2, F=24, seed=31)
    from libllsm2_b200 import stretch_map
    maps = [stretch_map(conf.nfrm, 30), stretch_map(conf.nfrm, 30)]
    maps[1] = (np.minimum(maps[1][0] + 1, conf.nfrm - 2).astype(np.int32), maps[1][1][::-1].copy(), maps[1][2])
    base, ratio, res = (np.stack([m[k] for m in maps]) for k in range(3))
    o = emu_stretch(fr, conf, l1, base, ratio, res, per_utt=True)
    for b in range(2):
        one = {k: (v[b:b + 1] if v is not None else None) for k, v in fr.items()}
        l1b = {k: v[b:b + 1] for k, v in l1.items()}
        c1 = abi.make_conf(1, conf.nfrm, conf.maxnhar, conf.maxnhar_e, conf.npsd, conf.nchannel, conf.fs, conf.thop)
        ref = S.ref_stretch(one, c1, l1b, base[b], ratio[b], res[b])
        S.check_stretch({k: o[k][b:b + 1] for k in S.STRETCH_KEYS}, ref, exact=True)


def test_out_of_range_maps_are_clamped():
    """base is clamped to [0, nfrm - 2] (frame base + 1 must exist), residx to [0, nfrm - 1]."""
    fr, conf, l1 = S.stretch_case(B=1, F=24, seed=41)
    F = conf.nfrm
    base = np.array([-5, 0, F - 2, F - 1, 1000], np.int32)
    ratio = np.array([0.25, 0.25, 0.5, 0.5, 0.75], np.float32)
    res = np.array([-1, 3, F - 1, F, 1000], np.int32)
    o = emu_stretch(fr, conf, l1, base, ratio, res)
    ref = S.ref_stretch(fr, conf, l1, np.clip(base, 0, F - 2), ratio, np.clip(res, 0, F - 1))
    S.check_stretch(o, ref, exact=True)
