"""Host logic: the index plans reproduce the reference's float knife-edge arithmetic
(SURVEY.md App. A/B/D)."""
import ctypes as C
import numpy as np
import support as S


def _query(nfrm, fs, thop, npsd=256):
    emu = S.load_emu()
    ints = (C.c_int * 8)()
    hb = np.zeros(nfrm, np.int32); eo = np.zeros(nfrm, np.int32)
    emu.emu_plan_query(nfrm, C.c_float(fs), C.c_float(thop), npsd, ints, hb.ctypes.data_as(C.c_void_p),
                       eo.ctypes.data_as(C.c_void_p))
    return list(ints), hb, eo


def test_c2_sizes_and_positions():
    ints, hb, eo = _query(400, 44100.0, 0.005)
    ny, n_hm, n_env, n_ns, nfft, ntemplate, nt = ints[:7]
    assert (ny, n_hm, n_env, n_ns, nfft) == (88420, 442, 441, 441, 1024)   # App. B
    assert (ntemplate, nt) == (20000, 20128)
    assert list(hb[:10]) == [0, 221, 441, 662, 882, 1102, 1323, 1544, 1764, 1984]  # App. A
    assert list(eo[:5]) == [-221, 0, 221, 441, 662]


def test_c1_sizes():
    ints, hb, eo = _query(100, 44100.0, 128 / 44100.0, 128)
    assert ints[1:5] == [256, 256, 256, 512]
    assert all(abs(int(hb[i]) - 128 * i) <= 1 for i in range(100))


def test_output_length_matches_reference():
    ref = S.load_ref()
    from libllsm2_b200 import output_length
    for nfrm in (1, 7, 400, 1154):
        for thop, fs in ((0.005, 44100.0), (128 / 44100.0, 44100.0), (0.005, 48000.0), (100.5 / 44100.0, 44100.0)):
            assert output_length(nfrm, thop, fs) == ref.ref_output_length(nfrm, C.c_float(thop), C.c_float(fs))


def _slice_plan(B, F, bytes_per_utt, cap=None, **env):
    """llsm_b200_host_slice_plan through the product library (host arithmetic only: loads without a GPU)."""
    import os
    from libllsm2_b200._lib import lib
    L = lib()
    L.llsm_b200_host_slice_plan.restype = C.c_int
    L.llsm_b200_host_slice_plan.argtypes = [C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int]
    cap = cap or B
    out = (C.c_int * cap)()
    old = {k: os.environ.get(k) for k in ("LLSM_B200_HOST_SLICES", "LLSM_B200_HOST_SLICE_LIST")}
    try:
        for k in old:
            os.environ.pop(k, None)
        os.environ.update(env)
        n = L.llsm_b200_host_slice_plan(B, F, bytes_per_utt, out, cap)
    finally:
        for k, v in old.items():
            os.environ.pop(k, None)
            if v is not None:
                os.environ[k] = v
    return list(out[:n])


def test_host_slice_plan_c2_is_tapered():
    """The anasynth host pipeline's slices at BASELINE configs[1]: small ends, two large slices (the measured optimum)."""
    per_utt = (88200 + 88420) * 4
    assert _slice_plan(1024, 400, per_utt) == [16, 144, 352, 352, 144, 16]
    big = _slice_plan(8192, 400, per_utt)                     # configs[4]'s per-GPU shard: large slices capped at 352
    assert sum(big) == 8192 and big[0] == 32 and big[-1] == 32 and max(big) <= 352 and min(big) >= 1


def test_host_slice_plan_small_batches_and_overrides():
    per_utt = (88200 + 88420) * 4
    for B in (1, 2, 3, 5, 17, 45, 46, 64, 100, 1000, 1025):
        for F in (1, 40, 400, 1154, 200000):
            s = _slice_plan(B, F, per_utt)
            assert sum(s) == B and min(s) >= 1, (B, F, s)
    assert _slice_plan(3, 70, per_utt) == [3]                                        # below 8 MB per slice: one slice
    assert _slice_plan(10, 400, per_utt, LLSM_B200_HOST_SLICES="4") == [3, 3, 3, 1]    # forced even split
    assert _slice_plan(10, 400, per_utt, LLSM_B200_HOST_SLICE_LIST="1,4") == [1, 4, 4, 1]   # the last entry repeats
    assert _slice_plan(10, 400, per_utt, LLSM_B200_HOST_SLICE_LIST="64") == [10]
    assert _slice_plan(1024, 400, per_utt, cap=3) == [16, 144, 864]                   # no room: the tail is merged
