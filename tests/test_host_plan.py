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
