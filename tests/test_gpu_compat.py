"""The drop-in entry points llsm_synthesize / llsm_analyze of libllsm2_b200.so against the same
calls on the reference build, chunk in / chunk out, same srand() state."""
import ctypes as C
import numpy as np
import pytest
import support as S
import compat_util as U
from test_emu_analysis import check_analysis

pytestmark = pytest.mark.gpu
libc = C.CDLL(None)


@pytest.fixture(scope="module")
def libs():
    from libllsm2_b200._lib import lib
    return U.bind(lib()), U.bind(S.load_ref())


def test_llsm_synthesize_dropin(libs):
    fr, conf = S.synth_frames(1, 60, seed=12, nhar=60, maxnhar=60)
    outs = []
    for L in libs:
        ck = U.build_chunk(L, fr, conf)
        so = L.llsm_create_soptions(C.c_float(conf.fs))
        libc.srand(77)
        o = L.llsm_synthesize(so, ck)
        assert o, "llsm_synthesize returned NULL"
        outs.append(U.output_arrays(o))
        L.llsm_delete_output(o); L.llsm_delete_soptions(so); L.llsm_delete_chunk(ck)
    for a, b, name in zip(outs[0], outs[1], ("y", "y_sin", "y_noise")):
        assert a.shape == b.shape
        assert S.rms(a - b) < 1e-4, (name, S.rms(a - b))


def test_llsm_analyze_dropin(libs):
    fr, conf = S.synth_frames(1, 80, seed=14, nhar=100, maxnhar=100)
    y, _, _ = S.ref_synthesize(fr, conf, seed=3)
    res = []
    for L in libs:
        ao = L.llsm_create_aoptions()
        ao.contents.maxnhar = conf.maxnhar
        f0 = fr["f0"][0].copy()
        x = np.ascontiguousarray(y[0])
        xap = U.fp()
        ck = L.llsm_analyze(ao, x.ctypes.data_as(U.fp), len(x), C.c_float(conf.fs), f0.ctypes.data_as(U.fp),
                            conf.nfrm, C.byref(xap))
        assert ck, "llsm_analyze returned NULL"
        o = U.chunk_to_flat(L, ck, conf)
        o["x_res"] = np.ctypeslib.as_array(xap, (len(x),)).copy()
        o["f0_inout"] = f0
        res.append(o)
        L.llsm_delete_chunk(ck); L.llsm_delete_aoptions(ao); libc.free(xap)
    a, b = res
    assert np.abs(a["f0_inout"] - b["f0_inout"]).max() < 1e-3      # caller's f0 is refined in place
    assert np.array_equal(a["nhar"], b["nhar"]) and np.array_equal(a["enhar"], b["enhar"])
    assert np.abs(a["ampl"] - b["ampl"]).max() < 1e-6
    assert S.rms(a["x_res"] - b["x_res"]) < 1e-6
    assert np.abs(a["psd"] - b["psd"]).max() < 0.05
