"""Drop-in data model (include/llsm.h): the host-side container / frame / chunk semantics that the
reference pins in test/test-structs.c, checked on libllsm2_b200.so and cross-checked against the
reference build. No GPU needed (no compute calls)."""
import ctypes as C
import numpy as np
import pytest
import support as S
import compat_util as U


@pytest.fixture(scope="module")
def libs():
    from libllsm2_b200._lib import lib
    return U.bind(lib()), U.bind(S.load_ref())


def test_container_attach_copy_remove(libs):
    for L in libs:
        c = L.llsm_create_container(2)
        a = L.llsm_create_fparray(5)
        for i in range(5):
            a[i] = i + 0.5
        L.llsm_container_attach_(c, 0, C.cast(a, C.c_void_p), U.fn_ptr(L, "llsm_delete_fparray"), U.fn_ptr(L, "llsm_copy_fparray"))
        shallow = L.llsm_create_fp(C.c_float(3.0))
        L.llsm_container_attach_(c, 7, C.cast(shallow, C.c_void_p), None, None)       # auto-expands
        assert c.contents.nmember == 8
        d = L.llsm_copy_container(c)
        deep = C.cast(L.llsm_container_get(d, 0), U.fp)
        assert [deep[i] for i in range(5)] == [0.5, 1.5, 2.5, 3.5, 4.5]
        assert L.llsm_fparray_length(deep) == 5
        assert C.cast(deep, C.c_void_p).value != C.cast(a, C.c_void_p).value           # deep copy
        assert L.llsm_container_get(d, 7) == C.cast(shallow, C.c_void_p).value          # NULL copy-ctor = alias
        L.llsm_container_attach_(c, 0, None, None, None)                                # the "remove" idiom
        assert not L.llsm_container_get(c, 0)
        assert not L.llsm_container_get(c, 100)
        L.llsm_delete_container(d); L.llsm_delete_container(c); L.llsm_delete_fp(shallow)


def test_hmframe_copy_and_phaseshift_roundtrip(libs):
    rng = np.random.default_rng(0)
    ph = rng.uniform(-3, 3, 40).astype(np.float32)
    res = []
    for L in libs:
        h = L.llsm_create_hmframe(40)
        for k in range(40):
            h.contents.phse[k] = float(ph[k]); h.contents.ampl[k] = 1.0 / (k + 1)
        g = L.llsm_copy_hmframe(h)
        L.llsm_hmframe_phaseshift(g, C.c_float(0.37))
        shifted = np.array([g.contents.phse[k] for k in range(40)])
        L.llsm_hmframe_phaseshift(g, C.c_float(-0.37))
        back = np.array([g.contents.phse[k] for k in range(40)])
        assert np.abs((back - ph + np.pi) % (2 * np.pi) - np.pi).max() < 1e-5
        assert np.all(np.abs(shifted) <= np.pi + 1e-6)
        res.append(shifted)
        L.llsm_delete_hmframe(g); L.llsm_delete_hmframe(h)
    assert np.abs(res[0] - res[1]).max() < 1e-6


def test_frame_defaults_and_chunk_copy(libs):
    fr, conf = S.synth_frames(1, 6, seed=2, nhar=10, maxnhar=10)
    for L in libs:
        f = L.llsm_create_frame(3, 4, 2, 16)
        nm = C.cast(L.llsm_container_get(f, U.NMI), C.POINTER(U.NM)).contents
        assert nm.npsd == 16 and nm.nchannel == 4 and abs(nm.psd[0] + 120.0) < 1e-6 and abs(nm.edc[0] - 1e-5) < 1e-9
        assert L.llsm_frame_checklayer0(f) == 1 and L.llsm_frame_checklayer1(f) == 0
        L.llsm_delete_container(f)
        ck = U.build_chunk(L, fr, conf)
        cp = L.llsm_copy_chunk(ck)
        a = U.chunk_to_flat(L, ck, conf); b = U.chunk_to_flat(L, cp, conf)
        for k in a:
            assert np.array_equal(a[k], b[k]), k
        assert np.allclose(a["ampl"], fr["ampl"][0]) and np.array_equal(a["nhar"], fr["nhar"][0])
        n = C.c_int(0)
        f0 = L.llsm_chunk_getf0(ck, C.byref(n))
        assert n.value == conf.nfrm and abs(f0[2] - fr["f0"][0, 2]) < 1e-6
        L.llsm_delete_chunk(cp); L.llsm_delete_chunk(ck)


def test_phasepropagate_matches_reference(libs):
    fr, conf = S.synth_frames(1, 12, seed=5, nhar=8, maxnhar=8)
    outs = []
    for L in libs:
        ck = U.build_chunk(L, fr, conf)
        L.llsm_chunk_phasepropagate(ck, -1)
        L.llsm_chunk_phasesync_rps(ck, 0)
        outs.append(U.chunk_to_flat(L, ck, conf)["phse"])
        L.llsm_delete_chunk(ck)
    d = (outs[0] - outs[1] + np.pi) % (2 * np.pi) - np.pi
    assert np.abs(d).max() < 2e-4        # float phase arithmetic at ~1e2 rad


def test_synthesize_rejects_bad_chunk_and_fails_without_gpu(libs):
    import torch
    L = libs[0]
    so = L.llsm_create_soptions(C.c_float(44100.0))
    fr, conf = S.synth_frames(1, 4, seed=1, nhar=4, maxnhar=4)
    ck = U.build_chunk(L, fr, conf)
    # remove NM from one frame: integrity check fails -> NULL (layer0.c:637)
    L.llsm_container_attach_(ck.contents.frames[1], U.NMI, None, None, None)
    assert not L.llsm_synthesize(so, ck)
    L.llsm_delete_chunk(ck)
    if not torch.cuda.is_available():
        ck = U.build_chunk(L, fr, conf)
        assert not L.llsm_synthesize(so, ck)      # no device: NULL, no CPU fallback
        L.llsm_last_error.restype = C.c_char_p
        assert b"CUDA" in L.llsm_last_error()
        L.llsm_delete_chunk(ck)
    L.llsm_delete_soptions(so)
