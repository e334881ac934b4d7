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


def _l1_members(L, ck, conf, nspec):
    F = conf.nfrm
    rd = np.zeros(F, np.float32); vt = np.zeros((F, nspec), np.float32); vs = np.zeros((F, conf.maxnhar), np.float32)
    nvs = np.zeros(F, np.int32)
    for i in range(F):
        fr = ck.contents.frames[i]
        r = L.llsm_container_get(fr, 10); v = L.llsm_container_get(fr, 11); s = L.llsm_container_get(fr, 12)
        if r:
            rd[i] = C.cast(r, U.fp)[0]
        if v:
            vt[i] = np.ctypeslib.as_array(C.cast(v, U.fp), (nspec,))
        if s:
            n = L.llsm_fparray_length(C.cast(s, U.fp)); nvs[i] = n
            vs[i, :n] = np.ctypeslib.as_array(C.cast(s, U.fp), (n,))
    return dict(rd=rd[None], vtmagn=vt[None], vsphse=vs[None], nvs=nvs[None])


GFM = C.CFUNCTYPE(None, C.POINTER(C.c_float * 5), C.POINTER(C.c_float), C.c_void_p, C.c_void_p)


@pytest.mark.parametrize("effect", [False, True])
def test_layer1_dropin_roundtrip(libs, effect):
    """llsm_chunk_tolayer1 -> remove HM -> PBPSYN on/off -> llsm_synthesize(use_l1 = 1), with and without an
    llsm_pbpeffect callback (a deterministic growl-like modifier: test/test-pbpeffects.c:70-85 pattern)."""
    fr, conf = S.synth_frames(1, 90, seed=21, nhar=80, maxnhar=80)
    outs, l1s = [], []
    for L in libs:
        state = {"n": 0}

        def modifier(g, delta_t, info, frame):
            state["n"] += 1
            k = state["n"]
            g.contents[4] = g.contents[4] * (1.0 + 0.3 * np.sin(0.7 * k))      # Ee
            g.contents[0] = g.contents[0] * (1.0 - 0.2 * np.cos(0.3 * k))      # Fa
            delta_t[0] = 2e-4 * np.sin(1.3 * k)
        cb = GFM(modifier)
        ck = U.build_chunk(L, fr, conf)
        L.llsm_chunk_tolayer1(ck, 2048)
        l1s.append(_l1_members(L, ck, conf, 1025))
        L.llsm_create_pbpeffect.restype = C.c_void_p
        for i in range(conf.nfrm):
            f = ck.contents.frames[i]
            L.llsm_container_attach_(f, U.HMI, None, None, None)                      # test-layer1-anasynth.c:34
            if i % 40 > 20:
                L.llsm_container_attach_(f, 9, C.cast(L.llsm_create_int(1), C.c_void_p), U.fn_ptr(L, "llsm_delete_int"),
                                         U.fn_ptr(L, "llsm_copy_int"))
            if effect and 25 <= i < 70:
                e = L.llsm_create_pbpeffect(cb, None)
                L.llsm_container_attach_(f, 8, C.c_void_p(e), U.fn_ptr(L, "llsm_delete_pbpeffect"),
                                         U.fn_ptr(L, "llsm_copy_pbpeffect"))
        so = L.llsm_create_soptions(C.c_float(conf.fs))
        so.contents.use_l1 = 1
        libc.srand(5)
        o = L.llsm_synthesize(so, ck)
        assert o, "llsm_synthesize(use_l1) returned NULL"
        outs.append(U.output_arrays(o))
        L.llsm_delete_output(o); L.llsm_delete_soptions(so); L.llsm_delete_chunk(ck)
    S.check_layer1(l1s[0], l1s[1], (fr["f0"] > 0))
    for a, b, name in zip(outs[0], outs[1], ("y", "y_sin", "y_noise")):
        assert S.rms(a - b) < 1e-4, (name, S.rms(a - b))


def test_llsmrt_dropin(libs):
    """test-llsmrt.c's loop (feed a frame, drain the ring) on both libraries, a clear in the middle, and the
    non-blocking fetch / latency / numoutput contracts."""
    fr, conf = S.synth_frames(1, 36, seed=21, nhar=60, maxnhar=60)
    outs = []
    for L in libs:
        L.llsm_create_rtsynth_buffer.restype = C.c_void_p
        L.llsm_create_rtsynth_buffer.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        for f in ("llsm_rtsynth_buffer_feed", "llsm_rtsynth_buffer_fetch_decomposed", "llsm_rtsynth_buffer_fetch",
                  "llsm_rtsynth_buffer_clear", "llsm_delete_rtsynth_buffer", "llsm_rtsynth_buffer_getlatency",
                  "llsm_rtsynth_buffer_numoutput"):
            getattr(L, f).argtypes = [C.c_void_p] + [C.c_void_p] * (2 if f.endswith("decomposed") else 1 if f.endswith(("feed", "fetch")) else 0)
        ck = U.build_chunk(L, fr, conf)
        so = L.llsm_create_soptions(C.c_float(conf.fs))
        libc.srand(31)
        rt = L.llsm_create_rtsynth_buffer(so, ck.contents.conf, 4096)
        assert rt, "llsm_create_rtsynth_buffer returned NULL"
        lat = L.llsm_rtsynth_buffer_getlatency(rt)
        p, ap = C.c_float(), C.c_float()
        assert L.llsm_rtsynth_buffer_fetch_decomposed(rt, C.byref(p), C.byref(ap)) == 0      # empty: no wait
        ys, counts = [], []
        for i in range(conf.nfrm):
            if i == 20:
                L.llsm_rtsynth_buffer_clear(rt)
                assert L.llsm_rtsynth_buffer_numoutput(rt) == 0
            L.llsm_rtsynth_buffer_feed(rt, ck.contents.frames[i])
            counts.append(L.llsm_rtsynth_buffer_numoutput(rt))
            if i % 2 == 0:
                while L.llsm_rtsynth_buffer_fetch_decomposed(rt, C.byref(p), C.byref(ap)):
                    ys.append((p.value, ap.value))
            else:
                y = C.c_float()
                while L.llsm_rtsynth_buffer_fetch(rt, C.byref(y)):
                    ys.append((y.value, 0.0))
        L.llsm_delete_rtsynth_buffer(rt); L.llsm_delete_soptions(so); L.llsm_delete_chunk(ck)
        outs.append((lat, counts, np.array(ys, np.float32)))
    assert outs[0][0] == outs[1][0] and outs[0][1] == outs[1][1]
    a, b = outs[0][2], outs[1][2]
    assert a.shape == b.shape and S.rms(b) > 1e-3
    assert S.rms(a - b) < 1e-4, S.rms(a - b)


@pytest.mark.parametrize("effect", [False, True])
def test_llsmrt_layer1_dropin(libs, effect):
    """test-llsmrt.c / test-pbpeffects.c pattern: tolayer1, HM removed, PBPSYN toggling, use_l1 streaming, with
    and without an llsm_pbpeffect callback per frame."""
    fr, conf = S.synth_frames(1, 90, seed=21, nhar=80, maxnhar=80)
    outs = []
    for L in libs:
        state = {"n": 0}

        def modifier(g, delta_t, info, frame):
            state["n"] += 1
            k = state["n"]
            g.contents[4] = g.contents[4] * (1.0 + 0.3 * np.sin(0.7 * k))
            g.contents[0] = g.contents[0] * (1.0 - 0.2 * np.cos(0.3 * k))
            delta_t[0] = 2e-4 * np.sin(1.3 * k)
        cb = GFM(modifier)
        L.llsm_create_rtsynth_buffer.restype = C.c_void_p
        L.llsm_create_rtsynth_buffer.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.llsm_rtsynth_buffer_feed.argtypes = [C.c_void_p, C.c_void_p]
        L.llsm_rtsynth_buffer_fetch_decomposed.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.llsm_delete_rtsynth_buffer.argtypes = [C.c_void_p]
        L.llsm_create_pbpeffect.restype = C.c_void_p
        ck = U.build_chunk(L, fr, conf)
        L.llsm_chunk_tolayer1(ck, 2048)
        for i in range(conf.nfrm):
            f = ck.contents.frames[i]
            L.llsm_container_attach_(f, U.HMI, None, None, None)
            if i % 40 > 20:
                L.llsm_container_attach_(f, 9, C.cast(L.llsm_create_int(1), C.c_void_p), U.fn_ptr(L, "llsm_delete_int"),
                                         U.fn_ptr(L, "llsm_copy_int"))
            if effect and 25 <= i < 70:
                e = L.llsm_create_pbpeffect(cb, None)
                L.llsm_container_attach_(f, 8, C.c_void_p(e), U.fn_ptr(L, "llsm_delete_pbpeffect"),
                                         U.fn_ptr(L, "llsm_copy_pbpeffect"))
        so = L.llsm_create_soptions(C.c_float(conf.fs))
        so.contents.use_l1 = 1
        libc.srand(15)
        rt = L.llsm_create_rtsynth_buffer(so, ck.contents.conf, 4096)
        assert rt, "llsm_create_rtsynth_buffer(use_l1) returned NULL"
        p, ap = C.c_float(), C.c_float()
        ys = []
        for i in range(conf.nfrm):
            L.llsm_rtsynth_buffer_feed(rt, ck.contents.frames[i])
            while L.llsm_rtsynth_buffer_fetch_decomposed(rt, C.byref(p), C.byref(ap)):
                ys.append((p.value, ap.value))
        L.llsm_delete_rtsynth_buffer(rt); L.llsm_delete_soptions(so); L.llsm_delete_chunk(ck)
        outs.append((np.array(ys, np.float32), state["n"]))
    a, b = outs[0][0], outs[1][0]
    assert outs[0][1] == outs[1][1], "the effect callback must run once per pulse on both sides"
    assert a.shape == b.shape and S.rms(b[:, 0]) > 1e-3
    assert S.rms(a - b) < 1e-4, S.rms(a - b)


def test_coder_dropin(libs):
    """test/test-coder.c:24-38 on both libraries: llsm_chunk_tolayer1, llsm_create_coder(conf, 64, 5), then per frame
    llsm_coder_encode -> llsm_coder_decode_layer0 / _layer1; vectors and decoded members must agree."""
    fr, conf = S.synth_frames(1, 14, seed=23, nhar=100, maxnhar=256, f0_lo=100, f0_hi=210)
    res = []
    for L in libs:
        L.llsm_create_coder.restype = C.c_void_p
        L.llsm_create_coder.argtypes = [C.POINTER(U.Container), C.c_int, C.c_int]
        L.llsm_coder_encode.restype = U.fp
        L.llsm_coder_encode.argtypes = [C.c_void_p, C.POINTER(U.Container)]
        for fn in (L.llsm_coder_decode_layer0, L.llsm_coder_decode_layer1):
            fn.restype = C.POINTER(U.Container); fn.argtypes = [C.c_void_p, U.fp]
        L.llsm_delete_coder.argtypes = [C.c_void_p]
        ck = U.build_chunk(L, fr, conf)
        L.llsm_chunk_tolayer1(ck, 2048)
        coder = L.llsm_create_coder(ck.contents.conf, 64, 5)
        assert coder
        encs, psd0, amp0, phs0, vt1, vs1 = [], [], [], [], [], []
        for i in range(conf.nfrm):
            e = L.llsm_coder_encode(coder, ck.contents.frames[i])
            assert e, "llsm_coder_encode returned NULL"
            encs.append(np.ctypeslib.as_array(e, (72,)).copy())
            f0d = L.llsm_coder_decode_layer0(coder, e)
            f1d = L.llsm_coder_decode_layer1(coder, e)
            assert f0d and f1d
            nm = C.cast(L.llsm_container_get(f0d, U.NMI), C.POINTER(U.NM)).contents
            psd0.append(np.ctypeslib.as_array(nm.psd, (conf.npsd,)).copy())
            hm = L.llsm_container_get(f0d, U.HMI)
            if fr["f0"][0, i] > 0:
                h = C.cast(hm, C.POINTER(U.HM)).contents
                amp0.append(np.ctypeslib.as_array(h.ampl, (h.nhar,)).copy()); phs0.append(np.ctypeslib.as_array(h.phse, (h.nhar,)).copy())
                v = L.llsm_container_get(f1d, 11); s = L.llsm_container_get(f1d, 12)
                assert v and s and not L.llsm_container_get(f1d, U.HMI)
                n = L.llsm_fparray_length(C.cast(s, U.fp))
                assert n == h.nhar
                vt1.append(np.ctypeslib.as_array(C.cast(v, U.fp), (1025,)).copy()); vs1.append(np.ctypeslib.as_array(C.cast(s, U.fp), (n,)).copy())
            else:
                assert not L.llsm_container_get(f1d, 11)
            L.llsm_delete_container(f0d); L.llsm_delete_container(f1d)
            libc.free(e)
        L.llsm_delete_coder(coder); L.llsm_delete_chunk(ck)
        res.append((np.stack(encs), np.stack(psd0), amp0, phs0, vt1, vs1))
    (e_a, p_a, a_a, h_a, vt_a, vs_a), (e_b, p_b, a_b, h_b, vt_b, vs_b) = res
    # the two chunks went through two implementations of llsm_chunk_tolayer1 first: layer-1 parity bars apply
    assert np.array_equal(e_a[:, :2], e_b[:, :2]) and np.abs(e_a[:, 2] - e_b[:, 2]).max() < 1e-5
    assert np.abs(e_a[:, 3:67] - e_b[:, 3:67]).max() < 1e-3 and np.abs(e_a[:, 67:] - e_b[:, 67:]).max() < 1e-4
    assert np.abs(p_a - p_b).max() < 2e-2
    for x, y in zip(a_a, a_b):
        assert x.shape == y.shape and np.abs(x - y).max() < 1e-3 * np.abs(y).max()
    for x, y in zip(vt_a, vt_b):
        fin = np.isfinite(y)
        assert np.array_equal(fin, np.isfinite(x)) and np.abs(x[fin] - y[fin]).max() < 5e-2
    for x, y in zip(vs_a, vs_b):
        assert np.abs(S.phase_err(x, y)).max() < 1e-4
