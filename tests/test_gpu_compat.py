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


def _bind_rt(L):
    L.llsm_create_rtsynth_buffer.restype = C.c_void_p
    L.llsm_create_rtsynth_buffer.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.llsm_rtsynth_buffer_feed.argtypes = [C.c_void_p, C.c_void_p]
    L.llsm_rtsynth_buffer_fetch_decomposed.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.llsm_rtsynth_buffer_numoutput.argtypes = [C.c_void_p]
    L.llsm_delete_rtsynth_buffer.argtypes = [C.c_void_p]


def test_llsmrt_two_threads_feed_blocks_on_full_ring(libs):
    """The reference's threading contract (test/test-llsmrt.c:29-63,123-124; llsmrt.c:489-499,523-543): a synthesis
    thread feeds frames and BLOCKS while the output ring is full, a reading thread polls fetch_decomposed (non-blocking,
    usleep(10) between misses). The ring holds 600 samples, less than three hops, so the feeder has to wait for the
    reader over and over. The samples must equal those of a single-threaded run of the same library, and the
    reference build run the same way must agree with it to the usual bar."""
    import threading
    import time
    fr, conf = S.synth_frames(1, 60, seed=27, nhar=60, maxnhar=60)
    results = []
    for L in libs:
        _bind_rt(L)
        runs = []
        for threaded in (False, True):
            ck = U.build_chunk(L, fr, conf)
            so = L.llsm_create_soptions(C.c_float(conf.fs))
            libc.srand(31)
            cap = 600 if threaded else 8192
            rt = L.llsm_create_rtsynth_buffer(so, ck.contents.conf, cap)
            assert rt
            ys = []
            if not threaded:
                p, ap = C.c_float(), C.c_float()
                for i in range(conf.nfrm):
                    L.llsm_rtsynth_buffer_feed(rt, ck.contents.frames[i])
                    while L.llsm_rtsynth_buffer_fetch_decomposed(rt, C.byref(p), C.byref(ap)):
                        ys.append((p.value, ap.value))
            else:
                done = threading.Event()
                blocked = {"full": 0}

                def feeder():
                    for i in range(conf.nfrm):
                        if L.llsm_rtsynth_buffer_numoutput(rt) > cap - 230:
                            blocked["full"] += 1                         # this feed will have to wait for the reader
                        L.llsm_rtsynth_buffer_feed(rt, ck.contents.frames[i])
                    done.set()

                def reader():
                    p, ap = C.c_float(), C.c_float()
                    idle = 0
                    while True:
                        if L.llsm_rtsynth_buffer_fetch_decomposed(rt, C.byref(p), C.byref(ap)):
                            ys.append((p.value, ap.value)); idle = 0
                        elif done.is_set():
                            idle += 1
                            if idle > 3:
                                break
                        else:
                            time.sleep(1e-5)
                t1, t2 = threading.Thread(target=feeder), threading.Thread(target=reader)
                t1.start(); t2.start()
                t1.join(timeout=120); t2.join(timeout=120)
                assert not t1.is_alive() and not t2.is_alive(), "feeder / reader deadlocked"
                assert blocked["full"] > 5, blocked
            L.llsm_delete_rtsynth_buffer(rt); L.llsm_delete_soptions(so); L.llsm_delete_chunk(ck)
            runs.append(np.array(ys, np.float32))
        assert runs[0].shape == runs[1].shape and np.array_equal(runs[0], runs[1]), "threaded run differs"
        results.append(runs[1])
    a, b = results
    assert a.shape == b.shape and S.rms(b) > 1e-3
    assert S.rms(a - b) < 1e-4, S.rms(a - b)


def test_compat_entries_are_thread_safe(libs):
    """Every drop-in call goes through one process-wide device context; concurrent callers must not corrupt each
    other's staging buffers (the reference's functions are reentrant). Two threads run llsm_chunk_tolayer1 +
    llsm_synthesize(use_l1) on different chunks at the same time, many times; each must reproduce its own
    single-threaded result bit for bit."""
    import threading
    L = libs[0]
    cases = []
    for seed, F in ((51, 50), (52, 64)):
        fr, conf = S.synth_frames(1, F, seed=seed, nhar=70, maxnhar=70)
        cases.append((fr, conf))

    def run(fr, conf):
        ck = U.build_chunk(L, fr, conf)
        L.llsm_chunk_tolayer1(ck, 2048)
        l1 = _l1_members(L, ck, conf, 1025)
        so = L.llsm_create_soptions(C.c_float(conf.fs))
        so.contents.use_l1 = 1
        o = L.llsm_synthesize(so, ck)
        assert o
        ys = U.output_arrays(o)[1]                       # y_sin: deterministic (the noise draws from the shared rand())
        L.llsm_delete_output(o); L.llsm_delete_soptions(so); L.llsm_delete_chunk(ck)
        return l1, ys

    base = [run(fr, conf) for fr, conf in cases]
    errs = []

    def worker(k):
        try:
            for _ in range(6):
                l1, ys = run(*cases[k])
                for key in ("rd", "vtmagn", "vsphse", "nvs"):
                    assert np.array_equal(l1[key], base[k][0][key]), key
                assert np.array_equal(ys, base[k][1])
        except BaseException as e:                       # noqa: BLE001
            errs.append((k, repr(e)))
    ts = [threading.Thread(target=worker, args=(k,)) for k in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=300)
    assert not errs, errs


def _bind_dsp(L):
    L.llsm_harmonic_analysis.argtypes = [U.fp, C.c_int, C.c_float, U.fp, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int,
                                         C.POINTER(C.c_int), C.POINTER(U.fp), C.POINTER(U.fp)]
    L.llsm_harmonic_analysis.restype = None
    for f in (L.llsm_synthesize_harmonic_frame, L.llsm_synthesize_harmonic_frame_iczt):
        f.argtypes = [U.fp, U.fp, C.c_int, C.c_float, C.c_int]
        f.restype = U.fp
    L.llsm_get_fftsize.argtypes = [U.fp, C.c_int, C.c_float, C.c_float]
    L.llsm_refine_f0.argtypes = [U.fp, C.c_int, C.c_float, U.fp, C.c_int, C.c_float]
    L.llsm_refine_f0.restype = None


@pytest.mark.parametrize("method", [0, 1])    # LLSM_AOPTION_HMPP, LLSM_AOPTION_HMCZT
def test_dsputils_harmonic_analysis_dropin(libs, method):
    """test/test-dsputils.c:44-133 through the exported llsm_harmonic_analysis of both libraries: the three-harmonic
    chirp, the reference's tolerances (amplitude mean / std error < 0.01, phase-advance error < 0.1 rad) on each, and
    the two libraries against each other."""
    import math
    nx, fs, thop = 100000, 20000.0, 0.005
    nfrm = int(math.floor(np.float32(nx) / np.float32(fs) / np.float32(thop)))
    center = np.round(np.arange(nfrm) * np.float32(thop) * np.float32(fs)).astype(int)
    rate = center.astype(np.float32) / nx
    f0 = (100 + 100 * rate).astype(np.float32)
    i = np.arange(nx)
    ph = np.cumsum((100 + 100 * i / nx) / fs * 2 * 3.1415927)
    x = np.ascontiguousarray(((i / nx) * np.sin(ph) + 0.5 * np.sin(2 * ph) + 0.25 * np.sin(3 * ph)).astype(np.float32))
    got = []
    for L in libs:
        _bind_dsp(L)
        nhar = (C.c_int * nfrm)()
        ampl = (U.fp * nfrm)(); phse = (U.fp * nfrm)()
        f0c = f0.copy()
        assert L.llsm_get_fftsize(f0c.ctypes.data_as(U.fp), nfrm, C.c_float(fs), C.c_float(4.0)) == 1024
        L.llsm_harmonic_analysis(x.ctypes.data_as(U.fp), nx, C.c_float(fs), f0c.ctypes.data_as(U.fp), nfrm, C.c_float(thop),
                                 C.c_float(4.0), 3, method, nhar, ampl, phse)
        A = np.zeros((nfrm, 3), np.float32); P = np.zeros((nfrm, 3), np.float32)
        for t in range(nfrm):
            assert nhar[t] == 3 and ampl[t] and phse[t]
            A[t] = np.ctypeslib.as_array(ampl[t], (3,)); P[t] = np.ctypeslib.as_array(phse[t], (3,))
            libc.free(ampl[t]); libc.free(phse[t])
        s = slice(5, nfrm - 5)
        for k, truth in enumerate([rate, 0.5, 0.25]):
            e = np.zeros(nfrm); e[s] = (A[:, k] - truth)[s]
            assert abs(e.mean()) < 0.01 and e.std() < 0.01, (k, e.mean(), e.std())
        pe = np.zeros(nfrm - 1)
        for t in range(5, nfrm - 5):
            d = P[t, 0] - (P[t - 1, 0] + f0[t] * 2 * 3.1415927 * thop)
            pe[t - 1] = (d + math.pi) % (2 * math.pi) - math.pi
        assert abs(pe.mean()) < 0.1 and pe.std() < 0.1, (pe.mean(), pe.std())
        got.append((A, P))
    (Aa, Pa), (Ab, Pb) = got
    # amplitudes reach 1.0 here: bars relative to that (the analysis tests' 1e-6 is on amplitudes <= 0.1)
    assert np.abs(Aa - Ab).max() < (5e-6 if method == 1 else 1e-4)
    assert np.abs(S.phase_err(Pa, Pb) * Ab).max() < (5e-6 if method == 1 else 1e-3)


def test_dsputils_harmonic_frame_dropin(libs):
    """test/test-harmonic.c:29-48 through both libraries: 100 harmonics, f0 = 0.01 cycles / sample, 1024 samples; the
    ICZT and the sinusoid-bank frame must agree (SNR printed by the reference, > 80 dB asserted here) and each must
    match the reference build's frame."""
    rng = np.random.default_rng(3)
    nhar, nx, f0 = 100, 1024, 0.01
    ampl = np.ascontiguousarray(rng.uniform(0, 1, nhar).astype(np.float32))
    phse = np.ascontiguousarray(rng.uniform(-np.pi, np.pi, nhar).astype(np.float32))
    frames = []
    for L in libs:
        _bind_dsp(L)
        out = []
        for fn in (L.llsm_synthesize_harmonic_frame_iczt, L.llsm_synthesize_harmonic_frame):
            y = fn(ampl.ctypes.data_as(U.fp), phse.ctypes.data_as(U.fp), nhar, C.c_float(f0), nx)
            assert y
            out.append(np.ctypeslib.as_array(y, (nx,)).copy())
            libc.free(y)
        snr = 10 * np.log10(np.sum(out[1].astype(np.float64) ** 2) / np.sum((out[0].astype(np.float64) - out[1]) ** 2))
        assert snr > 80, snr
        frames.append(out)
    for a, b in zip(frames[0], frames[1]):
        assert S.rms(a - b) < 1e-4 * S.rms(b), S.rms(a - b) / S.rms(b)
    # more harmonics than samples: the ICZT branch drops those beyond the transform length (dsputils.c:341-348)
    nh2, nx2 = 300, 256
    a2 = np.ascontiguousarray(rng.uniform(0, 1, nh2).astype(np.float32)); p2 = np.ascontiguousarray(rng.uniform(-3, 3, nh2).astype(np.float32))
    ys = []
    for L in libs:
        y = L.llsm_synthesize_harmonic_frame_iczt(a2.ctypes.data_as(U.fp), p2.ctypes.data_as(U.fp), nh2, C.c_float(0.0015), nx2)
        ys.append(np.ctypeslib.as_array(y, (nx2,)).copy()); libc.free(y)
    assert S.rms(ys[0] - ys[1]) < 1e-4 * S.rms(ys[1])


def test_dsputils_refine_f0_dropin(libs):
    fr, conf = S.synth_frames(1, 60, seed=61, nhar=60, maxnhar=60)
    y, _, _ = S.ref_synthesize(fr, conf, seed=3)
    x = np.ascontiguousarray(y[0])
    outs = []
    for L in libs:
        _bind_dsp(L)
        f0 = (fr["f0"][0] * np.float32(1.01)).astype(np.float32)         # start one per cent off
        L.llsm_refine_f0(x.ctypes.data_as(U.fp), len(x), C.c_float(conf.fs), f0.ctypes.data_as(U.fp), conf.nfrm, C.c_float(conf.thop))
        outs.append(f0)
    assert np.abs(outs[0] - outs[1]).max() < 1e-3
    v = fr["f0"][0] > 0
    start = (fr["f0"][0] * np.float32(1.01)).astype(np.float32)
    assert np.abs(outs[1][v] - start[v]).mean() > 0.1                  # the track was really re-estimated
    assert np.array_equal(outs[0][~v], start[~v])                      # unvoiced frames untouched


def test_llsmrt_attaches_hm_like_the_reference(libs):
    """Side effect of llsm_rtsynth_buffer_feed (llsmrt.c:341-347,387-390): a layer-1 frame without a harmonic model gets
    one attached (llsm_frame_tolayer0) whenever the stream plays its sinusoids -- PbP onset or not in PbP mode --, and
    keeps none while the stream stays in PbP mode. Same frames on both libraries, same model."""
    fr, conf = S.synth_frames(1, 70, seed=71, nhar=60, maxnhar=60)
    res = []
    for L in libs:
        _bind_rt(L)
        ck = U.build_chunk(L, fr, conf)
        L.llsm_chunk_tolayer1(ck, 2048)
        for i in range(conf.nfrm):
            f = ck.contents.frames[i]
            L.llsm_container_attach_(f, U.HMI, None, None, None)
            if 20 <= i < 45:
                L.llsm_container_attach_(f, 9, C.cast(L.llsm_create_int(1), C.c_void_p), U.fn_ptr(L, "llsm_delete_int"),
                                         U.fn_ptr(L, "llsm_copy_int"))
        so = L.llsm_create_soptions(C.c_float(conf.fs))
        so.contents.use_l1 = 1
        libc.srand(3)
        rt = L.llsm_create_rtsynth_buffer(so, ck.contents.conf, 8192)
        p, ap = C.c_float(), C.c_float()
        for i in range(conf.nfrm):
            L.llsm_rtsynth_buffer_feed(rt, ck.contents.frames[i])
            while L.llsm_rtsynth_buffer_fetch_decomposed(rt, C.byref(p), C.byref(ap)):
                pass
        has, amp = [], []
        for i in range(conf.nfrm):
            hm = L.llsm_container_get(ck.contents.frames[i], U.HMI)
            has.append(bool(hm))
            if hm:
                h = C.cast(hm, C.POINTER(U.HM)).contents
                amp.append(np.ctypeslib.as_array(h.ampl, (h.nhar,)).copy())
        L.llsm_delete_rtsynth_buffer(rt); L.llsm_delete_soptions(so); L.llsm_delete_chunk(ck)
        res.append((has, amp))
    assert res[0][0] == res[1][0], "HM members attached to different frames"
    assert any(res[1][0]) and not all(h for h, v in zip(res[1][0], fr["f0"][0] > 0) if v)
    for a, b in zip(res[0][1], res[1][1]):
        assert a.shape == b.shape and np.abs(a - b).max() < 1e-4 * max(float(np.abs(b).max()), 1e-9)
