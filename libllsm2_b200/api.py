"""Host-side mirror of the reference's synthesis / analysis interface over the C ABI.

Names and argument meaning follow llsm.h: a *conf* (LLSM_CONF_* entries), per-frame members
(F0, HM = ampl/phse/nhar, NM = psd/edc/eenv, PSDRES) -- here as flat [B][F][...] arrays -- and an
output triple (y, y_sin, y_noise) as in llsm_output (llsm.h:246-252).
"""
import ctypes as C
import numpy as np

from . import abi
from ._lib import lib, check, LlsmB200Error

FRAME_KEYS = ("nfrm_utt", "f0", "nhar", "ampl", "phse", "psd", "psdres", "edc", "enhar", "eampl", "ephse")


def output_length(nfrm, thop, fs):
    """ny = round((nfrm + 1) * thop * fs) in the reference's float arithmetic (layer0.c:643)."""
    return lib().llsm_b200_output_length(int(nfrm), C.c_float(thop), C.c_float(fs))


class Context:
    """Device context (stream, cached plans, scratch). One per process / GPU."""

    def __init__(self, device=0, use_torch_stream=True):
        L = lib()
        self._h = L.llsm_b200_create(int(device))
        if not self._h:
            raise LlsmB200Error("llsm_b200_create failed: " + L.llsm_b200_last_error().decode())
        self.device = int(device)
        if use_torch_stream:
            import torch
            with torch.cuda.device(self.device):
                self.set_stream(torch.cuda.current_stream().cuda_stream)

    def set_stream(self, cuda_stream_ptr):
        check(lib().llsm_b200_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def synchronize(self):
        check(lib().llsm_b200_synchronize(self._h))

    @property
    def launches(self):
        return int(lib().llsm_b200_launch_count(self._h))

    def set_kernel_timing(self, enable=True):
        check(lib().llsm_b200_set_kernel_timing(self._h, 1 if enable else 0))

    def kernel_times(self):
        """[(name, ms), ...] of the kernels launched since set_kernel_timing(True), in launch order."""
        names = (C.c_char_p * 48)()
        ms = (C.c_float * 48)()
        n = lib().llsm_b200_kernel_timing_read(self._h, 48, names, ms)
        if n < 0:
            raise LlsmB200Error(lib().llsm_b200_last_error().decode())
        return [(names[i].decode(), float(ms[i])) for i in range(n)]

    def close(self):
        if getattr(self, "_h", None):
            lib().llsm_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _ptr(a):
    if a is None:
        return None
    if hasattr(a, "data_ptr"):          # torch tensor
        assert a.is_contiguous()
        return a.data_ptr()
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


def _frames(frames):
    f = abi.Frames()
    for k in FRAME_KEYS:
        setattr(f, k, _ptr(frames.get(k)))
    return f


def _soptions(options, white, seed):
    o = abi.default_soptions(_ptr(white), int(seed))
    if options is not None:
        o.use_iczt = int(options.get("use_iczt", 1))
        o.iczt_param_a = float(options.get("iczt_param_a", 0.275))
        o.iczt_param_b = float(options.get("iczt_param_b", 2.26))
    return o


def synthesize_l0(ctx, conf, frames, white=None, seed=0, options=None, out=None):
    """llsm_synthesize (layer0.c:636-664) for a batch held in CUDA tensors.
    frames: dict of torch CUDA tensors keyed as FRAME_KEYS. Returns dict(y, y_sin, y_noise)."""
    import torch
    ny = output_length(conf.nfrm, conf.thop, conf.fs)
    dev = frames["f0"].device
    if out is None:
        out = {k: torch.empty((conf.nutt, ny), dtype=torch.float32, device=dev)
               for k in ("y", "y_sin", "y_noise")}
    o = abi.Output()
    o.y, o.y_sin, o.y_noise, o.stride = _ptr(out["y"]), _ptr(out["y_sin"]), _ptr(out["y_noise"]), out["y"].shape[1]
    f = _frames(frames)
    so = _soptions(options, white, seed)
    check(lib().llsm_b200_synthesize_l0(ctx._h, C.byref(conf), C.byref(f), C.byref(so), C.byref(o)))
    return out


def synthesize_l0_host(ctx, conf, frames, white=None, seed=0, options=None, out=None):
    """Same through host (numpy or pinned torch CPU) buffers: H2D, kernels, D2H, synchronise."""
    ny = output_length(conf.nfrm, conf.thop, conf.fs)
    if out is None:
        out = {k: np.empty((conf.nutt, ny), np.float32) for k in ("y", "y_sin", "y_noise")}
    o = abi.Output()
    o.y, o.y_sin, o.y_noise = _ptr(out.get("y")), _ptr(out.get("y_sin")), _ptr(out.get("y_noise"))
    first = next(v for v in out.values() if v is not None)
    o.stride = first.shape[1]
    f = _frames(frames)
    so = _soptions(options, white, seed)
    check(lib().llsm_b200_synthesize_l0_host(ctx._h, C.byref(conf), C.byref(f), C.byref(so), C.byref(o)))
    return out


def synthesize_harmonics(ctx, conf, frames, nsamp, options=None, with_options=True, out=None):
    """Harmonic component only (llsm_synthesize_harmonics_l0, layer0.c:117-146), CUDA tensors.
    with_options=False reproduces the options == NULL call of the analysis residual."""
    import torch
    dev = frames["f0"].device
    if out is None:
        out = torch.empty((conf.nutt, nsamp), dtype=torch.float32, device=dev)
    f = _frames(frames)
    so = _soptions(options, None, 0)
    check(lib().llsm_b200_synthesize_harmonics(ctx._h, C.byref(conf), C.byref(f),
                                               C.byref(so) if with_options else None,
                                               _ptr(out), int(nsamp), out.shape[1]))
    return out


OUT_KEYS = ("f0", "nhar", "ampl", "phse", "psd", "psdres", "edc", "enhar", "eampl", "ephse")


def _aoptions(options):
    """llsm_create_aoptions defaults that matter on this path (reference layer0.c:27-43)."""
    a = abi.AOptions()
    options = options or {}
    a.f0_refine = int(options.get("f0_refine", 1))
    a.hm_method = int(options.get("hm_method", 1))      # LLSM_AOPTION_HMCZT
    a.rel_winsize = float(options.get("rel_winsize", 4.0))
    return a


def analysis_shapes(conf):
    B, F, n = conf.nutt, conf.nfrm, conf.nchannel
    return {"f0": ((B, F), "f"), "nhar": ((B, F), "i"), "ampl": ((B, F, conf.maxnhar), "f"),
            "phse": ((B, F, conf.maxnhar), "f"), "psd": ((B, F, conf.npsd), "f"),
            "psdres": ((B, F, conf.npsd), "f"), "edc": ((B, F, n), "f"), "enhar": ((B, F, n), "i"),
            "eampl": ((B, F, n, conf.maxnhar_e), "f"), "ephse": ((B, F, n, conf.maxnhar_e), "f")}


def analyze_l0(ctx, conf, x, f0, options=None, want_residual=False, out=None):
    """llsm_analyze (layer0.c:478-511) for a batch of waveforms held in CUDA tensors.
    x: [B][nx] float32 CUDA tensor, f0: [B][F] (copied; the refined track is returned, as the
    reference overwrites the caller's f0). Returns dict of frame tensors (+ x_res). out: reuse the arrays of an
    earlier call (every row of every array is rewritten)."""
    import torch
    dev = x.device
    if out is None:
        out = {}
        for k, (shape, kind) in analysis_shapes(conf).items():
            out[k] = torch.zeros(shape, dtype=torch.float32 if kind == "f" else torch.int32, device=dev)
    out["f0"].copy_(f0)
    fo = abi.FramesOut()
    for k in OUT_KEYS:
        setattr(fo, k, _ptr(out[k]))
    xr = torch.empty_like(x) if want_residual else None
    a = _aoptions(options)
    check(lib().llsm_b200_analyze_l0(ctx._h, C.byref(conf), C.byref(a), _ptr(x), x.shape[1], x.stride(0),
                                     C.byref(fo), _ptr(xr)))
    if want_residual:
        out["x_res"] = xr
    return out


def analyze_l0_host(ctx, conf, x, f0, options=None, want_residual=False):
    """Same through host numpy buffers."""
    out = {}
    for k, (shape, kind) in analysis_shapes(conf).items():
        out[k] = np.zeros(shape, np.float32 if kind == "f" else np.int32)
    out["f0"][...] = f0
    fo = abi.FramesOut()
    for k in OUT_KEYS:
        setattr(fo, k, _ptr(out[k]))
    x = np.ascontiguousarray(x, np.float32)
    xr = np.empty_like(x) if want_residual else None
    a = _aoptions(options)
    check(lib().llsm_b200_analyze_l0_host(ctx._h, C.byref(conf), C.byref(a), _ptr(x), x.shape[1], x.shape[1],
                                          C.byref(fo), _ptr(xr)))
    if want_residual:
        out["x_res"] = xr
    return out


def anasynth_host(ctx, conf, x, f0, options=None, soptions=None, white=None, seed=0, phase_ops=0, out=None,
                  f0_refined=None):
    """llsm_analyze -> [phasesync_rps / phasepropagate] -> llsm_synthesize in one call (test/test-layer0-anasynth.c:40-66):
    host (numpy / pinned torch CPU) waveforms in, host waveforms out, the chunk stays on the device.
    out: dict with any of y / y_sin / y_noise ([B][>= ny] float32); default: y only."""
    ny = output_length(conf.nfrm, conf.thop, conf.fs)
    if out is None:
        out = {"y": np.empty((conf.nutt, ny), np.float32)}
    o = abi.Output()
    o.y, o.y_sin, o.y_noise = _ptr(out.get("y")), _ptr(out.get("y_sin")), _ptr(out.get("y_noise"))
    o.stride = next(v for v in out.values() if v is not None).shape[1]
    a = _aoptions(options)
    so = _soptions(soptions, white, seed)
    check(lib().llsm_b200_anasynth_host(ctx._h, C.byref(conf), C.byref(a), C.byref(so), _ptr(x), int(x.shape[1]),
                                        int(x.shape[1]), _ptr(f0), _ptr(f0_refined), int(phase_ops), C.byref(o)))
    return out


def tolayer1(ctx, conf, frames, nfft):
    """llsm_chunk_tolayer1 (layer1.c:129-149) on CUDA tensors: returns dict(rd, vtmagn, vsphse, nvs)."""
    import torch
    dev = frames["f0"].device
    B, F, nspec = conf.nutt, conf.nfrm, nfft // 2 + 1
    out = {"rd": torch.zeros((B, F), dtype=torch.float32, device=dev),
           "vtmagn": torch.zeros((B, F, nspec), dtype=torch.float32, device=dev),
           "vsphse": torch.zeros((B, F, conf.maxnhar), dtype=torch.float32, device=dev),
           "nvs": torch.zeros((B, F), dtype=torch.int32, device=dev)}
    l1 = abi.Layer1()
    l1.rd, l1.vtmagn, l1.vsphse, l1.nvs, l1.nspec = (_ptr(out["rd"]), _ptr(out["vtmagn"]), _ptr(out["vsphse"]),
                                                     _ptr(out["nvs"]), nspec)
    f = _frames(frames)
    check(lib().llsm_b200_tolayer1(ctx._h, C.byref(conf), C.byref(f), int(nfft), C.byref(l1)))
    return out


def tolayer0(ctx, conf, f0, layer1, nfrm_utt=None):
    """llsm_chunk_tolayer0 (layer1.c:151-201) on CUDA tensors: returns dict(nhar, ampl, phse)."""
    import torch
    dev = f0.device
    B, F = conf.nutt, conf.nfrm
    out = {"nhar": torch.zeros((B, F), dtype=torch.int32, device=dev),
           "ampl": torch.zeros((B, F, conf.maxnhar), dtype=torch.float32, device=dev),
           "phse": torch.zeros((B, F, conf.maxnhar), dtype=torch.float32, device=dev)}
    l1 = abi.Layer1()
    l1.rd, l1.vtmagn, l1.vsphse, l1.nvs = (_ptr(layer1["rd"]), _ptr(layer1["vtmagn"]), _ptr(layer1["vsphse"]),
                                           _ptr(layer1["nvs"]))
    l1.nspec = layer1["vtmagn"].shape[-1]
    check(lib().llsm_b200_tolayer0(ctx._h, C.byref(conf), _ptr(nfrm_utt), _ptr(f0), C.byref(l1),
                                   _ptr(out["nhar"]), _ptr(out["ampl"]), _ptr(out["phse"])))
    return out


def synthesize_l1(ctx, conf, frames, layer1, pbpsyn=None, white=None, seed=0, options=None, out=None):
    """llsm_synthesize with use_l1 (layer0.c:148-287) on CUDA tensors. frames: noise model (+ optional
    stored HM); layer1: dict(rd, vtmagn, vsphse, nvs); pbpsyn: [B][F] int32 flags or None."""
    import torch
    ny = output_length(conf.nfrm, conf.thop, conf.fs)
    dev = frames["f0"].device
    if out is None:
        out = {k: torch.empty((conf.nutt, ny), dtype=torch.float32, device=dev) for k in ("y", "y_sin", "y_noise")}
    o = abi.Output()
    o.y, o.y_sin, o.y_noise, o.stride = _ptr(out["y"]), _ptr(out["y_sin"]), _ptr(out["y_noise"]), out["y"].shape[1]
    l1 = abi.Layer1()
    l1.rd, l1.vtmagn, l1.vsphse, l1.nvs = (_ptr(layer1["rd"]), _ptr(layer1["vtmagn"]), _ptr(layer1["vsphse"]),
                                           _ptr(layer1["nvs"]))
    l1.nspec = layer1["vtmagn"].shape[-1]
    f = _frames(frames)
    so = _soptions(options, white, seed)
    check(lib().llsm_b200_synthesize_l1(ctx._h, C.byref(conf), C.byref(f), C.byref(l1), _ptr(pbpsyn), C.byref(so), C.byref(o)))
    return out


def synthesize_l0_shard(ctx, conf, frames, frame_lo, frame_hi, white=None, seed=0, options=None, out=None):
    """Partial synthesis of frames [frame_lo, frame_hi) (see include/llsm_b200.h, frame-range sharding)."""
    import torch
    ny = output_length(conf.nfrm, conf.thop, conf.fs)
    dev = frames["f0"].device
    if out is None:
        out = {k: torch.empty((conf.nutt, ny), dtype=torch.float32, device=dev) for k in ("y", "y_sin", "y_noise")}
    o = abi.Output()
    o.y, o.y_sin, o.y_noise, o.stride = _ptr(out["y"]), _ptr(out["y_sin"]), _ptr(out["y_noise"]), out["y"].shape[1]
    f = _frames(frames)
    so = _soptions(options, white, seed)
    check(lib().llsm_b200_synthesize_l0_shard(ctx._h, C.byref(conf), C.byref(f), C.byref(so), C.byref(o),
                                              int(frame_lo), int(frame_hi)))
    return out


class RtSynth:
    """llsm_rtsynth_buffer (llsmrt.h:33-56) for a batch of conf.nutt streams sharing fs / thop.

    feed() takes the next nfeed frames of every stream ([nutt][nfeed][..] arrays) and returns the
    samples they release, (periodic, aperiodic) as llsm_rtsynth_buffer_fetch_decomposed yields them.
    CUDA tensors in -> CUDA tensors out; numpy / CPU tensors in -> numpy out (copies inside)."""

    def __init__(self, ctx, conf, white=None, seed=0, options=None, nspec=None, host_tracker=False):
        """nspec (LLSM_CONF_NSPEC) given: a layer-1 stream (soptions.use_l1), fed with layer1= / pbpsyn=;
        host_tracker keeps the pulse tracker on the host so that a per-pulse hook can run (numpy feeds only)."""
        self.ctx, self.conf = ctx, conf
        so = _soptions(options, white, seed)
        on_host = 0 if (white is None or hasattr(white, "data_ptr") and white.is_cuda) else 1
        h = C.c_void_p()
        self.nspec = nspec
        if nspec is None:
            check(lib().llsm_b200_rt_create(ctx._h, C.byref(conf), C.byref(so), on_host, C.byref(h)))
        else:
            check(lib().llsm_b200_rt_create_l1(ctx._h, C.byref(conf), C.byref(so), on_host, int(nspec),
                                               1 if host_tracker else 0, C.byref(h)))
        self._h = h

    @staticmethod
    def template_length(fs):
        return lib().llsm_b200_rt_template_length(C.c_float(fs))

    @property
    def latency(self):
        return lib().llsm_b200_rt_latency(self._h)

    def output_length(self, nfeed=1):
        return lib().llsm_b200_rt_output_length(self._h, int(nfeed))

    def clear(self):
        check(lib().llsm_b200_rt_clear(self._h))

    def feed(self, frames, nfeed=1, layer1=None, pbpsyn=None, hook=None, user=None):
        n = self.output_length(nfeed)
        f = _frames({k: v for k, v in frames.items() if k != "nfrm_utt"})
        got = C.c_int(0)
        f0 = frames["f0"]
        l1 = None
        if self.nspec is not None:
            l1 = abi.Layer1()
            l1.rd, l1.vtmagn, l1.vsphse, l1.nvs = (_ptr(layer1["rd"]), _ptr(layer1["vtmagn"]), _ptr(layer1["vsphse"]),
                                                   _ptr(layer1["nvs"]))
            l1.nspec = layer1["vtmagn"].shape[-1]
        if hasattr(f0, "is_cuda") and f0.is_cuda:
            import torch
            p = torch.empty((self.conf.nutt, n), dtype=torch.float32, device=f0.device)
            ap = torch.empty_like(p)
            if l1 is None:
                check(lib().llsm_b200_rt_feed(self._h, C.byref(f), int(nfeed), _ptr(p), _ptr(ap), n, C.byref(got)))
            else:
                check(lib().llsm_b200_rt_feed_l1(self._h, C.byref(f), C.byref(l1), _ptr(pbpsyn), int(nfeed),
                                                 _ptr(p), _ptr(ap), n, C.byref(got)))
        else:
            p = np.empty((self.conf.nutt, n), np.float32); ap = np.empty_like(p)
            if l1 is None:
                check(lib().llsm_b200_rt_feed_host(self._h, C.byref(f), int(nfeed), _ptr(p), _ptr(ap), n, C.byref(got)))
            else:
                check(lib().llsm_b200_rt_feed_l1_host(self._h, C.byref(f), C.byref(l1), _ptr(pbpsyn), int(nfeed),
                                                      hook, user, _ptr(p), _ptr(ap), n, C.byref(got)))
        assert got.value == n
        return p, ap

    def close(self):
        if getattr(self, "_h", None):
            lib().llsm_b200_rt_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _phase_op(fn, ctx, conf, frames, layer1, arg):
    fo = abi.FramesOut()
    for k in ("f0", "nhar", "phse", "enhar", "ephse"):
        setattr(fo, k, _ptr(frames[k]))
    l1 = None
    if layer1 is not None:
        l1 = abi.Layer1()
        l1.vsphse, l1.nvs = _ptr(layer1["vsphse"]), _ptr(layer1["nvs"])
    check(fn(ctx._h, C.byref(conf), _ptr(frames.get("nfrm_utt")), C.byref(fo),
             C.byref(l1) if l1 is not None else None, int(arg)))


def chunk_phasepropagate(ctx, conf, frames, layer1=None, sign=1):
    """llsm_chunk_phasepropagate (layer0.c:694-706) on CUDA tensors, in place: frames["phse"], frames["ephse"]
    and layer1["vsphse"] are shifted by the running phase 2 pi thop sign cumsum(f0)."""
    _phase_op(lib().llsm_b200_chunk_phasepropagate, ctx, conf, frames, layer1, sign)


def chunk_phasesync_rps(ctx, conf, frames, layer1=None, layer1_based=0):
    """llsm_chunk_phasesync_rps (layer0.c:687-692) on CUDA tensors, in place: every frame is shifted so that its
    first harmonic (or first source harmonic when layer1_based) has zero phase."""
    _phase_op(lib().llsm_b200_chunk_phasesync_rps, ctx, conf, frames, layer1, layer1_based)


def coder_encode(ctx, conf, f0, psd, layer1, order_spec=64, order_bap=5, nfrm_utt=None):
    """llsm_coder_encode (coder.c:85-163) for every frame, CUDA tensors: f0 [B][F], psd [B][F][npsd], layer1 =
    dict(rd, vtmagn) -> [B][F][order_spec + order_bap + 3]."""
    import torch
    dim = order_spec + order_bap + 3
    enc = torch.zeros((conf.nutt, conf.nfrm, dim), dtype=torch.float32, device=f0.device)
    l1 = abi.Layer1()
    l1.rd, l1.vtmagn, l1.nspec = _ptr(layer1["rd"]), _ptr(layer1["vtmagn"]), layer1["vtmagn"].shape[-1]
    check(lib().llsm_b200_coder_encode(ctx._h, C.byref(conf), _ptr(nfrm_utt), _ptr(f0), _ptr(psd), C.byref(l1),
                                       int(order_spec), int(order_bap), _ptr(enc)))
    return enc


def coder_decode(ctx, conf, enc, nspec, order_spec=64, order_bap=5, use_layer1=True, nfrm_utt=None):
    """llsm_coder_decode_layer1 / _layer0 (coder.c:165-292) for every frame, CUDA tensors. Returns dict(f0, rd, psd,
    nhar) plus vtmagn, vsphse (layer 1) or ampl, phse (layer 0)."""
    import torch
    dev = enc.device
    B, F = conf.nutt, conf.nfrm
    o = {"f0": torch.zeros((B, F), dtype=torch.float32, device=dev), "rd": torch.zeros((B, F), dtype=torch.float32, device=dev),
         "psd": torch.zeros((B, F, conf.npsd), dtype=torch.float32, device=dev),
         "nhar": torch.zeros((B, F), dtype=torch.int32, device=dev)}
    fo, l1 = abi.FramesOut(), abi.Layer1()
    fo.f0, fo.psd, fo.nhar = _ptr(o["f0"]), _ptr(o["psd"]), _ptr(o["nhar"])
    l1.rd, l1.nspec = _ptr(o["rd"]), int(nspec)
    if use_layer1:
        o["vtmagn"] = torch.zeros((B, F, nspec), dtype=torch.float32, device=dev)
        o["vsphse"] = torch.zeros((B, F, conf.maxnhar), dtype=torch.float32, device=dev)
        l1.vtmagn, l1.vsphse = _ptr(o["vtmagn"]), _ptr(o["vsphse"])
    else:
        o["ampl"] = torch.zeros((B, F, conf.maxnhar), dtype=torch.float32, device=dev)
        o["phse"] = torch.zeros((B, F, conf.maxnhar), dtype=torch.float32, device=dev)
        fo.ampl, fo.phse = _ptr(o["ampl"]), _ptr(o["phse"])
    check(lib().llsm_b200_coder_decode(ctx._h, C.byref(conf), _ptr(nfrm_utt), _ptr(enc), int(order_spec), int(order_bap),
                                       1 if use_layer1 else 0, C.byref(fo), C.byref(l1)))
    return o


def stretch_map(nfrm, nfrm_new):
    """The uniform time map of test/demo-stretch.c:170-175: numpy (base int32, ratio float32, residx int32)."""
    base, ratio, res = np.zeros(nfrm_new, np.int32), np.zeros(nfrm_new, np.float32), np.zeros(nfrm_new, np.int32)
    check(lib().llsm_b200_stretch_map(int(nfrm), int(nfrm_new), base.ctypes.data, ratio.ctypes.data, res.ctypes.data))
    return base, ratio, res


def frames_stretch(ctx, conf, frames, layer1, base, ratio, residx=None):
    """Frame interpolation / time-stretch of a layer-1 batch (test/demo-stretch.c:16-129,169-185), CUDA tensors.
    frames: dict keyed as FRAME_KEYS, layer1: dict(rd, vtmagn, vsphse, nvs); base / ratio / residx: [nfrm_new] (one map
    for the batch) or [B][nfrm_new]. Returns (conf_new, frames_new, layer1_new)."""
    import torch
    dev = frames["f0"].device
    per_utt = base.dim() == 2
    Fn = base.shape[-1]
    B, n = conf.nutt, conf.nchannel
    nspec = layer1["vtmagn"].shape[-1]
    z = lambda shape, dt=torch.float32: torch.empty(shape, dtype=dt, device=dev)  # noqa: E731  (every row is written)
    fo = {"f0": z((B, Fn)), "psd": z((B, Fn, conf.npsd)), "edc": z((B, Fn, n)), "enhar": z((B, Fn, n), torch.int32),
          "eampl": z((B, Fn, n, conf.maxnhar_e)), "ephse": z((B, Fn, n, conf.maxnhar_e))}
    if frames.get("psdres") is not None:
        fo["psdres"] = z((B, Fn, conf.npsd))
    if all(frames.get(k) is not None for k in ("nhar", "ampl", "phse")):
        fo["nhar"], fo["ampl"], fo["phse"] = z((B, Fn), torch.int32), z((B, Fn, conf.maxnhar)), z((B, Fn, conf.maxnhar))
    lo = {"rd": z((B, Fn)), "vtmagn": z((B, Fn, nspec)), "vsphse": z((B, Fn, conf.maxnhar)), "nvs": z((B, Fn), torch.int32)}
    s, d = _frames(frames), abi.FramesOut()
    for k, v in fo.items():
        setattr(d, k, _ptr(v))
    sl, dl = abi.Layer1(), abi.Layer1()
    for k in ("rd", "vtmagn", "vsphse", "nvs"):
        setattr(sl, k, _ptr(layer1[k]))
        setattr(dl, k, _ptr(lo[k]))
    sl.nspec = dl.nspec = int(nspec)
    check(lib().llsm_b200_frames_stretch(ctx._h, C.byref(conf), C.byref(s), C.byref(sl), int(Fn), _ptr(base), _ptr(ratio),
                                         _ptr(residx), 1 if per_utt else 0, C.byref(d), C.byref(dl)))
    conf_new = abi.make_conf(B, Fn, conf.maxnhar, conf.maxnhar_e, conf.npsd, n, conf.fs, conf.thop,
                             list(conf.chanfreq)[:max(n - 1, 0)], conf.lip_radius)
    return conf_new, fo, lo


def frames_to_blob(conf, frames):
    """Serialise a batch (dict of numpy arrays keyed as FRAME_KEYS) into one relocatable buffer (numpy uint8):
    llsm_b200_frames_pack. What parallel.py scatters and what a file would hold."""
    f = _frames({k: (np.ascontiguousarray(v) if v is not None else None) for k, v in frames.items()})
    n = lib().llsm_b200_frames_blob_size(C.byref(conf), C.byref(f))
    blob = np.zeros(n, np.uint8)
    check(lib().llsm_b200_frames_pack(C.byref(conf), C.byref(f), blob.ctypes.data, n))
    return blob


def blob_to_frames(blob):
    """llsm_b200_frames_unpack: (conf, dict of numpy views into the blob)."""
    blob = np.ascontiguousarray(blob, np.uint8)
    conf, f = abi.Conf(), abi.Frames()
    check(lib().llsm_b200_frames_unpack(blob.ctypes.data, blob.size, C.byref(conf), C.byref(f)))
    B, F, n = conf.nutt, conf.nfrm, conf.nchannel
    shapes = {"nfrm_utt": ((B,), np.int32), "f0": ((B, F), np.float32), "nhar": ((B, F), np.int32),
              "ampl": ((B, F, conf.maxnhar), np.float32), "phse": ((B, F, conf.maxnhar), np.float32),
              "psd": ((B, F, conf.npsd), np.float32), "psdres": ((B, F, conf.npsd), np.float32),
              "edc": ((B, F, n), np.float32), "enhar": ((B, F, n), np.int32),
              "eampl": ((B, F, n, conf.maxnhar_e), np.float32), "ephse": ((B, F, n, conf.maxnhar_e), np.float32)}
    out = {}
    base = blob.ctypes.data
    for k in FRAME_KEYS:
        p = getattr(f, k)
        if not p:
            out[k] = None
            continue
        shape, dt = shapes[k]
        cnt = int(np.prod(shape))
        out[k] = np.frombuffer(blob, dtype=dt, count=cnt, offset=p - base).reshape(shape)
    return conf, out
