"""Timing of BASELINE configs 3 and 4 on one GPU (device-resident, CUDA events), next to the reference
sources on one host thread for a bounded sample.
  C3: batch PbP synthesis (use_l1, every voiced frame PBPSYN = 1), 256 harmonics, f0 in [60, 86] Hz
  C4: layer0 -> layer1 (Rd fit + spectral envelope, nfft 2048) over 4096 frames of the C2 workload
Prints one JSON line per configuration."""
import json, sys, time, argparse, ctypes as C
import numpy as np
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import libllsm2_b200 as L
from libllsm2_b200.synthetic import synth_frames

ap = argparse.ArgumentParser()
ap.add_argument("--batch3", type=int, default=1024)
ap.add_argument("--nfrm3", type=int, default=100)
ap.add_argument("--no-cpu", action="store_true")
a = ap.parse_args()
ctx = L.Context(0)


def ev_time(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def rep(fr, B):
    n = fr["f0"].shape[0]
    r = (B + n - 1) // n
    return {k: (np.ascontiguousarray(np.concatenate([v] * r, 0)[:B]) if v is not None else None) for k, v in fr.items()}


# ---------------------------------------------------------------- C4
fr, conf = synth_frames(16, 256, seed=3)            # 4096 frames
d = {k: (torch.from_numpy(v).cuda() if v is not None else None) for k, v in fr.items()}
ms = ev_time(lambda: L.tolayer1(ctx, conf, d, 2048))
out = {"config": "C4 layer0->layer1, 4096 frames (16 x 256), nfft 2048", "ms": ms, "frames_per_s": 4096 / ms * 1e3}
if not a.no_cpu:
    import support as S
    t0 = time.perf_counter(); S.ref_tolayer1({k: v[:2] if v is not None else None for k, v in fr.items()},
                                             L.abi.make_conf(2, 256, conf.maxnhar, conf.maxnhar_e, conf.npsd, conf.nchannel, conf.fs, conf.thop), 2048)
    out["cpu_frames_per_s_1thread_parity_build"] = 2 * 256 / (time.perf_counter() - t0)
print(json.dumps(out), flush=True)

# larger batch for throughput
fr, conf = synth_frames(16, 400, seed=3)
frb = rep(fr, 256); conf.nutt = 256
d = {k: (torch.from_numpy(v).cuda() if v is not None else None) for k, v in frb.items()}
ms = ev_time(lambda: L.tolayer1(ctx, conf, d, 2048), iters=3)
print(json.dumps({"config": "layer0->layer1, 256 x 400 frames", "ms": ms, "frames_per_s": 256 * 400 / ms * 1e3}), flush=True)

# ---------------------------------------------------------------- coder
# 64 mel-cepstral numbers + 5 band aperiodicities (test/test-coder.c:31) on the 256 x 400 layer-1 frames above
l1c = L.tolayer1(ctx, conf, d, 2048)
enc = L.coder_encode(ctx, conf, d["f0"], d["psd"], l1c, 64, 5)
nfr = 256 * 400
ms_e = ev_time(lambda: L.coder_encode(ctx, conf, d["f0"], d["psd"], l1c, 64, 5), iters=3)
ms_1 = ev_time(lambda: L.coder_decode(ctx, conf, enc, 1025, 64, 5, True), iters=3)
ms_0 = ev_time(lambda: L.coder_decode(ctx, conf, enc, 1025, 64, 5, False), iters=3)
out = {"config": "coder, 256 x 400 frames, order_spec 64, order_bap 5, nspec 1025",
       "encode_frames_per_s": nfr / ms_e * 1e3, "decode_layer1_frames_per_s": nfr / ms_1 * 1e3,
       "decode_layer0_frames_per_s": nfr / ms_0 * 1e3, "ms": [ms_e, ms_1, ms_0]}
if not a.no_cpu:
    import support as S
    c2 = L.abi.make_conf(1, 64, conf.maxnhar, conf.maxnhar_e, conf.npsd, conf.nchannel, conf.fs, conf.thop)
    l1h = {k: v[:1, :64].cpu().numpy().copy() for k, v in l1c.items()}
    t0 = time.perf_counter()
    e = S.ref_coder_encode(frb["f0"][:1, :64].copy(), frb["psd"][:1, :64].copy(), l1h, c2, 64, 5)
    t1 = time.perf_counter()
    S.ref_coder_decode(e, c2, 1025, 64, 5, 1)
    t2 = time.perf_counter()
    out["cpu_1thread_parity_build_frames_per_s"] = {"encode": 64 / (t1 - t0), "decode_layer1": 64 / (t2 - t1),
        "note": "the oracle's ddct is a direct O(n^2) sum, the reference's is Ooura's O(n log n): not a fair CPU baseline"}
print(json.dumps(out), flush=True)

# ---------------------------------------------------------------- C3
F = a.nfrm3
fr, conf = synth_frames(16, F, seed=5, nhar=256, maxnhar=256, f0_lo=60, f0_hi=86)
frb = rep(fr, a.batch3); conf.nutt = a.batch3
d = {k: (torch.from_numpy(v).cuda() if v is not None else None) for k, v in frb.items()}
l1 = L.tolayer1(ctx, conf, d, 2048)
pbp = torch.ones((a.batch3, F), dtype=torch.int32, device="cuda")
dn = dict(d); dn["nhar"] = dn["ampl"] = dn["phse"] = None          # HM removed: everything from layer 1
l0 = ctx.launches
ms = ev_time(lambda: L.synthesize_l1(ctx, conf, dn, l1, pbpsyn=pbp, seed=3), iters=3)
print(json.dumps({"config": "C3 PbP synthesis (use_l1, all voiced frames pulse-by-pulse), batch %d x %d frames, 256 harmonics, "
                            "f0 60-86 Hz, device tracker (no effect callbacks)" % (a.batch3, F),
                  "ms": ms, "frames_per_s": a.batch3 * F / ms * 1e3, "launches_per_call": (ctx.launches - l0) // 5}), flush=True)
ms = ev_time(lambda: L.tolayer1(ctx, conf, d, 2048), iters=3)
print(json.dumps({"config": "layer0->layer1 on the C3 frames (256 harmonics)", "ms": ms,
                  "frames_per_s": a.batch3 * F / ms * 1e3}), flush=True)
if not a.no_cpu:
    import support as S
    c1 = L.abi.make_conf(1, F, conf.maxnhar, conf.maxnhar_e, conf.npsd, conf.nchannel, conf.fs, conf.thop)
    one = {k: v[:1] if v is not None else None for k, v in fr.items()}
    t0 = time.perf_counter(); S.ref_synthesize_l1(one, c1, np.ones((1, F), np.int32), seed=9)
    print(json.dumps({"config": "C3 reference (tolayer1 + llsm_synthesize use_l1), 1 utterance, 1 thread, parity build",
                      "frames_per_s": F / (time.perf_counter() - t0)}), flush=True)

# streaming PbP: S streams, K frames per feed, device tracker
S_, K = 1024, 8
fr, conf = synth_frames(16, K, seed=5, nhar=256, maxnhar=256, f0_lo=60, f0_hi=86)
frb = rep(fr, S_); conf.nutt = S_
d = {k: (torch.from_numpy(v).cuda() if v is not None and k != "nfrm_utt" else None) for k, v in frb.items()}
l1 = L.tolayer1(ctx, conf, d, 2048)
pbp = torch.ones((S_, K), dtype=torch.int32, device="cuda")
rt = L.RtSynth(ctx, conf, seed=5, nspec=1025)
dn = dict(d); dn["nhar"] = dn["ampl"] = dn["phse"] = None
ms = ev_time(lambda: rt.feed(dn, K, layer1=l1, pbpsyn=pbp), iters=5) / K
print(json.dumps({"config": "C3 streaming (llsmrt use_l1, PbP engaged), %d streams, %d frames per feed" % (S_, K),
                  "ms_per_frame_step": ms, "frames_per_s": S_ / ms * 1e3}), flush=True)
