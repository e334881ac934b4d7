"""Streaming synthesizer timing: S streams, one kernel launch per fed frame.
  device arm : frames resident in HBM, K frames per feed call, CUDA events
  host arm   : llsm_b200_rt_feed_host with 1 frame per call (the drop-in's per-frame path), wall clock
Prints one JSON line per configuration."""
import json, sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import libllsm2_b200 as L
from libllsm2_b200.synthetic import synth_frames

ctx = L.Context(0)
for S_, K in ((1, 1), (1, 16), (64, 16), (1024, 16), (4096, 8)):
    fr, conf = synth_frames(min(S_, 16), K, seed=3)
    rep = (S_ + fr["f0"].shape[0] - 1) // fr["f0"].shape[0]
    frt = {k: (np.ascontiguousarray(np.concatenate([v] * rep, 0)[:S_]) if v is not None and k != "nfrm_utt" else None)
           for k, v in fr.items()}
    conf.nutt = S_
    dev = {k: (torch.from_numpy(v).cuda() if v is not None else None) for k, v in frt.items()}
    rt = L.RtSynth(ctx, conf, seed=7)
    for _ in range(3):
        rt.feed(dev, K)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 20
    e0.record()
    for _ in range(iters):
        rt.feed(dev, K)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (iters * K)
    out = {"streams": S_, "frames_per_feed": K, "ms_per_frame_step": ms, "frames_per_s": S_ / ms * 1e3,
           "realtime_factor_per_stream": conf.thop * 1e3 / ms}
    if S_ <= 64:
        one = {k: (np.ascontiguousarray(v[:, :1]) if v is not None else None) for k, v in frt.items()}
        for _ in range(5):
            rt.feed(one, 1)
        t0 = time.perf_counter()
        for _ in range(50):
            rt.feed(one, 1)
        out["host_feed_ms"] = (time.perf_counter() - t0) / 50 * 1e3
    rt.close()
    print(json.dumps(out), flush=True)
