"""Analysis (llsm_analyze) timing on the device: B utterances x 400 frames, input = the waveform synthesised
from the C2-shaped frames. Prints one JSON line per (batch, method). Use under ncu for the per-kernel list:
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv python tools/ana_bench.py --once"""
import json, sys, time, argparse
import numpy as np
import torch
sys.path.insert(0, ".")
import libllsm2_b200 as L
from libllsm2_b200.synthetic import synth_frames

ap = argparse.ArgumentParser()
ap.add_argument("--once", action="store_true")
ap.add_argument("--batch", type=int, default=0)
a = ap.parse_args()
ctx = L.Context(0)
for B in ([a.batch] if a.batch else ([64] if a.once else [16, 128, 512])):
    fr, conf = synth_frames(min(B, 16), 400, seed=3)
    rep = (B + 15) // 16
    frt = {k: (np.ascontiguousarray(np.concatenate([v] * rep, 0)[:B]) if v is not None else None) for k, v in fr.items()}
    conf.nutt = B
    d = {k: (torch.from_numpy(v).cuda() if v is not None else None) for k, v in frt.items()}
    y = L.synthesize_l0(ctx, conf, d, seed=11)["y"]
    for method in ((1,) if a.once else (1, 0)):
        opts = {"hm_method": method}
        iters = 1 if a.once else 5
        if not a.once:
            L.analyze_l0(ctx, conf, y, d["f0"], options=opts)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launches
        e0.record()
        for _ in range(iters):
            o = L.analyze_l0(ctx, conf, y, d["f0"], options=opts)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print(json.dumps({"batch": B, "nfrm": 400, "hm_method": "czt" if method else "pp", "ms": ms,
                          "frames_per_s": B * 400 / ms * 1e3, "launches": (ctx.launches - l0) // iters}), flush=True)
