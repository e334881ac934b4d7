#!/usr/bin/env python
"""Time the harmonic-bank kernel alone for each launch variant (LLSM_BANK_VARIANT), fresh process each."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, json; sys.path.insert(0, %r)
import numpy as np, torch
import libllsm2_b200 as L
from libllsm2_b200.synthetic import synth_frames
fr, conf = synth_frames(32, 400, nhar=128, seed=0)
B = 1024
full = {k: (None if v is None else np.ascontiguousarray(np.concatenate([v] * (B // 32), 0))) for k, v in fr.items()}
conf.nutt = B
d = {k: (torch.from_numpy(v).cuda() if v is not None else None) for k, v in full.items()}
ctx = L.Context(0)
ny = L.output_length(conf.nfrm, conf.thop, conf.fs)
ys = torch.empty((B, ny), dtype=torch.float32, device="cuda")
for i in range(3): L.synthesize_harmonics(ctx, conf, d, ny, out=ys)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(10): L.synthesize_harmonics(ctx, conf, d, ny, out=ys)
e1.record(); torch.cuda.synchronize()
print(json.dumps({"ms": e0.elapsed_time(e1) / 10, "checksum": float(ys.double().abs().sum())}))
''' % ROOT
for v in sys.argv[1:] or ["0", "1", "2", "3", "4", "5"]:
    env = dict(os.environ, LLSM_BANK_VARIANT=v)
    out = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True)
    print("variant", v, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-500:])
