"""Helper of tests/test_gpu_variants.py: one analysis on the GPU, outputs saved to argv[1] (a process per kernel-selection
setting: the LLSM_*_VARIANT switches are read once per process). argv[2]: frames per utterance."""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import libllsm2_b200 as L

d = np.load(sys.argv[3])
conf = L.abi.make_conf(int(d["B"]), int(d["F"]), int(d["maxnhar"]), int(d["maxnhar_e"]), int(d["npsd"]), int(d["nch"]),
                       float(d["fs"]), float(d["thop"]))
ctx = L.Context(0)
out = L.analyze_l0(ctx, conf, torch.from_numpy(d["x"]).cuda(), torch.from_numpy(d["f0"]).cuda(), want_residual=True,
                   options={"hm_method": 1})
torch.cuda.synchronize()
np.savez(sys.argv[1], **{k: v.cpu().numpy() for k, v in out.items()})
ctx.close()
