"""Timing of the frame interpolation / time-stretch operation (test/demo-stretch.c:16-129,
169-185) on one GPU: a batch of layer-1 frames stretched to twice its length, device-resident, CUDA events; next to
the reference's own interp_llsm_frame on one host thread for one utterance. Prints one JSON line."""
import json, sys, time, argparse
import numpy as np
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import libllsm2_b200 as L
from libllsm2_b200.synthetic import synth_frames

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--nfrm", type=int, default=400)
ap.add_argument("--no-cpu", action="store_true")
a = ap.parse_args()
ctx = L.Context(0)
B, F = a.batch, a.nfrm

fr, conf = synth_frames(16, F, seed=3)
r = (B + 15) // 16
frb = {k: (np.ascontiguousarray(np.concatenate([v] * r, 0)[:B]) if v is not None else None) for k, v in fr.items()}
conf.nutt = B
d = {k: (torch.from_numpy(v).cuda() if v is not None else None) for k, v in frb.items()}
l1 = L.tolayer1(ctx, conf, d, 2048)
L.chunk_phasepropagate(ctx, conf, d, l1, sign=-1)
base, ratio, res = L.stretch_map(F, 2 * F)
tb, tr, ts = (torch.from_numpy(x).cuda() for x in (base, ratio, res))


def run():
    return L.frames_stretch(ctx, conf, d, l1, tb, tr, ts)


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 10
e0.record()
for _ in range(iters):
    run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
# algorithmic bytes of one output frame: two source rows read and one written (f0, rd, nvs, vsphse, vtmagn, psd, per
# channel edc + enhar + envelope harmonics), the PSDRES row and the layer-0 harmonics copied (read + write)
row = 4.0 * (3 + conf.maxnhar + 1025 + conf.npsd + conf.nchannel * (2 + 2 * conf.maxnhar_e))
byt = (3 * row + 8.0 * conf.npsd + 2 * 4.0 * (1 + 2 * conf.maxnhar)) * B * 2 * F
out = {"config": "frame interpolation, %d x %d -> %d frames, nspec 1025, %d harmonics" % (B, F, 2 * F, conf.maxnhar),
       "ms": ms, "out_frames_per_s": B * 2 * F / ms * 1e3, "algorithmic_bytes": byt, "algorithmic_GB_per_s": byt / ms / 1e6}
if not a.no_cpu:
    import support as S
    c1 = L.abi.make_conf(1, F, conf.maxnhar, conf.maxnhar_e, conf.npsd, conf.nchannel, conf.fs, conf.thop)
    one = {k: (v[:1].cpu().numpy().copy() if v is not None else None) for k, v in d.items()}
    l1h = {k: v[:1].cpu().numpy().copy() for k, v in l1.items()}
    t0 = time.perf_counter()
    S.ref_stretch(one, c1, l1h, base, ratio, res)
    out["cpu_1thread_parity_build_out_frames_per_s"] = 2 * F / (time.perf_counter() - t0)
print(json.dumps(out), flush=True)
