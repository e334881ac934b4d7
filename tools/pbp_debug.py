"""Diagnostic: device tracker vs host tracker vs oracle for layer-1 synthesis (run on the GPU box)."""
import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import support as S
import libllsm2_b200 as L
from libllsm2_b200 import abi
from libllsm2_b200._lib import lib, check

B, F = 1, 60
fr, conf = S.synth_frames(B, F, seed=3, nhar=100, maxnhar=100)
pbp = np.ones((B, F), np.int32)
ref, l1 = S.ref_synthesize_l1(fr, conf, pbp, seed=9)
white = S.ref_white_noise(conf, seed=9)
ctx = L.Context(0)
fr2 = dict(fr); fr2["nhar"] = None; fr2["ampl"] = None; fr2["phse"] = None
dev = lambda d: {k: (torch.from_numpy(np.ascontiguousarray(v)).cuda() if v is not None else None) for k, v in d.items()}
out = L.synthesize_l1(ctx, conf, dev(fr2), dev(l1), pbpsyn=torch.from_numpy(pbp).cuda(), white=torch.from_numpy(white).cuda())
torch.cuda.synchronize()
ys_dev = out["y_sin"].cpu().numpy()
# host-tracker path through the host entry with a no-op hook
HOOK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, *([C.POINTER(C.c_float)] * 6))
hook = HOOK(lambda *a: 0)
ny = ref[0].shape[1]
y = np.zeros((B, ny), np.float32); ys = np.zeros_like(y); yn = np.zeros_like(y)
o = abi.Output(); o.y = y.ctypes.data; o.y_sin = ys.ctypes.data; o.y_noise = yn.ctypes.data; o.stride = ny
f = abi.Frames()
for k in ("f0", "psd", "psdres", "edc", "enhar", "eampl", "ephse"):
    setattr(f, k, fr[k].ctypes.data)
s = abi.Layer1(); s.rd = l1["rd"].ctypes.data; s.vtmagn = l1["vtmagn"].ctypes.data; s.vsphse = l1["vsphse"].ctypes.data
s.nvs = l1["nvs"].ctypes.data; s.nspec = l1["vtmagn"].shape[-1]
so = abi.default_soptions(white.ctypes.data, 0)
Lb = lib()
Lb.llsm_b200_synthesize_l1_host.argtypes = [C.c_void_p, C.POINTER(abi.Conf), C.POINTER(abi.Frames), C.POINTER(abi.Layer1), C.c_void_p,
                                            C.POINTER(abi.SOptions), C.POINTER(abi.Output), C.c_void_p, C.c_void_p]
check(Lb.llsm_b200_synthesize_l1_host(ctx._h, C.byref(conf), C.byref(f), C.byref(s), pbp.ctypes.data_as(C.c_void_p), C.byref(so),
                                      C.byref(o), C.cast(hook, C.c_void_p), None))
print("rms ref y_sin", S.rms(ref[1]))
print("device-tracker vs ref", S.rms(ys_dev - ref[1]), " host-tracker vs ref", S.rms(ys - ref[1]), " dev vs host", S.rms(ys_dev - ys))
d = np.abs(ys_dev[0] - ref[1][0])
hop = 220.5
per = [float(np.sqrt((d[int(i * hop):int((i + 1) * hop)] ** 2).mean())) for i in range(F)]
print("per-frame err (device tracker):", ["%.1e" % v for v in per[:30]])
