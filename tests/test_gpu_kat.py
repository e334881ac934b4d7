"""The reference's own known-answer test for harmonic analysis (test/test-dsputils.c:44-133: a three-harmonic chirp,
amplitude error < 0.01, phase-advance error < 0.1 rad, both estimators) run against the CUDA analysis path itself,
not just against the oracle that pins it (tests/test_oracle_kat.py)."""
import math
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("method", [0, 1])   # LLSM_AOPTION_HMPP, LLSM_AOPTION_HMCZT
def test_chirp_harmonic_analysis_on_device(method):
    import torch
    import libllsm2_b200 as L
    from libllsm2_b200 import abi
    nx, fs, thop = 100000, 20000.0, 0.005
    nfrm = int(math.floor(np.float32(nx) / np.float32(fs) / np.float32(thop)))
    center = np.round(np.arange(nfrm) * np.float32(thop) * np.float32(fs)).astype(int)
    rate = center.astype(np.float32) / nx
    f0 = (100 + 100 * rate).astype(np.float32)
    i = np.arange(nx)
    ph = np.cumsum((100 + 100 * i / nx) / fs * 2 * 3.1415927)
    x = ((i / nx) * np.sin(ph) + 0.5 * np.sin(2 * ph) + 0.25 * np.sin(3 * ph)).astype(np.float32)
    conf = abi.make_conf(1, nfrm, 3, 2, 64, 4, fs, thop, [2000.0, 4000.0, 8000.0])
    ctx = L.Context(0)
    o = L.analyze_l0(ctx, conf, torch.from_numpy(x).cuda().view(1, -1), torch.from_numpy(f0).cuda().view(1, -1),
                     options={"hm_method": method, "f0_refine": 0})
    torch.cuda.synchronize()
    A = o["ampl"][0].cpu().numpy(); P = o["phse"][0, :, 0].cpu().numpy(); nh = o["nhar"][0].cpu().numpy()
    ctx.close()
    assert np.all(nh == 3)
    s = slice(5, nfrm - 5)
    for k, truth in enumerate([rate, 0.5, 0.25]):
        e = np.zeros(nfrm); e[s] = (A[:, k] - truth)[s]
        assert abs(e.mean()) < 0.01 and e.std() < 0.01, (k, e.mean(), e.std())
    pe = np.zeros(nfrm - 1)
    for t in range(5, nfrm - 5):
        d = P[t] - (P[t - 1] + f0[t] * 2 * 3.1415927 * thop)
        pe[t - 1] = (d + math.pi) % (2 * math.pi) - math.pi
    assert abs(pe.mean()) < 0.1 and pe.std() < 0.1, (pe.mean(), pe.std())
