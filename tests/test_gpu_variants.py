"""Every fast analysis kernel of round 2 against the general kernel it replaced, ON THE GPU, at sizes the CPU emulation
cannot reach: 400- and 1501-frame utterances put the shared-memory sub-band filter on thread-block clusters of two and
eight CTAs (carry through distributed shared memory), which only exist on the device. A process per setting (the
selection switches are read once per process); arrays compared far below the parity bars against the oracle."""
import os
import subprocess
import sys
import numpy as np
import pytest
import support as S

pytestmark = pytest.mark.gpu
HELPER = os.path.join(os.path.dirname(os.path.abspath(__file__)), "helpers", "gpu_analyze_dump.py")


def _run(tmp_path, tag, inp, **env):
    out = str(tmp_path / (tag + ".npz"))
    e = dict(os.environ); e.update(env)
    subprocess.check_call([sys.executable, HELPER, out, "0", inp], env=e)
    return np.load(out)


@pytest.mark.parametrize("F,f0_lo,f0_hi", [(400, 90, 170), (1501, 60, 140)])
def test_fast_kernels_match_the_general_ones_on_gpu(tmp_path, F, f0_lo, f0_hi):
    fr, conf = S.synth_frames(2, F, seed=41 + F, nhar=100, maxnhar=128, f0_lo=f0_lo, f0_hi=f0_hi)
    y, ys, yn = S.ref_synthesize(fr, conf, seed=7)
    inp = str(tmp_path / "in.npz")
    np.savez(inp, x=np.ascontiguousarray(y), f0=fr["f0"], B=conf.nutt, F=conf.nfrm, maxnhar=conf.maxnhar,
             maxnhar_e=conf.maxnhar_e, npsd=conf.npsd, nch=conf.nchannel, fs=conf.fs, thop=conf.thop)
    base = _run(tmp_path, "base", inp)
    amp = float(np.abs(base["ampl"]).max()); eamp = float(np.abs(base["eampl"]).max()); edc = float(np.abs(base["edc"]).max())
    for tag, env, bars in (
        ("iir0", {"LLSM_IIR_VARIANT": "0"}, {"eampl": 2e-6 * eamp, "edc": 2e-6 * edc, "ampl": 0.0, "psd": 0.0}),
        ("env0", {"LLSM_ENV_VARIANT": "0"}, {"eampl": 2e-6 * eamp, "edc": 2e-6 * edc, "ampl": 0.0, "psd": 0.0}),
        ("ns0", {"LLSM_NS_VARIANT": "0"}, {"psd": 0.03, "psdres": 0.05, "ampl": 0.0, "eampl": 0.0}),
        ("dft0", {"LLSM_DFT_VARIANT": "0"}, {"ampl": 5e-6 * amp, "x_res": 5e-6 * amp, "psd": 0.03, "eampl": 1e-4 * eamp}),
        ("serial", {"LLSM_ANA_OVERLAP": "0"}, {"ampl": 0.0, "psd": 0.0, "psdres": 0.0, "eampl": 0.0, "edc": 0.0, "x_res": 0.0}),
    ):
        other = _run(tmp_path, tag, inp, **env)
        assert np.array_equal(other["nhar"], base["nhar"]) and np.array_equal(other["enhar"], base["enhar"]), tag
        assert np.array_equal(other["f0"], base["f0"]), tag
        for k, bar in bars.items():
            d = float(np.abs(other[k].astype(np.float64) - base[k]).max())
            assert d <= bar, (tag, k, d, bar)
