"""Every fast analysis kernel of round 2 against the general kernel it replaced, on the CPU thread emulation: the
selection switches pick the older kernel in a second process and the two analyses of the same
waveform are compared array by array. Much tighter than the parity bars against the oracle (which include the oracle's
own float noise): the pairs compute the same mathematics in a different order."""
import os
import subprocess
import sys
import numpy as np
import pytest
import support as S

HELPER = os.path.join(os.path.dirname(os.path.abspath(__file__)), "helpers", "emu_analyze_dump.py")


def _run(tmp_path, tag, mode, **env):
    S.load_emu()                                             # build once, in this process
    out = str(tmp_path / (tag + ".npz"))
    e = dict(os.environ); e.update(env); e["LLSM_EMU_NOBUILD"] = "1"
    subprocess.check_call([sys.executable, HELPER, out, mode], env=e)
    return np.load(out)


@pytest.mark.parametrize("mode", ["norm", "low"])            # low: f0 50-78 Hz, the frame-list / time-aliased paths
def test_fast_kernels_match_the_general_ones(tmp_path, mode):
    base = _run(tmp_path, "base", mode)
    amp = float(np.abs(base["ampl"]).max()); eamp = float(np.abs(base["eampl"]).max())
    for tag, env, bars in (
        # noise spectra: register-FFT warp kernel vs block-FFT kernel (float-level differences through the Kalman smoother)
        ("ns0", {"LLSM_NS_VARIANT": "0"}, {"psd": 0.02, "psdres": 0.02, "ampl": 0.0, "eampl": 0.0, "x_res": 0.0}),
        # main pass: FP16-split tensor-core products vs direct FP32 DFT
        ("dft0", {"LLSM_DFT_VARIANT": "0"}, {"ampl": 5e-6 * amp, "x_res": 5e-6 * amp, "psd": 0.02, "eampl": 1e-4 * eamp}),
        # envelope pass: bulk-copied segments, a warp per frame vs one CTA per frame
        ("env0", {"LLSM_ENV_VARIANT": "0"}, {"eampl": 1e-6 * eamp, "edc": 1e-6 * float(np.abs(base["edc"]).max()), "ampl": 0.0, "psd": 0.0}),
        # sub-band filter: shared-memory resident vs streaming
        ("iir0", {"LLSM_IIR_VARIANT": "0"}, {"eampl": 1e-6 * eamp, "edc": 1e-6 * float(np.abs(base["edc"]).max()), "ampl": 0.0, "psd": 0.0}),
    ):
        other = _run(tmp_path, tag, mode, **env)
        assert np.array_equal(other["nhar"], base["nhar"]) and np.array_equal(other["enhar"], base["enhar"]), tag
        assert np.array_equal(other["f0"], base["f0"]), tag
        for k, bar in bars.items():
            d = float(np.abs(other[k].astype(np.float64) - base[k]).max())
            assert d <= bar, (tag, k, d, bar)
