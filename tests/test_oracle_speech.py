"""The reference's own end-to-end acceptance test (test/test-layer0-anasynth.c:13-83, BASELINE configs[0]) run on
the oracle build: arctic_a0001.wav -> llsm_analyze (HMPP and CZT) -> llsm_synthesize, before and after
llsm_chunk_phasesync_rps + llsm_chunk_phasepropagate, judged with the reference's statistical bars
(test/verify-utils.h: KL divergence < 0.05 on the waveform / first / second difference and on the STFT
distribution, spectral correlation > 0.95). It pins the ciglet shim at the level the reference itself pins ciglet."""
import numpy as np
import pytest

import compat_util as U
import speech_util as SU
import support as S


@pytest.mark.parametrize("method", ["pp", "czt"])
def test_reference_acceptance_on_arctic(method):
    L = U.bind(S.load_ref())
    fx = SU.fixtures()["arctic"]
    assert len(fx["x"]) == 147734 and fx["fs"] == 44100.0 and len(fx["f0"]) == 1154
    r = SU.anasynth(L, fx["x"], fx["fs"], fx["f0"], fx["nhop"], method)
    assert r["chunk"]["nhar"].max() > 100                      # real speech: up to fnyq / f0 harmonics
    for key in ("out1", "out2"):
        y = r[key][0]
        assert np.isfinite(y).all()
        for k in SU.verify_data_distribution(fx["x"], y):
            assert k < 0.05, (key, k)
        cc, k0, k1 = SU.verify_spectral_distribution(fx["x"], y)
        assert cc > 0.95 and k0 < 0.05 and k1 < 0.05, (key, cc, k0, k1)


def test_empirical_kld_selftest():
    """test/verify-utils.h:32-69."""
    rng = np.random.default_rng(0)
    x = rng.normal(1.0, 1.0, 100000)
    k1 = SU.empirical_kld(x, rng.normal(1.0, 1.0, 50000))
    k2 = SU.empirical_kld(x, rng.normal(0.0, 1.0, 50000))
    k3 = SU.empirical_kld(x, rng.normal(1.0, np.sqrt(3.0), 50000))
    k4 = SU.empirical_kld(x, rng.uniform(0, 1, 50000))
    assert abs(k1) < 0.05 and k2 > k1 and k3 > k1 and k4 > k1
    assert abs(k2 - 0.5) < 0.1
