"""Synthetic layer-0 model parameters: used by bench.py, smoke() and the tests.
Pure numpy; nothing here touches the oracle."""
import numpy as np

from . import abi


def synth_frames(B, F, nhar=128, npsd=256, nch=4, nhar_e=4, fs=44100.0, thop=0.005, seed=0,
                 f0_lo=90.0, f0_hi=170.0, maxnhar=None, maxnhar_e=None, unvoiced=0.2, rms=0.05):
    """Synthetic layer-0 frames shaped like SURVEY.md section 8(d) (C2): smooth random-walk F0 with
    unvoiced runs, 1/k^1.2 harmonic amplitudes with log-normal scatter, propagated phases, sloping
    noise PSD with residual, per-channel envelope harmonics."""
    rng = np.random.default_rng(seed)
    maxnhar = maxnhar or nhar
    maxnhar_e = maxnhar_e or nhar_e
    f0 = np.zeros((B, F), np.float32)
    for b in range(B):
        walk = np.cumsum(rng.normal(0, 1.5, F)) + rng.uniform(f0_lo + 10, f0_hi - 10)
        walk = f0_lo + np.abs((walk - f0_lo) % (2 * (f0_hi - f0_lo)) - (f0_hi - f0_lo))
        v = np.ones(F, bool)
        i = 0
        while i < F:
            if rng.random() < unvoiced / 10.0 * 1.0:
                run = int(rng.integers(10, 25)); v[i:i + run] = False; i += run
            else:
                i += 1
        f0[b] = np.where(v, walk, 0).astype(np.float32)
    voiced = f0 > 0
    nh = np.where(voiced, np.minimum(nhar, np.floor(fs / 2 / np.maximum(f0, 1.0)).astype(int)), 0).astype(np.int32)
    k = np.arange(1, maxnhar + 1, dtype=np.float64)
    ampl = (k[None, None, :] ** -1.2) * np.exp(rng.normal(0, 0.3, (B, F, maxnhar)))
    ampl *= (np.arange(maxnhar)[None, None, :] < nh[:, :, None])
    scale = rms / np.sqrt(np.maximum((ampl ** 2).sum(-1) / 2, 1e-12))
    ampl = (ampl * scale[:, :, None]).astype(np.float32)
    ph0 = rng.uniform(-np.pi, np.pi, (B, 1, maxnhar))
    adv = 2 * np.pi * thop * np.cumsum(f0.astype(np.float64), axis=1)
    phse = ph0 + adv[:, :, None] * k[None, None, :] + rng.normal(0, 0.05, (B, F, maxnhar))
    phse = ((phse + np.pi) % (2 * np.pi) - np.pi).astype(np.float32)
    phse *= (np.arange(maxnhar)[None, None, :] < nh[:, :, None])
    j = np.arange(npsd)
    psd = (-60.0 - 20.0 * j / npsd)[None, None, :] + rng.normal(0, 2.0, (B, F, npsd))
    psd = np.where(voiced[:, :, None], psd, -40.0 + rng.normal(0, 2.0, (B, F, npsd))).astype(np.float32)
    psdres = rng.normal(0, 1.0, (B, F, npsd)).astype(np.float32)
    edc = rng.uniform(1e-4, 1e-2, (B, F, nch)).astype(np.float32)
    enhar = np.where(voiced[:, :, None], nhar_e, 0).astype(np.int32) * np.ones((1, 1, nch), np.int32)
    eampl = (0.3 * edc[..., None] * np.ones((1, 1, 1, maxnhar_e)))
    eampl = (eampl * (np.arange(maxnhar_e)[None, None, None, :] < enhar[..., None])).astype(np.float32)
    ephse = rng.uniform(-np.pi, np.pi, (B, F, nch, maxnhar_e)).astype(np.float32)
    ephse *= (np.arange(maxnhar_e)[None, None, None, :] < enhar[..., None])
    fr = dict(f0=f0, nhar=nh, ampl=ampl, phse=phse, psd=psd, psdres=psdres, edc=edc,
              enhar=np.ascontiguousarray(enhar), eampl=np.ascontiguousarray(eampl),
              ephse=np.ascontiguousarray(ephse), nfrm_utt=None)
    conf = abi.make_conf(B, F, maxnhar, maxnhar_e, npsd, nch, fs, thop)
    return fr, conf


