"""Chunk phase utilities (llsm_chunk_phasepropagate / llsm_chunk_phasesync_rps, layer0.c:687-706) of the CPU thread
emulation of the kernels against the reference build. Phases are FP_TYPE values wrapped in double: bit-exact."""
import ctypes as C
import numpy as np
import pytest
import support as S
from libllsm2_b200 import abi


def _run_emu(fr, conf, mode, arg, vsphse=None, nvs=None):
    emu = S.load_emu()
    o = dict(phse=fr["phse"].copy(), ephse=fr["ephse"].copy())
    fo = abi.FramesOut()
    fo.f0, fo.nhar, fo.enhar = fr["f0"].ctypes.data, fr["nhar"].ctypes.data, fr["enhar"].ctypes.data
    fo.phse, fo.ephse = o["phse"].ctypes.data, o["ephse"].ctypes.data
    l1 = None
    if vsphse is not None:
        o["vsphse"] = vsphse.copy()
        l1 = abi.Layer1(); l1.vsphse, l1.nvs = o["vsphse"].ctypes.data, nvs.ctypes.data
    nu = fr.get("nfrm_utt")
    assert emu.emu_phase_op(C.byref(conf), nu.ctypes.data_as(C.c_void_p) if nu is not None else None, C.byref(fo),
                            C.byref(l1) if l1 is not None else None, mode, arg) == 0
    return o


def _source_phases(fr, conf, seed):
    rng = np.random.default_rng(seed)
    nvs = np.where(fr["f0"] > 0, np.minimum(fr["nhar"], conf.maxnhar - 3), 0).astype(np.int32)
    vs = rng.uniform(-np.pi, np.pi, fr["phse"].shape).astype(np.float32)
    vs[np.arange(conf.maxnhar)[None, None, :] >= nvs[..., None]] = 0
    return vs, nvs


@pytest.mark.parametrize("mode,arg", [(0, 1), (0, -1), (1, 0), (1, 1)])
def test_phase_ops_match_reference(mode, arg):
    fr, conf = S.synth_frames(2, 50, seed=31, nhar=60, maxnhar=64)
    fr["nfrm_utt"] = np.asarray([50, 37], np.int32)
    vs, nvs = _source_phases(fr, conf, 5)
    ref = S.ref_phase_op(fr, conf, mode, arg, vs, nvs)
    got = _run_emu(fr, conf, mode, arg, vs, nvs)
    for k in ("phse", "ephse", "vsphse"):
        assert np.array_equal(got[k], ref[k]), (k, np.abs(got[k] - ref[k]).max())
    if mode == 1 and arg == 0:
        v = fr["f0"] > 0; v[1, 37:] = False
        assert np.all(got["phse"][..., 0][v] == 0)              # relative phase shift: first harmonic at zero


def test_propagate_then_back_is_identity_up_to_rounding():
    fr, conf = S.synth_frames(1, 80, seed=32)
    a = _run_emu(fr, conf, 0, 1)
    fr2 = dict(fr); fr2["phse"], fr2["ephse"] = a["phse"], a["ephse"]
    b = _run_emu(fr2, conf, 0, -1)
    # the shifted phase is rounded to float before it is wrapped (frame.c:59): at harmonic 128 it reaches 4e4 rad
    assert np.abs(S.phase_err(b["phse"], fr["phse"])).max() < 2e-2
    assert np.abs(S.phase_err(b["phse"], fr["phse"]))[..., :8].max() < 1e-3
