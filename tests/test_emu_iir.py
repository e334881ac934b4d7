"""The shared-memory resident sub-band filter (kernels_iir_smem.cuh, single-CTA instance: clusters exist on the GPU
only, where the sharded-state carry is covered by the analysis parity tests at 88 k / 148 k / 331 k samples) against the
streaming kernel it replaces (kernels_iir.cuh) on the CPU thread emulation: same chunk-parallel mathematics, so the
float outputs must agree to the last bits -- including rows that are not 16-byte aligned (plain loads instead of the
bulk copy), lengths that are not multiples of four, and the squared output. (This test found the first version of the
shared-memory kernel letting the forward pass ring on into the zero padding behind the sequence, which the backward pass
then read: filtfilt filters exactly n samples each way.)"""
import ctypes as C
import numpy as np
import pytest
import support as S


@pytest.mark.parametrize("n,sstride,ystride,fs,square", [
    (4099, 4099, 4100, 44100.0, 1),      # odd row stride: the second utterance's row is misaligned
    (2500, 2504, 2500, 16000.0, 0),      # 16 kHz: other filter selections, plain output
    (9001, 9004, 9004, 44100.0, 1),
])
def test_smem_filter_matches_streaming_filter(n, sstride, ystride, fs, square):
    emu = S.load_emu()
    rng = np.random.default_rng(n)
    nutt, nch = 2, 4
    src = np.zeros((nutt, sstride), np.float32)
    src[:, :n] = rng.normal(0, 0.1, (nutt, n)).astype(np.float32)
    cf = np.asarray([2000, 4000, 8000, 0, 0, 0, 0, 0], np.float32)
    ya = np.full((nutt * nch, ystride), np.nan, np.float32)
    yb = np.full((nutt * nch, ystride), np.nan, np.float32)
    emu.emu_iir_both.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                 C.c_void_p, C.c_void_p, C.c_int]
    rc = emu.emu_iir_both(nutt, nch, n, fs, cf.ctypes.data, src.ctypes.data, sstride, square, ya.ctypes.data,
                          yb.ctypes.data, ystride)
    assert rc == 0
    a, b = ya[:, :n], yb[:, :n]
    assert np.isfinite(a).all() and np.isfinite(b).all()
    scale = float(np.abs(a).max())
    assert scale > 0
    assert np.abs(a - b).max() <= 2e-7 * scale, np.abs(a - b).max() / scale
    assert np.mean(a == b) > 0.99                                       # the same chunked arithmetic: nearly all bits equal
