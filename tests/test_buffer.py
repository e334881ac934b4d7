"""include/buffer.h (the ring buffers of the streaming synthesizer, used by libllsm2_b200's llsmrt layer) against the
contract of the reference's header of the same name: the reference's own equivalence test (test/test-structs.c:168-214)
and a seeded script of every operation whose read-back log must equal the log of the reference's implementation
(tests/golden/buffer_script.npz, generated here from /root/reference/buffer.h by this file's __main__; compared live
as well whenever /root/reference is present)."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c", "buffer_harness.c")
GOLD = os.path.join(ROOT, "tests", "golden", "buffer_script.npz")
CASES = [(1, 4000, 64), (2, 6000, 257), (3, 3000, 4096)]


def _build(incdir, tag):
    so = os.path.join(tempfile.gettempdir(), "llsm_buffer_%s_%d.so" % (tag, os.getpid()))
    subprocess.check_call(["gcc", "-std=c99", "-O1", "-g", "-shared", "-fPIC", "-I" + incdir, SRC, "-o", so])
    L = C.CDLL(so)
    L.buffer_script.argtypes = [C.c_uint, C.c_int, C.c_int, C.c_void_p, C.c_int]
    return L


def _log(L, seed, nops, cap):
    out = np.zeros(4_000_000, np.float32)
    n = L.buffer_script(seed, nops, cap, out.ctypes.data, len(out))
    assert 0 < n < len(out)
    return out[:n].copy()


@pytest.fixture(scope="module")
def ours():
    return _build(os.path.join(ROOT, "include"), "ours")


def test_reference_equivalences(ours):
    assert ours.buffer_equivalences() == 0


def test_script_matches_golden(ours):
    g = np.load(GOLD)
    for seed, nops, cap in CASES:
        assert np.array_equal(_log(ours, seed, nops, cap), g["log_%d" % seed]), seed


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree not present")
def test_script_matches_reference_header(ours):
    ref = _build("/root/reference", "ref")
    assert ref.buffer_equivalences() == 0
    for seed, nops, cap in CASES + [(11, 5000, 100)]:
        assert np.array_equal(_log(ours, seed, nops, cap), _log(ref, seed, nops, cap)), seed


if __name__ == "__main__":          # regenerate the golden logs from the reference's header
    ref = _build("/root/reference", "ref")
    np.savez_compressed(GOLD, **{"log_%d" % s: _log(ref, s, n, c) for s, n, c in CASES})
    print("wrote", GOLD)
