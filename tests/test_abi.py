"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol the
headers declare, and refuses to run without a GPU (no CPU fallback)."""
import os
import re

import pytest

import nv_wavenet_b200 as nw
from nv_wavenet_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b((?:nvwn_|wavenet_infer|get_[RSA])\w*)\s*\(", src))


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()
    declared = _declared("nvwn_b200.h") | _declared("wavenet_infer.h")
    assert {"wavenet_infer", "get_R", "get_S", "get_A", "nvwn_create", "nvwn_run_partial"} <= declared
    for name in declared:
        assert hasattr(lib, name), f"libwavenet_infer.so does not export {name}"
    # and the ctypes table covers every declaration
    assert declared == set(_lib.SYMBOLS)


def test_channel_counts_match_reference_build():
    lib = _lib.lib()                       # pytorch/wavenet_infer.cu:34-37
    assert (lib.get_R(), lib.get_S(), lib.get_A()) == (64, 256, 256)


def test_no_cpu_fallback():
    lib = _lib.lib()
    if lib.nvwn_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(nw.NvwnError, match="no CUDA device"):
        nw.NVWavenetInfer(2, 2, 1, 4)


def test_argument_validation_messages():
    lib = _lib.lib()
    import ctypes as C
    h = C.c_void_p()
    assert lib.nvwn_create(C.byref(h), 7, 64, 256, 256, 2, 2, 1, 4, 0, 1) == -1
    assert b"dtype" in lib.nvwn_last_error()
    assert lib.nvwn_create(C.byref(h), 0, 48, 256, 256, 2, 2, 1, 4, 0, 1) == -2
    assert b"unsupported channel" in lib.nvwn_last_error()
    assert lib.nvwn_set_inputs(None, None, None) == -1


def test_cpp_facade_compiles():
    """include/nv_wavenet.hpp (the nvWavenetInfer<> template facade) is valid C++ against the C-ABI."""
    import subprocess
    import tempfile
    src = r'''
#include "nv_wavenet.hpp"
int main() {
    typedef nvWavenetInfer<float, float, 64, 256, 256> W32;
    typedef nvWavenetInfer<half2, half, 64, 256, 256> W16;
    static_assert(W32::AUTO == 0 && W32::PERSISTENT == 3 && W16::MANYBLOCK_NONPERSISTENT == 4, "enum");
    if (nvwn_device_count() == 0) return 0;
    W32 w(2, 2, 1, 4);
    int y[4];
    return w.run(4, 1, y) ? 0 : 1;
}
'''
    with tempfile.TemporaryDirectory() as d:
        f = os.path.join(d, "t.cpp")
        open(f, "w").write(src)
        exe = os.path.join(d, "t")
        cmd = ["g++", "-std=c++14", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include", f, "-o", exe,
               "-L", os.path.dirname(nw.LIB_PATH), "-lwavenet_infer", "-Wl,-rpath," + os.path.dirname(nw.LIB_PATH)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        assert res.returncode == 0, res.stderr
        assert subprocess.run([exe]).returncode == 0


def test_libc_selectors_replay_the_reference_draw():
    """nvwn_libc_selectors (host helper, no GPU): Matrix(batch, samples).randomize(0.5, 1.0) on the caller's rand() stream
    (pytorch/wavenet_infer.cu:92-93, matrix.cpp:38-56), checked against the glibc replay of tests/refgen.py."""
    import ctypes as C

    import numpy as np

    from nv_wavenet_b200 import _lib
    from tests import refgen
    lib = _lib.lib()
    B, N = 5, 7
    C.CDLL(None).srand(321)
    sel = np.empty(N * B, np.float32)
    assert lib.nvwn_libc_selectors(C.c_void_p(sel.ctypes.data), B, N) == 0
    want = refgen.randomize(refgen.GlibcRand(321), B, N, np.float32(0.5), np.float32(1.0)).reshape(N, B)
    assert np.array_equal(sel.reshape(N, B), want)
    assert lib.nvwn_libc_selectors(None, B, N) != 0
