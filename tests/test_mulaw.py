"""Output side: mu-law decode + int16 conversion.  CPU part: the oracle restatement and the
library's host-computed decode table against the golden vectors generated from the reference's own pytorch/utils.py
(tests/golden/make_golden_mulaw.py)."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import mulaw

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "mulaw_decode.npz"))


@pytest.mark.parametrize("A", [256, 512, 1024])
def test_oracle_matches_reference_vectors(A):
    x = np.arange(A)
    audio = mulaw.mu_law_decode(x, A)
    assert audio.dtype == np.float64
    assert np.array_equal(audio, GOLD[f"audio_{A}"])                       # bit for bit
    assert np.array_equal(mulaw.to_int16(audio), GOLD[f"int16_{A}"])
    sat = mulaw.to_int16(audio, saturate=True)
    assert sat[-1] == 32767 and np.array_equal(sat[:-1], GOLD[f"int16_{A}"][:-1])


@pytest.mark.parametrize("A", [256, 512, 1024])
def test_library_table_matches_reference_vectors(A):
    """nvwn_mulaw_table is host code (no GPU needed): the table nvwn_get_audio gathers from on the device."""
    from nv_wavenet_b200 import _lib
    lib = _lib.lib()
    f = np.empty(A, np.float32); w = np.empty(A, np.int16); s = np.empty(A, np.int16)
    assert lib.nvwn_mulaw_table(A, C.c_void_p(f.ctypes.data), C.c_void_p(w.ctypes.data), C.c_void_p(s.ctypes.data)) == 0
    assert np.array_equal(f, GOLD[f"audio_{A}"].astype(np.float32))
    assert np.array_equal(w, GOLD[f"int16_{A}"])
    assert s[-1] == 32767 and np.array_equal(s[:-1], GOLD[f"int16_{A}"][:-1])
    assert lib.nvwn_mulaw_table(1, None, None, None) != 0


def test_properties():
    audio = mulaw.mu_law_decode(np.arange(256), 256)
    assert np.all(np.diff(audio) > 0)                                      # monotone
    assert np.allclose(audio, -audio[::-1], rtol=0, atol=1e-15)            # odd symmetry around the mid code pair
    assert audio[0] == -1.0 and audio[-1] == 1.0
