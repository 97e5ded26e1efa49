"""Output side on the device: nvwn_get_audio decodes the engine's yOut with the reference's
mu-law expansion (pytorch/utils.py:62-70, pytorch/nv_wavenet_inference.py:55-60).  Bit-exact against the oracle:
the device only gathers from a host-computed table."""
import numpy as np
import pytest

from oracle import mulaw
from tests import refgen

pytestmark = pytest.mark.gpu


def _engine(dtype_name="fp32", B=5, N=48):
    import nv_wavenet_b200 as nw
    L, R, S, A, md = 4, 64, 256, 256, 4
    w = refgen.lively_inputs(11, R, S, A, L, B, N)
    e = nw.NVWavenetInfer(L, md, B, N, R=R, S=S, A=A, dtype=nw.FP16 if dtype_name == "fp16" else nw.FP32)
    e.load(w); e.set_inputs(w["Lh"], w["selectors"])
    y = np.zeros((B, N), np.int32)
    e.run(N, B, y); e.synchronize()
    return e, y


@pytest.mark.parametrize("dtype_name", ["fp32", "fp16"])
def test_audio_matches_oracle(dtype_name):
    e, y = _engine(dtype_name)
    assert y.min() >= 0 and y.max() < 256 and len(np.unique(y)) > 2
    want = mulaw.mu_law_decode(y, 256)
    got = e.get_audio()
    assert got.dtype == np.float32 and got.shape == y.shape
    assert np.array_equal(got, want.astype(np.float32))
    assert np.array_equal(e.get_audio(int16=True), mulaw.to_int16(want))
    assert np.array_equal(e.get_audio(int16=True, saturate=True), mulaw.to_int16(want, saturate=True))
    # a chunk in the middle, as a streaming consumer would ask for it
    assert np.array_equal(e.get_audio(offset=7, size=20), want[:, 7:27].astype(np.float32))
    with pytest.raises(Exception):
        e.get_audio(offset=40, size=20)                                    # out of range


def test_audio_into_device_tensor_on_a_stream():
    import torch
    e, y = _engine()
    st = torch.cuda.Stream()
    out = torch.empty((y.shape[0], 16), dtype=torch.int16, device="cuda")
    e.get_audio(offset=32, size=16, int16=True, out=out, stream=st)
    st.synchronize()
    assert np.array_equal(out.cpu().numpy(), mulaw.to_int16(mulaw.mu_law_decode(y[:, 32:48], 256)))
