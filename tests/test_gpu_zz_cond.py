"""Conditioning producer on the device: nvwn_set_conditioning_from_features must leave the
engine in exactly the state nvwn_set_inputs(Lh) leaves it in when Lh is the host restatement of the same arithmetic
(itself checked against the reference's module in tests/test_cond_producer.py): identical sampled indices."""
import ctypes as C
import os

import numpy as np
import pytest

from tests import refgen
from tests.test_cond_producer import GOLD, host_cond

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype_name", ["fp32", "fp16"])
def test_generation_from_features_equals_generation_from_host_conditioning(dtype_name):
    import nv_wavenet_b200 as nw
    rng = np.random.default_rng(5)
    L, R, S, A, md = 4, 64, 256, 256, 4
    B, Cc, T, window, stride = 3, 6, 5, 24, 8                    # 40 samples
    N = T * stride + 6                                           # the producer fills [6, 46)
    first = 6
    w = refgen.lively_inputs(31, R, S, A, L, B, N)
    g = {"x_features": rng.standard_normal((B, Cc, T)).astype(np.float32),
         "x_upsample_weight": (0.3 * rng.standard_normal((Cc, Cc, window))).astype(np.float32),
         "x_upsample_bias": (0.1 * rng.standard_normal(Cc)).astype(np.float32),
         "x_cond_weight": (np.abs(w["Lh"]).max() * rng.standard_normal((L * 2 * R, Cc))).astype(np.float32),
         "x_cond_bias": (0.1 * np.abs(w["Lh"]).max() * rng.standard_normal(L * 2 * R)).astype(np.float32),
         "x_geometry": np.array([Cc, T, window, stride, L, R, B])}
    lh = np.array(w["Lh"], np.float32)                           # [N][L][B][2R]; samples before `first` keep the synthetic values
    lh[first:] = host_cond(g, "x")

    dt = nw.FP16 if dtype_name == "fp16" else nw.FP32
    ys = []
    for mode in ("host", "device"):
        e = nw.NVWavenetInfer(L, md, B, N, R=R, S=S, A=A, dtype=dt)
        e.load(w)
        if mode == "host":
            e.set_inputs(lh, w["selectors"])
        else:
            e.set_inputs(np.ascontiguousarray(w["Lh"], np.float32), w["selectors"])
            n = e.set_conditioning_from_features(g["x_features"], g["x_upsample_weight"], g["x_upsample_bias"], g["x_cond_weight"],
                                                 g["x_cond_bias"], stride, first_sample=first)
            assert n == T * stride
        y = np.zeros((B, N), np.int32)
        e.run(N, B, y); e.synchronize()
        ys.append(y)
    assert np.array_equal(ys[0], ys[1])
    assert len(np.unique(ys[0])) > 2


def test_out_of_range_is_rejected():
    import nv_wavenet_b200 as nw
    e = nw.NVWavenetInfer(2, 2, 2, 16, R=64, S=256, A=256, dtype=nw.FP32)
    z = np.zeros((2, 3, 5), np.float32)
    with pytest.raises(Exception):
        e.set_conditioning_from_features(z, np.zeros((3, 3, 8), np.float32), np.zeros(3, np.float32), np.zeros((2 * 2 * 64, 3), np.float32),
                                         np.zeros(2 * 2 * 64, np.float32), 4)        # 5 * 4 = 20 samples > 16
