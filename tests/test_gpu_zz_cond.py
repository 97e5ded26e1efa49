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


@pytest.mark.parametrize("kernel", ["fp32", "stream", "tc", "lat"])
def test_device_conditioning_store_matches_reference_module(kernel, monkeypatch):
    """The tensor the device producer leaves in the engine's conditioning store (read back through a debug getter, whatever the
    kernel-native layout) against vectors from the reference's own WaveNet.get_cond_input (tests/golden/make_golden_cond.py):
    fp32 within 1e-5 of the tensor scale, fp16 stores within 1e-3."""
    import nv_wavenet_b200 as nw
    from nv_wavenet_b200 import _lib
    for k in ("NVWN_FP16_KERNEL", "NVWN_TC_TILE", "NVWN_TC_NODUP"):
        monkeypatch.delenv(k, raising=False)
    if kernel != "fp32":
        monkeypatch.setenv("NVWN_FP16_KERNEL", kernel)
    Cc, T, window, stride, L, R, B = [int(v) for v in GOLD["c_geometry"]]
    N = T * stride + 5
    first = 5
    e = nw.NVWavenetInfer(L, 2, B, N, R=R, S=256, A=256, dtype=nw.FP32 if kernel == "fp32" else nw.FP16)
    n = e.set_conditioning_from_features(GOLD["c_features"], GOLD["c_upsample_weight"], GOLD["c_upsample_bias"], GOLD["c_cond_weight"],
                                         GOLD["c_cond_bias"], stride, first_sample=first)
    assert n == T * stride
    lib = _lib.lib()
    lib.nvwn_debug_get_conditioning.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    got = np.full((n, L, B, 2 * R), np.nan, np.float32)
    assert lib.nvwn_debug_get_conditioning(e._h, C.c_void_p(got.ctypes.data), first, n) == 0
    want = GOLD["c_Lh"]
    scale = np.abs(want).max()
    tol = 1e-5 if kernel == "fp32" else 1e-3
    assert np.isfinite(got).all() and np.abs(got - want).max() <= tol * scale, np.abs(got - want).max() / scale


def test_out_of_range_is_rejected():
    import nv_wavenet_b200 as nw
    e = nw.NVWavenetInfer(2, 2, 2, 16, R=64, S=256, A=256, dtype=nw.FP32)
    z = np.zeros((2, 3, 5), np.float32)
    with pytest.raises(Exception):
        e.set_conditioning_from_features(z, np.zeros((3, 3, 8), np.float32), np.zeros(3, np.float32), np.zeros((2 * 2 * 64, 3), np.float32),
                                         np.zeros(2 * 2 * 64, np.float32), 4)        # 5 * 4 = 20 samples > 16
