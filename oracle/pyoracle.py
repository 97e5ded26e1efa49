"""ctypes doors to the CPU checkers.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg may import this module; nothing in nv_wavenet_b200/ does.

  Oracle   -> oracle/liboracle.so          plain-C restatement (wavenet_oracle.c)
  RefCPU   -> oracle/_ref/libnvwn_ref.so   the reference's own nv_wavenet_reference.cpp,
                                           compiled unmodified (oracle/Makefile)
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int)

MATH_LIBM, MATH_PORTABLE = 0, 1
PREC_FP32, PREC_FP16 = 0, 1


def build(quiet=True):
    """(Re)build liboracle.so and, where /root/reference exists, _ref/libnvwn_ref.so."""
    subprocess.run(["make", "-C", _HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _fp(a):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_f32p)


def _ip(a):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_i32p)


def have_ref():
    return os.path.exists(os.path.join(_HERE, "_ref", "libnvwn_ref.so"))


class _Base:
    """Shared setter/getter surface (mirrors nvWavenetReference, nv_wavenet_reference.h:83-100)."""

    def load(self, w):
        """w: dict from tests.refgen / ref_gen_test_inputs."""
        self.set_embeddings(w["embPrev"], w["embCur"])
        for l in range(self.L):
            self.set_layer_weights(l, w["Wprev"][l], w["Wcur"][l], w["Bh"][l], w["Wres"][l],
                                   w["Bres"][l], w["Wskip"][l], w["Bskip"][l])
        self.set_out_weights(w["Wzs"], w["Bzs"], w["Wza"], w["Bza"])

    def activations(self):
        return {
            "xt": np.stack([self.get_xt_out(l) for l in range(self.L)]),
            "skip": np.stack([self.get_skip_out(l) for l in range(self.L)]),
            "zs": self.get_zs(), "za": self.get_za(), "p": self.get_p(),
        }


class Oracle(_Base):
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            path = os.path.join(_HERE, "liboracle.so")
            if not os.path.exists(path):
                build()
            lib = C.CDLL(path)
            lib.wno_create.restype = C.c_void_p
            lib.wno_create.argtypes = [C.c_int] * 7
            lib.wno_destroy.argtypes = [C.c_void_p]
            lib.wno_set_math.argtypes = [C.c_void_p, C.c_int]
            lib.wno_set_precision.argtypes = [C.c_void_p, C.c_int]
            lib.wno_set_tanh_embed.argtypes = [C.c_void_p, C.c_int]
            lib.wno_set_forced.argtypes = [C.c_void_p, _i32p]
            lib.wno_set_logit_trace.argtypes = [C.c_void_p, _f32p]
            lib.wno_set_embeddings.argtypes = [C.c_void_p, _f32p, _f32p]
            lib.wno_set_layer_weights.argtypes = [C.c_void_p, C.c_int] + [_f32p] * 7
            lib.wno_set_out_weights.argtypes = [C.c_void_p] + [_f32p] * 4
            lib.wno_set_inputs.argtypes = [C.c_void_p, _f32p, _f32p]
            for g in ("wno_get_xt_out", "wno_get_skip_out"):
                getattr(lib, g).argtypes = [C.c_void_p, C.c_int, _f32p]
            for g in ("wno_get_zs", "wno_get_za", "wno_get_p"):
                getattr(lib, g).argtypes = [C.c_void_p, _f32p]
            lib.wno_run.restype = C.c_int
            lib.wno_run.argtypes = [C.c_void_p, C.c_int, C.c_int, _i32p]
            for f in ("wno_expf_portable", "wno_tanhf_portable", "wno_sigmoidf_portable", "wno_round_fp16"):
                getattr(lib, f).restype = C.c_float
                getattr(lib, f).argtypes = [C.c_float]
            cls._lib = lib
        return cls._lib

    def __init__(self, L, B, N, R, S, A, max_dilation, math=MATH_LIBM, prec=PREC_FP32, tanh_embed=True):
        self.L, self.B, self.N, self.R, self.S, self.A, self.max_dilation = L, B, N, R, S, A, max_dilation
        self._l = self.lib()
        self._h = C.c_void_p(self._l.wno_create(L, B, N, R, S, A, max_dilation))
        self._l.wno_set_math(self._h, math)
        self._l.wno_set_precision(self._h, prec)
        self._l.wno_set_tanh_embed(self._h, int(tanh_embed))
        self._keep = []

    def __del__(self):
        if getattr(self, "_h", None):
            self._l.wno_destroy(self._h)
            self._h = None

    def set_embeddings(self, prev, cur):
        self._l.wno_set_embeddings(self._h, _fp(prev), _fp(cur))

    def set_layer_weights(self, l, Wprev, Wcur, Bh, Wres, Bres, Wskip, Bskip):
        self._l.wno_set_layer_weights(self._h, l, *[_fp(np.ascontiguousarray(x)) for x in (Wprev, Wcur, Bh, Wres, Bres, Wskip, Bskip)])

    def set_out_weights(self, Wzs, Bzs, Wza, Bza):
        self._l.wno_set_out_weights(self._h, _fp(Wzs), _fp(Bzs), _fp(Wza), _fp(Bza))

    def set_inputs(self, Lh, selectors):
        self._l.wno_set_inputs(self._h, _fp(Lh), _fp(selectors))

    def set_forced(self, forced):
        if forced is None:
            self._l.wno_set_forced(self._h, None)
        else:
            forced = np.ascontiguousarray(forced, dtype=np.int32)
            self._keep.append(forced)
            self._l.wno_set_forced(self._h, _ip(forced))

    def set_logit_trace(self, trace):
        if trace is None:
            self._l.wno_set_logit_trace(self._h, None)
        else:
            self._keep.append(trace)
            self._l.wno_set_logit_trace(self._h, _fp(trace))

    def _get2(self, fn, layer, dim):
        out = np.empty((self.B, dim), np.float32)
        getattr(self._l, fn)(self._h, layer, _fp(out))
        return out

    def _get1(self, fn):
        out = np.empty((self.B, self.A), np.float32)
        getattr(self._l, fn)(self._h, _fp(out))
        return out

    def get_xt_out(self, l): return self._get2("wno_get_xt_out", l, self.R)
    def get_skip_out(self, l): return self._get2("wno_get_skip_out", l, self.S)
    def get_zs(self): return self._get1("wno_get_zs")
    def get_za(self): return self._get1("wno_get_za")
    def get_p(self): return self._get1("wno_get_p")

    def run(self, num_samples, batch_size):
        y = np.empty((batch_size, num_samples), np.int32)
        status = self._l.wno_run(self._h, num_samples, batch_size, _ip(y))
        self.last_status = status
        return y


class RefCPU(_Base):
    """The reference's own CPU model (oracle/_ref/libnvwn_ref.so)."""
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            path = os.path.join(_HERE, "_ref", "libnvwn_ref.so")
            if not os.path.exists(path):
                build()
            lib = C.CDLL(path)
            lib.ref_create.restype = C.c_void_p
            lib.ref_create.argtypes = [C.c_int] * 7
            lib.ref_destroy.argtypes = [C.c_void_p]
            lib.ref_set_embeddings.argtypes = [C.c_void_p, _f32p, _f32p]
            lib.ref_set_layer_weights.argtypes = [C.c_void_p, C.c_int] + [_f32p] * 7
            lib.ref_set_out_weights.argtypes = [C.c_void_p] + [_f32p] * 4
            lib.ref_set_inputs.argtypes = [C.c_void_p, _f32p, _f32p]
            for g in ("ref_get_xt_out", "ref_get_skip_out"):
                getattr(lib, g).argtypes = [C.c_void_p, C.c_int, _f32p]
            for g in ("ref_get_zs", "ref_get_za", "ref_get_p"):
                getattr(lib, g).argtypes = [C.c_void_p, _f32p]
            lib.ref_run.argtypes = [C.c_void_p, C.c_int, C.c_int, _i32p]
            lib.ref_srand.argtypes = [C.c_uint]
            lib.ref_rand.restype = C.c_int
            lib.ref_randomize.argtypes = [_f32p, C.c_int, C.c_int, C.c_float, C.c_float]
            lib.ref_gen_test_inputs.argtypes = [C.c_int] * 6 + [_f32p] * 15
            cls._lib = lib
        return cls._lib

    def __init__(self, L, B, N, R, S, A, max_dilation):
        self.L, self.B, self.N, self.R, self.S, self.A, self.max_dilation = L, B, N, R, S, A, max_dilation
        self._l = self.lib()
        self._h = C.c_void_p(self._l.ref_create(L, B, N, R, S, A, max_dilation))

    def set_embeddings(self, prev, cur):
        self._l.ref_set_embeddings(self._h, _fp(prev), _fp(cur))

    def set_layer_weights(self, l, *ws):
        self._l.ref_set_layer_weights(self._h, l, *[_fp(np.ascontiguousarray(x)) for x in ws])

    def set_out_weights(self, Wzs, Bzs, Wza, Bza):
        self._l.ref_set_out_weights(self._h, _fp(Wzs), _fp(Bzs), _fp(Wza), _fp(Bza))

    def set_inputs(self, Lh, selectors):
        self._l.ref_set_inputs(self._h, _fp(Lh), _fp(selectors))

    def _get2(self, fn, layer, dim):
        out = np.empty((self.B, dim), np.float32)
        getattr(self._l, fn)(self._h, layer, _fp(out))
        return out

    def _get1(self, fn):
        out = np.empty((self.B, self.A), np.float32)
        getattr(self._l, fn)(self._h, _fp(out))
        return out

    def get_xt_out(self, l): return self._get2("ref_get_xt_out", l, self.R)
    def get_skip_out(self, l): return self._get2("ref_get_skip_out", l, self.S)
    def get_zs(self): return self._get1("ref_get_zs")
    def get_za(self): return self._get1("ref_get_za")
    def get_p(self): return self._get1("ref_get_p")

    def run(self, num_samples, batch_size):
        y = np.empty((batch_size, num_samples), np.int32)
        self._l.ref_run(self._h, num_samples, batch_size, _ip(y))
        return y


def ref_gen_test_inputs(seed, R, S, A, L, B, N, reseed=True):
    """Inputs of the reference test (nv_wavenet_test.cu:44-220) through the reference's own
    Matrix::randomize and libc rand().  reseed=False continues the current rand() stream
    (the reference test seeds once per channel-config group, nv_wavenet_test.cu:343-387)."""
    lib = RefCPU.lib()
    if reseed:
        lib.ref_srand(seed)
    f = np.float32
    w = {
        "selectors": np.empty((N, B), f), "embPrev": np.empty((A, R), f), "embCur": np.empty((A, R), f),
        "Wprev": np.empty((L, 2 * R * R), f), "Wcur": np.empty((L, 2 * R * R), f), "Bh": np.empty((L, 2 * R), f),
        "Wres": np.empty((L, R * R), f), "Bres": np.empty((L, R), f),
        "Wskip": np.empty((L, S * R), f), "Bskip": np.empty((L, S), f),
        "Wzs": np.empty(A * S, f), "Bzs": np.empty(A, f), "Wza": np.empty(A * A, f), "Bza": np.empty(A, f),
        "Lh": np.empty((N, L, B, 2 * R), f),
    }
    order = ["selectors", "embPrev", "embCur", "Wprev", "Wcur", "Bh", "Wres", "Bres", "Wskip", "Bskip",
             "Wzs", "Bzs", "Wza", "Bza", "Lh"]
    lib.ref_gen_test_inputs(R, S, A, L, B, N, *[_fp(w[k]) for k in order])
    return w
