"""Debug tool: prints a clock64 timeline of one sample of the tensor-core kernel (block 0)."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import nv_wavenet_b200 as nw
from nv_wavenet_b200 import _lib
from tests import refgen

L, R, S, A, md = 20, 64, 256, 256, 512
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N, T = 40, 30
w = refgen.lively_inputs(3, R, S, A, L, B, N)
e = nw.NVWavenetInfer(L, md, B, N, R=R, S=S, A=A, dtype=nw.FP16)
e.load(w); e.set_inputs(w["Lh"], w["selectors"])
lib = _lib.lib()
lib.nvwn_debug_trace.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
TID = int(sys.argv[4]) if len(sys.argv) > 4 else 0
assert lib.nvwn_debug_trace(e._h, T | (TID << 16), None, 0) == 0
e.run(N, B, None); torch.cuda.synchronize()
buf = np.zeros(3 * 1024, np.uint64)
assert lib.nvwn_debug_trace(e._h, T, buf.ctypes.data, 1) == 0
names = {28: "mma: hist waited", 26: "mma: res mmas issued", 29: "mma: Wf stage ready", 32: "mma: open_f done", 33: "mma: curx stage ready", 34: "mma: curx issued", 13: "E2 done", 14: "cond seen", 15: "prestore st done", 16: "gate ld done", 5: "cond loaded", 25: "mma: res issued(F)", 30: " E1 ld0 done", 31: " E1 half0 done", 32: " E1 ld1 done", 33: " E1 half1 done", 40: " E2 ld0 done", 41: " E2 half0 done",
         42: " E2 ld1 done", 43: " E2 half1 done", 1: "x0 arrive", 2: "D1 full seen", 12: "E1 math done", 3: "h arrive", 4: "Dx full seen", 5: "x arrive", 6: "skip full seen",
         7: "skq arrive", 9: "zsq arrive", 10: "Dza seen", 11: "sample done", 20: "mma: x seen", 21: "mma: cur issued",
         22: "mma: h seen", 23: "mma: res issued", 24: "mma: layer issued"}
ev = []
for role in range(3):
    for v in buf[role * 1024:(role + 1) * 1024]:
        v = int(v)
        if v:
            ev.append((v & 0xFFFFFFFFFFFF, role, v >> 48))
ev.sort()
t0 = ev[0][0]
prev = t0
lo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
for clk, role, tag in ev[lo:lo + (int(sys.argv[2]) if len(sys.argv) > 2 else 70)]:
    print(f"{clk - t0:8d} (+{clk - prev:6d})  role{role}  {names.get(tag, 'prod layer %d' % (tag - 100) if tag >= 100 else tag)}")
    prev = clk
print("total cycles in sample:", ev[-1][0] - t0, "events", len(ev))
