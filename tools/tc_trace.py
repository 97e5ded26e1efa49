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
names = {1: "x0 published", 14: "cond tile landed", 5: "cond in registers", 4: "Dx full seen", 13: "x tile published (E2 done)",
         2: "D1 full seen", 16: "gate: TMEM read done", 12: "gate: math done", 3: "h published", 6: "skip full seen",
         7: "relu(skip) published", 9: "relu(Zs) published", 10: "Dza seen", 11: "sample done", 15: "prestore st done",
         20: "A: x0 seen", 22: "A: h seen", 25: "A: res issued", 21: "A: Wf issued", 23: "issuer: skip issued", 24: "A: layer issued"}
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
