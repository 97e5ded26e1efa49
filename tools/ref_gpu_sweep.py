"""Batch sweep: reference kernels (oracle/_ref/ref_gpu_harness) vs ours, fp16 C3 model, kHz and samples/s."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import nv_wavenet_b200 as nw
from oracle import ref_gpu
from tests import refgen

R, S, A, L, MD = 64, 256, 256, 20, 512
N = 600
out = {}
for B in (128, 256, 512, 1024):
    w = refgen.synthetic_inputs(1, R, S, A, L, B, N)
    row = {}
    for mode, name in ((2, "dual_block"), (3, "persistent")):
        try:
            r = ref_gpu.run(w, 16, R, S, A, L, MD, B, N, mode=mode, chunk=2048, reps=1, timeout=200)
            row["reference_" + name] = {"khz_per_utterance": round(r["khz"], 2), "samples_per_s": round(r["samples_per_s"])}
        except Exception as ex:          # noqa: BLE001
            row["reference_" + name] = {"error": str(ex)[:160]}
    e = nw.NVWavenetInfer(L, MD, B, N, R=R, S=S, A=A, dtype=nw.FP16)
    e.load(w); e.set_inputs(w["Lh"], w["selectors"])
    y = torch.zeros((B, N), dtype=torch.int32).pin_memory()
    for it in range(2):
        e.reset_history(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); e.run_chunks(2048, lambda *a: None, N, B, y); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    row["ours"] = {"khz_per_utterance": round(N / ms, 2), "samples_per_s": round(N * B / (ms * 1e-3))}
    out[f"fp16_B{B}"] = row
    print(f"B={B}", json.dumps(row), flush=True)
    del e
json.dump(out, open("gpurun_out/ref_gpu_sweep.json", "w"), indent=1)
