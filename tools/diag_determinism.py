"""Run-to-run determinism of the fp16 kernels over a 2000-sample soak (C3 model): python tools/diag_determinism.py
Round 2 finding: the 128-row tensor-core tiles (NVWN_TC_NODUP) flip a sampled index about once per 1e5 utterance-samples, most
often when the tile is only partially filled; every other variant (32- / 64-utterance tiles full or ragged, the latency kernel) is
clean.  wn_tc_tile_utt() therefore never selects 128-row tiles."""
import os
import sys

sys.path.insert(0, ".")
import numpy as np

from tests import refgen

R, S, A, L, N, md = 64, 256, 256, 20, 2000, 512
CASES = (("lat", {"NVWN_FP16_KERNEL": "lat"}, 64), ("lat_ragged", {"NVWN_FP16_KERNEL": "lat"}, 40), ("tc32", {}, 64),
         ("tc64", {"NVWN_TC_TILE": "64"}, 64), ("tc64_ragged", {"NVWN_TC_TILE": "64"}, 40), ("tc64_unfused", {"NVWN_TC_TILE": "64", "NVWN_TC_FUSED": "0"}, 64),
         ("nodup128", {"NVWN_TC_NODUP": "1"}, 128), ("nodup64", {"NVWN_TC_NODUP": "1"}, 64), ("nodup200", {"NVWN_TC_NODUP": "1"}, 200))
for name, env, B in CASES:
    for k in ("NVWN_TC_TILE", "NVWN_TC_NODUP", "NVWN_TC_FUSED"):
        os.environ.pop(k, None)
    os.environ["NVWN_FP16_KERNEL"] = "tc"
    os.environ.update(env)
    import nv_wavenet_b200 as nw
    w = refgen.lively_inputs(31, R, S, A, L, B, N)
    e = nw.NVWavenetInfer(L, md, B, N, R=R, S=S, A=A, dtype=nw.FP16)
    e.load(w); e.set_inputs(w["Lh"], w["selectors"])
    ys = []
    for it in range(8):
        e.reset_history(); y = np.zeros((B, N), np.int32); e.run(N, B, y); e.synchronize(); ys.append(y)
    bad = [int((y != ys[0]).any(axis=1).sum()) for y in ys[1:]]
    print(name, "B", B, "tile", e.launch_info()["batch_per_cta"], "utterances differing from run 0 in runs 1..7:", bad, flush=True)
    e.close()
