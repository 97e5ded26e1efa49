"""Driver for oracle/_ref/ref_gpu_harness: the reference's OWN CUDA kernels (unmodified, rebuilt for sm_100a).
TEST / BASELINE INFRASTRUCTURE ONLY (GPU oracle + the GPU baseline to beat); never imported by the product."""
import json
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(HERE, "_ref", "ref_gpu_harness")


def available():
    return os.path.exists(BIN)


def run(w, precision, R, S, A, L, max_dilation, B, N, mode=3, tanh_embed=True, chunk=2048, reps=1, timeout=600):
    """w: dict as tests/refgen.py produces.  Returns dict(ms, khz, y [B][N], za [B][A], p [B][A])."""
    with tempfile.TemporaryDirectory() as d:
        fin, fout = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        with open(fin, "wb") as f:
            np.array([precision, R, S, A, L, max_dilation, B, N, mode, int(tanh_embed), chunk, reps], np.int32).tofile(f)
            w["embPrev"].astype(np.float32).tofile(f); w["embCur"].astype(np.float32).tofile(f)
            for l in range(L):
                for k in ("Wprev", "Wcur", "Bh", "Wres", "Bres", "Wskip", "Bskip"):
                    np.ascontiguousarray(w[k][l], np.float32).tofile(f)
            for k in ("Wzs", "Bzs", "Wza", "Bza"):
                np.ascontiguousarray(w[k], np.float32).tofile(f)
            np.ascontiguousarray(w["Lh"], np.float32).tofile(f)
            np.ascontiguousarray(w["selectors"], np.float32).tofile(f)
        res = subprocess.run([BIN, fin, fout], capture_output=True, text=True, timeout=timeout)
        if res.returncode != 0:
            raise RuntimeError(f"ref_gpu_harness failed ({res.returncode}): {res.stderr[-500:]} {res.stdout[-300:]}")
        info = json.loads([ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1])
        with open(fout, "rb") as f:
            ms = np.fromfile(f, np.float32, 1)[0]
            y = np.fromfile(f, np.int32, B * N).reshape(B, N)
            za = np.fromfile(f, np.float32, B * A).reshape(B, A)
            p = np.fromfile(f, np.float32, B * A).reshape(B, A)
    return {"ms": float(ms), "khz": info["khz_per_utterance"], "samples_per_s": info["samples_per_s"], "y": y, "za": za, "p": p}
