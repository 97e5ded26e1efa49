"""Generates tests/golden/mulaw_decode.npz by calling the reference's own post-processing
(pytorch/utils.py:62-70 mu_law_decode_numpy; pytorch/nv_wavenet_inference.py:58-60) in the build container.
/root/reference does not exist on the GPU box: only the committed .npz travels.

    python tests/golden/make_golden_mulaw.py
"""
import os
import sys
import warnings

import numpy as np

sys.path.insert(0, "/root/reference/pytorch")
import utils as ref_utils  # noqa: E402  (the reference's module)

out = {}
for A in (256, 512, 1024):
    x = np.arange(A)
    audio = ref_utils.mu_law_decode_numpy(x, A)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                    # numpy warns about the one out-of-range value (32768)
        wav = (ref_utils.MAX_WAV_VALUE * audio).astype("int16")
    out[f"audio_{A}"] = audio
    out[f"int16_{A}"] = wav
np.savez(os.path.join(os.path.dirname(__file__), "mulaw_decode.npz"), **out)
print({k: (v.dtype, v.shape, v[:2], v[-2:]) for k, v in out.items()})
