"""Golden vectors for the decode-bound configuration (BASELINE.json cfg4: T = 4000, beam 8 / 32 / 128): the CPU oracle's
transcript and fp32 score for ONE synthetic utterance per beam width.  The oracle needs minutes at beam 128 (explicit path
strings), so the GPU tests compare against this file instead of running it on the GPU box.

    python tests/golden/make_golden_ctc.py        # writes tests/golden/ctc_cfg4.json

Inputs are regenerated from the seed by synth.random_logprobs / synth.peaky_logprobs (nothing but the seed is stored)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "gpu-accelerated-speech-recognition_b200"))
sys.path.insert(0, ROOT)
import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

T, V = 4000, 29
out = {"T": T, "V": V, "cases": []}
for kind, seed in (("random", 77), ("peaky", 78)):
    gen = synth.random_logprobs if kind == "random" else synth.peaky_logprobs
    lp = gen(seed, T, 1, V)
    for beam in (8, 32, 128):
        p, s = O.ctc_decode(lp, synth.VOCAB29, 0, beam, domain="log")
        out["cases"].append({"kind": kind, "seed": seed, "beam": beam, "path_hex": p[0].hex(),
                             "score_bits": int(np.float32(s[0]).view(np.uint32))})
        print(kind, beam, len(p[0]), s[0], flush=True)
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "ctc_cfg4.json"), "w"), indent=1)
