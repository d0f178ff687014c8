#!/bin/bash
# same-box A/B: 64 vs 32 probe cells of the warp decoder's prune bound
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
summ='import json,sys
d=json.loads(sys.stdin.read()); print("value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "clk", d["clocks"]["sm_mhz"], d["stages_ms_sum_of_launches"]["ctc_decode"])'
for rep in 1 2 3; do
for c in 64 32; do
echo -n "cells $c: "; GASR_CTC_CELLS=$c GASR_WAVE_TIMEOUT_S=20 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-checks 2>/dev/null | tail -1 | python -c "$summ"
done; done
echo "decoder alone (serial, 4096 utterances): 64 / 32"
summ2='import json,sys
d=json.loads(sys.stdin.read()); print(d["stages_ms_sum_of_launches"])'
for c in 64 32; do
GASR_WAVE_SERIAL=1 GASR_CTC_CELLS=$c GASR_WAVE_TIMEOUT_S=30 timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-checks --utts 4096 --wave 4096 --lanes 1 2>/dev/null | tail -1 | python -c "$summ2"
done
echo "survivor statistics (warp kernel, N = 512, T = 300, beam 16)"
for c in 64 32; do
GASR_CTC_KERNEL=w GASR_CTC_CELLS=$c timeout 200 python - <<'PY'
import os, sys, time
sys.path.insert(0, "gpu-accelerated-speech-recognition_b200"); sys.path.insert(0, ".")
import numpy as np, gasr, synth
ctx = gasr.Context(0)
for kind in ("random", "peaky", "flat"):
    T, N, V = 300, 512, 29
    if kind == "random": lp = synth.random_logprobs(5, T, N, V)
    elif kind == "flat": lp = np.log(np.full((T, N, V), 1.0 / V, dtype=np.float32) * (1 + 1e-3 * np.random.default_rng(1).normal(size=(T, N, V)).astype(np.float32)))
    else:
        rng = np.random.default_rng(2); z = rng.normal(size=(T, N, V)).astype(np.float32) * 6; z -= z.max(-1, keepdims=True); lp = (z - np.log(np.exp(z).sum(-1, keepdims=True))).astype(np.float32)
    d = ctx.to_device(lp)
    for i in range(2):
        t0 = time.time(); r = ctx.ctc_decode(d, gasr.DOMAIN_LOG, T, N, V, V, 16, 0, synth.VOCAB29); dt = time.time() - t0
    fb, sv = ctx.ctc_last_stats()
    print("cells", os.environ["GASR_CTC_CELLS"], kind, "ms %.2f" % (dt * 1e3), "fallback frames", fb, "mean survivors %.1f" % (sv / (T * N)))
    ctx.free(d)
PY
done
} > gpurun_out/probe44.log 2>&1
echo done
