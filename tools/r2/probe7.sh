#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_wide.py tests/test_gpu_parity.py -x -q -m gpu -k "wide or pipeline_end_to_end" 2>&1 | tail -3
summ='import json,sys
d=json.loads(sys.stdin.read()); print("wave", d["config"]["utterances_per_batch"], "value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "stages", {k: round(v,1) for k,v in d["stages_ms_sum_of_launches"].items()})'
for env in "GASR_RNN_G=2 GASR_RNN_MC=1" "GASR_RNN_G=1 GASR_RNN_MC=1" "GASR_RNN_G=2 GASR_RNN_MC=0" "GASR_CHUNK=100"; do
echo "== $env"
env $env timeout 600 python bench.py --steps 2 --warmup 3 --wave 2048 --no-cpu-baseline --no-checks 2>&1 | tail -1 | python -c "$summ"
done
echo "== trace (wave engine, T=100 N=2048)"
GASR_LIB=$PWD/gpu-accelerated-speech-recognition_b200/build_trace/libgasr.so timeout 300 python - <<'PY' 2>&1 | grep "rw trace" | tail -3
import sys; sys.path.insert(0, "gpu-accelerated-speech-recognition_b200")
import gasr, synth
T, N, D, H, L, V, beam = 100, 2048, 161, 512, 3, 29, 16
ctx = gasr.Context(0)
pipe = gasr.AsrPipeline(ctx, gasr.CELL_TANH, False, T, N, D, H, L, V, beam, 0, synth.VOCAB29)
pipe.set_weights(*synth.rnn_weights(1, D, H, L), *synth.fc_weights(2, H, V))
x = synth.spectrogram_batch(3, T, N, D)
pipe.run_host(x); pipe.run_host(x)
PY
} > gpurun_out/probe7.log 2>&1
echo done
