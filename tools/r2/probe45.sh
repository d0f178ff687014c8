#!/bin/bash
# short chunks at both ends of a batch (fill / drain): parity, then same-box A/B at the shares of N = 8 / 4 / 1
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_wide.py tests/test_gpu_parity.py tests/test_gpu_lengths.py tests/test_gpu_sizes.py -x -q -m gpu -k "cfg5 or cfg2 or lengths or pipeline or job" 2>&1 | tail -3
summ='import json,sys
d=json.loads(sys.stdin.read()); print("utts", d["config"]["utterances"], "value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "clk", d["clocks"]["sm_mhz"])'
for a in "1024 1024" "2048 2048" "8192 4096"; do set -- $a
for rep in 1 2; do for r in 0 1; do
echo -n "ramp $r: "; GASR_CHUNK_RAMP=$r GASR_WAVE_TIMEOUT_S=20 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-checks --utts $1 --wave $2 2>/dev/null | tail -1 | python -c "$summ"
done; done; done
} > gpurun_out/probe45.log 2>&1
echo done
