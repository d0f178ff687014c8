#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_lengths.py tests/test_gpu_sizes.py tests/test_gpu_parity.py -x -q -m gpu -k "lengths or cfg5 or job or pipeline_end_to_end or variable" 2>&1 | tail -3
summ='import json,sys
d=json.loads(sys.stdin.read()); print("utts", d["config"]["utterances"], "value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "clk", d["clocks"]["sm_mhz"], d["pipeline"]["chunk_frames"], d.get("parity_checked",{}).get("ok"))'
for a in "1024 1024" "512 512" "8192 4096"; do set -- $a
GASR_WAVE_TIMEOUT_S=20 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --utts $1 --wave $2 2>/dev/null | tail -1 | python -c "$summ"
done
} > gpurun_out/probe60.log 2>&1
echo done
