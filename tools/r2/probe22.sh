#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
echo "== bench default"
GASR_WAVE_TIMEOUT_S=20 timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?"; tail -3 gpurun_out/bench_n1.err
summ='import json,sys
d=json.loads(sys.stdin.read()); print("utts", d["config"]["utterances"], "wave", d["config"]["utterances_per_batch"], "value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],3))'
tail -1 gpurun_out/bench_n1.json | python -c "$summ"
echo "== shares: N=2 (4096), N=4 (2048), N=8 (1024)"
for a in "--utts 4096 --wave 4096" "--utts 4096 --wave 2048" "--utts 2048 --wave 2048" "--utts 2048 --wave 1024" "--utts 1024 --wave 1024" "--utts 1024 --wave 512"; do
GASR_WAVE_TIMEOUT_S=20 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-checks $a 2>/dev/null | tail -1 | python -c "$summ"
done
} > gpurun_out/probe22.log 2>&1
echo done
