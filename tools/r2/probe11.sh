#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
echo "== bench default"
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 6000 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
summ='import json,sys
d=json.loads(sys.stdin.read()); print("wave", d["config"]["utterances_per_batch"], "value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]))'
echo "== one GPU's share at N=8 (1024 utterances): one batch of 1024 vs two of 512 vs four of 256"
for wv in 1024 512 256; do
timeout 600 python bench.py --steps 3 --warmup 3 --utts 1024 --wave $wv --no-cpu-baseline --no-checks 2>&1 | tail -1 | python -c "$summ"
done
echo "== one GPU's share at N=4 (2048 utterances): 2048 vs 2 x 1024"
for wv in 2048 1024; do
timeout 600 python bench.py --steps 3 --warmup 3 --utts 2048 --wave $wv --no-cpu-baseline --no-checks 2>&1 | tail -1 | python -c "$summ"
done
} > gpurun_out/probe11.log 2>&1
echo done
