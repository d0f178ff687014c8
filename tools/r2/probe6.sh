#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
echo "== bench N=1"
timeout 900 python bench.py --steps 2 --warmup 3 2>&1 | tail -3
echo "== wave sizes"
for wv in 512 2048; do
timeout 600 python bench.py --steps 2 --warmup 3 --wave $wv --no-cpu-baseline --no-checks 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('wave', d['config']['utterances_per_batch'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'stages', d['stages_ms_sum_of_launches'])"
done
} > gpurun_out/probe6.log 2>&1
echo done
