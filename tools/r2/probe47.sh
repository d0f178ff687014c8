#!/bin/bash
# two GPUs: sharded bench with gather check, the 2-GPU test, reference arm under torchrun
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=index,name --format=csv
timeout 300 python -m pytest tests/test_gpu_sizes.py -x -q -m gpu -k "two_gpu" 2>&1 | tail -3
GASR_WAVE_TIMEOUT_S=30 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "rc=$?"; tail -2 gpurun_out/r2_bench_n2.err | cut -c1-300
summ='import json,sys
d=json.loads(sys.stdin.read()); print("gpus", d["n_gpus"], "utts", d["config"]["utterances"], "wave", d["config"]["utterances_per_batch"], "value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],3), d.get("gather"), d.get("parity_checked",{}).get("ok"))'
tail -1 gpurun_out/r2_bench_n2.json | python -c "$summ"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
} > gpurun_out/probe47.log 2>&1
echo done
