#!/bin/bash
# two GPUs: the sharded bench (gather check) + the 2-GPU test; then single-GPU share checks of the groups-per-cluster rule
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=index,name --format=csv
timeout 300 python -m pytest tests/test_gpu_sizes.py -x -q -m gpu -k "two_gpu" 2>&1 | tail -3
echo "== bench --gpus 2"
GASR_WAVE_TIMEOUT_S=20 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?"; tail -3 gpurun_out/bench_n2.err
summ='import json,sys
d=json.loads(sys.stdin.read()); print("gpus", d["n_gpus"], "utts", d["config"]["utterances"], "wave", d["config"]["utterances_per_batch"], "value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],3), d.get("gather"), d.get("parity_checked",{}).get("ok"))'
tail -1 gpurun_out/bench_n2.json | python -c "$summ"
echo "== reference arm under torchrun"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
echo "== shares on one GPU with the groups-per-cluster rule"
for a in "--utts 2048 --wave 2048" "--utts 1024 --wave 1024"; do
GASR_WAVE_TIMEOUT_S=20 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-checks $a 2>/dev/null | tail -1 | python -c "$summ"
done
} > gpurun_out/probe23.log 2>&1
echo done
