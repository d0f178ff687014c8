#!/bin/bash
# eight GPUs (and four): the sharded bench with the gather check
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=index,name --format=csv | tail -8 | wc -l
summ='import json,sys
d=json.loads(sys.stdin.read()); print("gpus", d["n_gpus"], "utts", d["config"]["utterances"], "wave", d["config"]["utterances_per_batch"], "value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],3), d.get("gather"), d.get("parity_checked",{}).get("ok"), d["clocks"])'
for n in 8 4; do
GASR_WAVE_TIMEOUT_S=30 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r2_bench_n$n.json 2> gpurun_out/r2_bench_n$n.err; echo "rc=$?"; tail -2 gpurun_out/r2_bench_n$n.err | cut -c1-300
tail -1 gpurun_out/r2_bench_n$n.json | python -c "$summ"
done
} > gpurun_out/probe48.log 2>&1
echo done
