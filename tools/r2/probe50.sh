#!/bin/bash
# four GPUs: pinned input buffers on the GPU's NUMA node (on / off)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
ls /sys/devices/system/node | grep node; nvidia-smi topo -m 2>/dev/null | head -12; cat /sys/fs/cgroup/cpuset.mems.effective 2>/dev/null; python -c "import os; print(sorted(os.sched_getaffinity(0)))"
summ='import json,sys
d=json.loads(sys.stdin.read()); print("gpus", d["n_gpus"], "value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "e2e ms", round(d["e2e"]["ms_per_step"],2), "node", d["config"].get("pinned_host_memory_numa_node_rank0"))'
for numa in 1 0 1 0; do
GASR_BENCH_NUMA=$numa GASR_WAVE_TIMEOUT_S=30 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2952$numa bench.py --gpus 4 --steps 5 --warmup 2 --no-checks > gpurun_out/tmp_n4.json 2> gpurun_out/tmp_n4.err; echo -n "numa=$numa rc=$? "
tail -1 gpurun_out/tmp_n4.json | python -c "$summ"
done
} > gpurun_out/probe50.log 2>&1
echo done
