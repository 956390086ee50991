#!/bin/bash
# multi-GPU check: N ranks over NCCL (graphed step: engine gradient all-reduce + SyncBN(G) + packed embedding all-gather)
N=${1:-2}
mkdir -p gpurun_out
export CB200_BENCH_WATCHDOG=${CB200_BENCH_WATCHDOG:-100}
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 $2 > gpurun_out/bench_mgpu_$N.json 2> gpurun_out/bench_mgpu_$N.err
echo "rc=$?"; cut -c1-700 gpurun_out/bench_mgpu_$N.json; grep -E "bench rank|Error|error|File \"/root" gpurun_out/bench_mgpu_$N.err | head -40
