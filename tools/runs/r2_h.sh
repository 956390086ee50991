#!/bin/bash
# round 2, call H (8 GPUs): bench N=8 (parity check + eager DDP reference) and the kernel timeline of rank 0
mkdir -p gpurun_out
export CB200_BENCH_WATCHDOG=300
echo "== bench N=8"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2h_bench_n8.json 2> gpurun_out/r2h_bench_n8.err
echo "rc=$?"; cut -c1-2600 gpurun_out/r2h_bench_n8.json; grep -E "bench rank 0|Error|error|File \"/root" gpurun_out/r2h_bench_n8.err | tail -12
echo "== trace N=8"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/trace_step.py --out gpurun_out/r2h_trace_n8 2> gpurun_out/r2h_trace.err | cut -c1-400
tail -2 gpurun_out/r2h_trace.err
nvidia-smi topo -m 2>/dev/null | head -12 > gpurun_out/r2h_topo.txt
