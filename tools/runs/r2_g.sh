#!/bin/bash
# round 2, call G (2 GPUs): multi-GPU bench with parity check + eager DDP reference leg, trace at 64 images per rank
mkdir -p gpurun_out
export CB200_BENCH_WATCHDOG=300
echo "== bench N=2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err
echo "rc=$?"; cut -c1-2500 gpurun_out/r2g_bench_n2.json; grep -E "bench rank 0|Error|error|File \"/root" gpurun_out/r2g_bench_n2.err | tail -20
echo "== trace N=2, 64 per rank"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/trace_step.py --out gpurun_out/r2g_trace_n2_b64 --global-batch 128 2> gpurun_out/r2g_trace.err | cut -c1-500
tail -3 gpurun_out/r2g_trace.err
echo "== config1 strict"
timeout 300 python -m pytest tests/test_gpu_model.py -x -q -k "config1" 2>&1 | grep -E "assert|Error|passed|failed|step [12]" | head -20
echo "== dropin worker"
timeout 600 python -m pytest tests/test_gpu_dropin.py -x -q -s 2>&1 | tail -15
