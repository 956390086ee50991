#!/bin/bash
# round 2, call N (8 GPUs): bench N=8 (parity check, reference DDP leg in child processes), rank-0 kernel timeline,
# config 5 (StyleGAN2_512, nn.DataParallel over 8 GPUs) native and reference
mkdir -p gpurun_out
export CB200_BENCH_WATCHDOG=500
echo "== bench N=8"
timeout 560 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2n_bench_n8.json 2> gpurun_out/r2n_bench_n8.err
echo "rc=$?"; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2n_bench_n8.json'))
    print({k: d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e'], d.get('eager_gpu_baseline'), d.get('parity_check'))
except Exception as e: print("no line:", e)
PY
grep -E "bench rank 0|Error|Timeout" gpurun_out/r2n_bench_n8.err | tail -10
echo "== trace N=8"
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/trace_step.py --out gpurun_out/r2n_trace_n8 2> gpurun_out/r2n_trace.err | cut -c1-300
echo "== config 5 native (DataParallel x8)"
timeout 300 python tools/bench_sg2.py --data-parallel 8 --batch 64 --steps 16 --warmup 2 > gpurun_out/r2n_config5_native.json 2> gpurun_out/r2n_config5_native.err; tail -c 900 gpurun_out/r2n_config5_native.json; tail -3 gpurun_out/r2n_config5_native.err | cut -c1-300
echo "== config 5 reference (DataParallel x8)"
timeout 240 python tools/bench_sg2.py --impl reference --data-parallel 8 --batch 64 --steps 16 --warmup 2 > gpurun_out/r2n_config5_reference.json 2> gpurun_out/r2n_config5_reference.err; tail -c 600 gpurun_out/r2n_config5_reference.json; tail -3 gpurun_out/r2n_config5_reference.err | cut -c1-300
