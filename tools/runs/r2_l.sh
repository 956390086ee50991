#!/bin/bash
# round 2, call L (2 GPUs): contiguous-buffer gradient all-reduce + reference leg in child processes
mkdir -p gpurun_out
export CB200_BENCH_WATCHDOG=600
echo "== bench N=2"
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2l_bench_n2.json 2> gpurun_out/r2l_bench_n2.err
echo "rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2l_bench_n2.json'))
print({k: d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e'], d.get('eager_gpu_baseline'), d.get('parity_check'))
PY
grep -E "bench rank 0|Error|error" gpurun_out/r2l_bench_n2.err | tail -10
echo "== trace N=2, 64 per rank"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/trace_step.py --out gpurun_out/r2l_trace_n2_b64 --global-batch 128 2> gpurun_out/r2l_trace.err | cut -c1-300
echo "== multi-GPU parity test"
timeout 400 python -m pytest tests/test_gpu_multi.py -x -q -s 2>&1 | tail -4 | cut -c1-600
