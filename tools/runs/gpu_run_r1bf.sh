#!/bin/bash
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q -x 2>&1 | grep -E "^E|passed|failed|^FAILED|^ERROR" | head -20
timeout 200 python tools/bench_sg2.py > gpurun_out/bench_sg2.json 2> gpurun_out/bench_sg2.err; tail -2 gpurun_out/bench_sg2.err; cat gpurun_out/bench_sg2.json; cat gpurun_out/eager_gpu_sndcgan.json
