#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_sg2.py -q -x 2>&1 | grep -E "^E|passed|failed|^FAILED|^ERROR|Error" | head -20
timeout 200 python tools/bench_sg2.py > gpurun_out/bench_sg2_eager.json 2> gpurun_out/bench_sg2.err; tail -2 gpurun_out/bench_sg2.err; cut -c1-260 gpurun_out/bench_sg2_eager.json
timeout 200 python tools/bench_sg2.py --graph --steps 20 > gpurun_out/bench_sg2_graph.json 2> gpurun_out/bench_sg2g.err; tail -5 gpurun_out/bench_sg2g.err; cut -c1-260 gpurun_out/bench_sg2_graph.json
