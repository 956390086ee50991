#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python bench.py --steps 40 --warmup 10 > gpurun_out/bench_r1t.json 2> gpurun_out/bench_r1t.err; cat gpurun_out/bench_r1t.json | cut -c1-400; tail -3 gpurun_out/bench_r1t.err
timeout 200 python tools/host_profile.py 2>&1 | head -24
timeout 100 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
