#!/bin/bash
# round 2, call M (1 GPU): the whole GPU suite as the driver runs it, then smoke()
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -12 | cut -c1-400
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
