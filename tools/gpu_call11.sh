#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "field or hash or render" > gpurun_out/r2_pytest_sel.log 2>&1; echo "pytest-sel rc=$?"; tail -2 gpurun_out/r2_pytest_sel.log
timeout 300 python bench.py --steps 64 --warmup 5 --train-only > gpurun_out/r2_t_pairfw.json 2>/dev/null; echo "train-only $(grep value gpurun_out/r2_t_pairfw.json | cut -c1-200)"
timeout 300 python tools/frame_rate.py 2>&1 | grep "one loop"
