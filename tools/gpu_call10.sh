#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest.log
for t in "pdl=1" "pdl=0"; do
  ARN_TUNABLES=$t timeout 300 python bench.py --steps 64 --warmup 5 --train-only > gpurun_out/r2_t_$t.json 2> gpurun_out/r2_t_$t.err; echo "$t rc=$? $(grep value gpurun_out/r2_t_$t.json | cut -c1-200)"; tail -2 gpurun_out/r2_t_$t.err
  ARN_TUNABLES=$t timeout 300 python tools/frame_rate.py 2>&1 | grep -v Grid | tee gpurun_out/r2_frame_$t.log
done
