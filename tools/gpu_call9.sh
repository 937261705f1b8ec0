#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for t in "mlp_wide=0" "mlp_wide=2" "mlp_wide=3"; do
  ARN_TUNABLES=$t timeout 300 python bench.py --steps 64 --warmup 5 --train-only > gpurun_out/r2_t_$t.json 2> gpurun_out/r2_t_$t.err; echo "$t rc=$? $(grep value gpurun_out/r2_t_$t.json | cut -c1-200)"; tail -2 gpurun_out/r2_t_$t.err
done
for f in 1 3; do
ARN_FORK_STAGE=$f ARN_TUNABLES=mlp_wide=3 timeout 300 python bench.py --steps 64 --warmup 5 --train-only > gpurun_out/r2_t_f$f.json 2>/dev/null; echo "fork $f wide=3 $(grep value gpurun_out/r2_t_f$f.json | cut -c1-200)"
done
timeout 600 python -m pytest tests -m gpu -q -x -k "field or render_train or fused_train" > gpurun_out/r2_pytest_sel.log 2>&1; echo "pytest-sel rc=$?"; tail -2 gpurun_out/r2_pytest_sel.log
