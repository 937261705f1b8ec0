#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "field or render_train or fused_train or level_grouped or pipelined" > gpurun_out/r2_pytest_sel.log 2>&1; echo "pytest-sel rc=$?"; tail -4 gpurun_out/r2_pytest_sel.log
for t in "mlp_wide=1" "mlp_wide=0"; do
  ARN_TUNABLES=$t timeout 300 python bench.py --steps 64 --warmup 5 --train-only > gpurun_out/r2_t_$t.json 2> gpurun_out/r2_t_$t.err; echo "$t rc=$? $(grep value gpurun_out/r2_t_$t.json | cut -c1-200)"; tail -2 gpurun_out/r2_t_$t.err
done
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?"; python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench1.json').read().strip().splitlines()[-1])
for k in ("value","ms_per_step","ms_per_step_no_refresh","e2e","frames_per_s_800x800","refcuda","kernel_ms_per_step"):
    print(k, json.dumps(d[k])[:1500])
P
