#!/bin/bash
# GPU call: parity tests (report mode, then asserting), 1-GPU bench, 1-GPU level-group pipelining A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
ARN_PARITY_REPORT=1 timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/r2_pytest_report.log 2>&1; echo "report rc=$?"
timeout 300 python tests/report_field_error.py > gpurun_out/r2_field_error.log 2>&1; echo "field_error rc=$?"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?"
for g in "0,16" "0,8,16" "0,8,11,13,16" "0,6,9,11,12,13,14,15,16"; do
  ARN_LEVEL_GROUPS_1GPU=$g timeout 300 python bench.py --steps 64 --warmup 5 --train-only > gpurun_out/r2_lg1_$g.json 2> gpurun_out/r2_lg1_$g.err; echo "groups $g rc=$? $(cat gpurun_out/r2_lg1_$g.json | cut -c1-400)"
done
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_pytest.log; tail -5 gpurun_out/r2_bench1.err
