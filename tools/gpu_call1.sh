#!/bin/bash
# GPU call: parity tests (report mode first, then the asserting run), field error report, 1-GPU bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
ARN_PARITY_REPORT=1 timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/r2_pytest_report.log 2>&1; echo "report rc=$?" 
timeout 300 python tests/report_field_error.py > gpurun_out/r2_field_error.log 2>&1; echo "field_error rc=$?"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_pytest.log; tail -c 1500 gpurun_out/r2_bench1.json; tail -5 gpurun_out/r2_bench1.err
