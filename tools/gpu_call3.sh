#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest.log
timeout 300 python tools/bench_hash_bw.py > gpurun_out/r2_hash_bw_sweep.log 2>&1; echo "sweep rc=$?"; cat gpurun_out/r2_hash_bw_sweep.log | grep -v Grid
for g in "0,16" "0,8,16" "0,8,11,13,16"; do
  ARN_LEVEL_GROUPS_1GPU=$g timeout 300 python bench.py --steps 64 --warmup 5 --train-only > gpurun_out/r2_lg1_$g.json 2> gpurun_out/r2_lg1_$g.err; echo "groups $g rc=$? $(grep value gpurun_out/r2_lg1_$g.json | cut -c1-200)"
  ARN_NO_PIPELINED_OPT=1 ARN_LEVEL_GROUPS_1GPU=$g timeout 300 python bench.py --steps 64 --warmup 5 --train-only > gpurun_out/r2_lg1np_$g.json 2> gpurun_out/r2_lg1np_$g.err; echo "groups $g, optimizer behind rc=$? $(grep value gpurun_out/r2_lg1np_$g.json | cut -c1-200)"
done
