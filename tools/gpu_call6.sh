#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "hash_backward or level_grouped or field_backward or render_train_end_to_end or fused_train or pipelined" > gpurun_out/r2_pytest_sel.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_sel.log
timeout 300 python tools/bench_hash_bw_levels.py 2>&1 | grep -v Grid | tee gpurun_out/r2_hash_bw_levels2.log
timeout 300 python tools/bench_hash_bw.py 2>&1 | grep -E "blocks/SM|level groups" | tee gpurun_out/r2_hash_bw_sweep4.log
for t in "hash_bw_mode=1" "hash_bw_mode=16"; do
  ARN_TUNABLES=$t timeout 300 python bench.py --steps 64 --warmup 5 --train-only > gpurun_out/r2_t_$t.json 2> gpurun_out/r2_t_$t.err; echo "$t rc=$? $(grep value gpurun_out/r2_t_$t.json | cut -c1-200)"
done
