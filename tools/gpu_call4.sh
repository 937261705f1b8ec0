#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/bench_hash_bw.py > gpurun_out/r2_hash_bw_sweep2.log 2>&1; echo "sweep rc=$?"; grep -E "level groups|blocks/SM [023] min_run  16" gpurun_out/r2_hash_bw_sweep2.log
for t in "hash_bw_blocks=0" "hash_bw_blocks=2" "hash_bw_blocks=3"; do
  ARN_TUNABLES=$t timeout 300 python bench.py --steps 64 --warmup 5 --train-only > gpurun_out/r2_t_$t.json 2> gpurun_out/r2_t_$t.err; echo "$t rc=$? $(grep value gpurun_out/r2_t_$t.json | cut -c1-160)"
done
timeout 900 python -m pytest tests -m gpu -q -k "render_train_end_to_end or level_grouped or hash_backward or fused_train" > gpurun_out/r2_pytest_sel.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_sel.log
