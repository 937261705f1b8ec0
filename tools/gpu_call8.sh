#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for t in "mlp_wide=0" "mlp_wide=1" "mlp_wide=2" "mlp_wide=3" "mlp_wide=2,hash_bw_mode=16" ; do
  ARN_TUNABLES=$t timeout 300 python bench.py --steps 64 --warmup 5 --train-only > gpurun_out/r2_t_$t.json 2> gpurun_out/r2_t_$t.err; echo "$t rc=$? $(grep value gpurun_out/r2_t_$t.json | cut -c1-200)"; tail -2 gpurun_out/r2_t_$t.err
done
ARN_NO_PREFETCH=1 ARN_TUNABLES=mlp_wide=3 timeout 300 python bench.py --steps 64 --warmup 5 --train-only > gpurun_out/r2_t_np3.json 2>/dev/null; echo "no-prefetch wide=3 $(grep value gpurun_out/r2_t_np3.json | cut -c1-200)"
ARN_NO_PREFETCH=1 ARN_TUNABLES=mlp_wide=0 timeout 300 python bench.py --steps 64 --warmup 5 --train-only > gpurun_out/r2_t_np0.json 2>/dev/null; echo "no-prefetch wide=0 $(grep value gpurun_out/r2_t_np0.json | cut -c1-200)"
