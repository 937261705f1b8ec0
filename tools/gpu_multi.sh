#!/bin/bash
# usage: gpu_multi.sh N "variant;variant;..."  where variant = "LG FORK" ; then one full bench with the last variant
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=$1; shift
IFS=';' read -ra VARS <<< "$1"
PORT=29600
for v in "${VARS[@]}"; do
  set -- $v; LG=$1; FK=$2; EXTRA=$3
  PORT=$((PORT+1))
  tag="n${N}_lg${LG}_f${FK}${EXTRA:+_$EXTRA}"
  env ARN_LEVEL_GROUPS=$LG ARN_FORK_STAGE=$FK $EXTRA timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus $N --steps 64 --warmup 5 --train-only > gpurun_out/r2_$tag.json 2> gpurun_out/r2_$tag.err
  echo "$tag rc=$? $(grep value gpurun_out/r2_$tag.json | cut -c1-330)"
done
if [ -n "$FULL" ]; then
  set -- $FULL; LG=$1; FK=$2
  env ARN_LEVEL_GROUPS=$LG ARN_FORK_STAGE=$FK timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29700 \
      bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_full_n$N.json 2> gpurun_out/r2_full_n$N.err
  echo "full rc=$?"; tail -c 2500 gpurun_out/r2_full_n$N.json; tail -5 gpurun_out/r2_full_n$N.err
fi
