#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29741 tools/p2p_exchange_bench.py 2>&1 | grep -E "world|rror|symmetric" | head -5
ARN_P2P_NO_MULTICAST=1 $TR --master-port 29742 tools/p2p_exchange_bench.py 2>&1 | grep -E "world|rror|symmetric" | sed 's/^/[symm, no multicast] /' | head -5
ARN_P2P_BACKEND=ipc $TR --master-port 29743 tools/p2p_exchange_bench.py 2>&1 | grep -E "world|rror|symmetric" | sed 's/^/[ipc] /' | head -5
timeout 300 $TR --master-port 29744 bench.py --gpus $N --steps 64 --warmup 5 --train-only > gpurun_out/r2_mc_n$N.json 2> gpurun_out/r2_mc_n$N.err; echo "train-only multicast rc=$? $(grep value gpurun_out/r2_mc_n$N.json | cut -c1-220)"; grep -iE "error|Traceback|symmetric" gpurun_out/r2_mc_n$N.err | head -5
ARN_P2P_BACKEND=ipc timeout 300 $TR --master-port 29745 bench.py --gpus $N --steps 64 --warmup 5 --train-only > gpurun_out/r2_ipc_n$N.json 2> gpurun_out/r2_ipc_n$N.err; echo "train-only ipc rc=$? $(grep value gpurun_out/r2_ipc_n$N.json | cut -c1-220)"
