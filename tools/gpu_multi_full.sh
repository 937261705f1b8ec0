#!/bin/bash
# usage: gpu_multi_full.sh N  -- the driver's command line for the N-GPU bench, reference arm first
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
echo "bench N=$N rc=$?"; tail -3 gpurun_out/r02_bench_${N}gpu.err | cut -c1-300
python - <<P
import json
d=json.loads(open('gpurun_out/r02_bench_${N}gpu.json').read().strip().splitlines()[-1])
for k in ("value","ms_per_step","ms_per_step_no_refresh","frames_per_s_800x800","other_configs","multi_gpu_check","roofline","kernel_ms_per_step"):
    print(k, json.dumps(d[k])[:700])
print("e2e", d["e2e"]["value"])
P
