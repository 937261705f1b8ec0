#!/bin/bash
# round-2 profiles: launch list of bench.py (training steps), ncu --set full of the step's kernels, launch list of a test frame
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export ARN_BENCH_MIN_WARMUP=48
python bench.py --steps 16 --warmup 3 --train-only > gpurun_out/r02_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 16 --warmup 3 --train-only > gpurun_out/r02_ncu_bench.log 2>&1
echo "launch list rc=$?"
python tools/ncu_step.py 6 > gpurun_out/r02_plain_step.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"hash_encode|field_mlp|adam_vec4|composite_train|march_train_count" -s 35 -c 8 -o gpurun_out/r02_prof_step python tools/ncu_step.py 6 > gpurun_out/r02_ncu_step.log 2>&1
echo "full capture rc=$?"
python tools/ncu_frame_sg.py > gpurun_out/r02_plain_frame.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches_frame.csv python tools/ncu_frame_sg.py > gpurun_out/r02_ncu_frame.log 2>&1
echo "frame launch list rc=$?"
ls -la gpurun_out | grep r02
