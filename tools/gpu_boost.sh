#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
(timeout 300 python tools/frame_breakdown.py 8 1 16; timeout 300 python tools/frame_breakdown.py 1 1 4) > gpurun_out/r02_frame_breakdown_boost.txt 2> gpurun_out/r02_frame_breakdown_boost.err; echo "rc=$?"; cat gpurun_out/r02_frame_breakdown_boost.txt; tail -5 gpurun_out/r02_frame_breakdown_boost.err
