mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "field or fused_train or render_test or test_loop or hdr" 2>&1 | tail -4
timeout 300 python bench.py --no-refcuda --skip-w3 --skip-w4 --no-cpu-baseline > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/q_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/q_bench.json'))
print({k:d.get(k) for k in ('value','ms_per_step','ms_per_step_no_refresh','frames_per_s_800x800','frames_per_s_800x800_reference_schedule')})
print(d['kernel_ms_per_step'])
PY
