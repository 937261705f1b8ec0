mkdir -p gpurun_out
S=$(date +%s)
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$? $(( $(date +%s)-S ))s"; tail -3 gpurun_out/r02_pytest_gpu.log
S=$(date +%s)
python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$? $(( $(date +%s)-S ))s"
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_1gpu.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','frames_per_s_800x800','frames_per_s_800x800_reference_schedule','roofline')})
print(d['refcuda']['test_frame_ms'] if d.get('refcuda') else None, d['refcuda'].get('pixels_identical'), d['refcuda'].get('max_abs_rgb_diff_production'))
"
