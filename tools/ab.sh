# usage: bash tools/ab.sh "tun1" "tun2" ...   (each an ARN_TUNABLES string; "-" = defaults)
mkdir -p gpurun_out
for t in "$@"; do
  if [ "$t" = "-" ]; then unset ARN_TUNABLES; else export ARN_TUNABLES="$t"; fi
  timeout 300 python bench.py --train-only --no-refcuda --skip-w3 --skip-w4 --no-cpu-baseline > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -5 gpurun_out/ab.err
  python - "$t" <<'PY'
import json,sys
d=json.load(open('gpurun_out/ab.json'))
k=d.get('kernel_ms_per_step', {})
print(sys.argv[1], 'ms/step %.4f no_refresh %.4f refresh %.4f'%(d['ms_per_step'], d.get('ms_per_step_no_refresh') or 0, d.get('refresh_ms') or 0), {a:round(b*1e3,1) for a,b in k.items() if b>0.02})
PY
done
