for e in "X=1" "ARN_SIDE_PRIO=-1" "ARN_FORK_STAGE=3" "ARN_FORK_STAGE=1"; do
  env $e timeout 200 python bench.py --train-only --no-refcuda --skip-w3 --skip-w4 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$e', round(d['ms_per_step'],4), round(d['ms_per_step_no_refresh'],4))"
done
