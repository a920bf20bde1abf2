for m in 0 1 2; do
FQL_B200_SPLIT_ADAM=$m python bench.py --steps 300 --warmup 20 --no-cpu-baseline --no-fp32-leg --no-scaling-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('split_adam $m: ms', round(d['ms_per_step'],4), 'launches', d['gpu_launches_per_step'])"
done
