mkdir -p gpurun_out
for st in 0 2; do
FQL_B200_CHAIN2_STAGES=$st python bench.py --workload humanoidmaze-medium --batch 16384 --steps 10 --warmup 3 --no-cpu-baseline --no-fp32-leg --no-scaling-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('stages cap $st: step ms', round(d['ms_per_step'],3), 'euler kernel us', round(d['roofline'].get('us_per_launch',0),1))"
done
python -m pytest tests/test_step_gpu.py -m gpu -q -k "best_of_n" 2>&1 | tail -3
