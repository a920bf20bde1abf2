mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_large_gpu.py tests/test_tc_gpu.py -m gpu -q -x > gpurun_out/r2_gputest_e.log 2>&1; tail -6 gpurun_out/r2_gputest_e.log | cut -c1-400
python bench.py --workload humanoidmaze-medium --batch 16384 --steps 20 --warmup 5 --no-cpu-baseline --no-fp32-leg --no-scaling-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B=16384 ms', round(d['ms_per_step'],4), round(d['value']))"
