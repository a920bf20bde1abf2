mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_tc_large_gpu.py -m gpu -q -x > gpurun_out/r2_gputest_e.log 2>&1; tail -3 gpurun_out/r2_gputest_e.log | cut -c1-300
python bench.py --workload humanoidmaze-medium --batch 16384 --steps 20 --warmup 5 --no-cpu-baseline --no-fp32-leg --no-scaling-configs > gpurun_out/r2_e_h16384.json 2> gpurun_out/r2_e_h16384.err
python bench.py --workload puzzle-4x4 --batch 256 --seeds 64 --steps 20 --warmup 5 --no-cpu-baseline --no-fp32-leg --no-scaling-configs > gpurun_out/r2_e_p64.json 2> gpurun_out/r2_e_p64.err
python bench.py --workload puzzle-4x4 --batch 256 --seeds 8 --steps 20 --warmup 5 --no-cpu-baseline --no-fp32-leg --no-scaling-configs > gpurun_out/r2_e_p8.json 2> gpurun_out/r2_e_p8.err
for f in r2_e_h16384 r2_e_p64 r2_e_p8; do python -c "
import json
try:
    d=json.load(open('gpurun_out/$f.json')); print('$f', round(d['ms_per_step'],4), round(d['value']), d['roofline'].get('tensor_frac'))
except Exception as e: print('$f ERR', e)
"; tail -2 gpurun_out/$f.err; done
