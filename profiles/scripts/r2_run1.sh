python -m pytest tests/test_tc_gpu.py tests/test_tc_large_gpu.py tests/test_step_gpu.py tests/test_sampler_gpu.py -m gpu -q -s > gpurun_out/r2_gputest_c.log 2>&1; tail -30 gpurun_out/r2_gputest_c.log | cut -c1-250
B=16384 python profiles/dbg_timeline.py 2>&1 | head -14 > gpurun_out/r2_tl_16384_c.log; cat gpurun_out/r2_tl_16384_c.log
python bench.py --workload humanoidmaze-medium --batch 16384 --steps 20 --warmup 5 --no-cpu-baseline --no-fp32-leg --no-scaling-configs > gpurun_out/r2_c_h16384.json 2> gpurun_out/r2_c_h16384.err
python bench.py --workload puzzle-4x4 --batch 256 --seeds 64 --steps 20 --warmup 5 --no-cpu-baseline --no-fp32-leg --no-scaling-configs > gpurun_out/r2_c_p64.json 2> gpurun_out/r2_c_p64.err
for f in r2_c_h16384 r2_c_p64; do python -c "
import json
try:
    d=json.load(open('gpurun_out/$f.json')); print('$f', d['ms_per_step'], d['value'], d['roofline'].get('tensor_frac'))
except Exception as e: print('$f ERR', e)
"; tail -2 gpurun_out/$f.err; done
