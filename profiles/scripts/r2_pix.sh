mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q -x > gpurun_out/r2_conv_tc.log 2>&1; tail -3 gpurun_out/r2_conv_tc.log | cut -c1-300
python profiles/micro/conv_bench.py 256 2>&1 | tail -6
python bench.py --workload visual-cube-single --steps 20 --warmup 5 --no-cpu-baseline --no-fp32-leg --no-scaling-configs --precision bf16 > gpurun_out/r2_pix_bench_bf16.json 2> gpurun_out/r2_pix_bench_bf16.err
for f in bf16; do python -c "
import json
d=json.load(open('gpurun_out/r2_pix_bench_$f.json')); print('$f', d['ms_per_step'], d['value'], d['dtype'], d['gpu_launches_per_step'], d['e2e'])
" || tail -3 gpurun_out/r2_pix_bench_$f.err; done
