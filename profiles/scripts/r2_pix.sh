mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pixels_gpu.py -m gpu -q -s -k tensor_core > gpurun_out/r2_pix_tc.log 2>&1; grep -v "^$" gpurun_out/r2_pix_tc.log | grep "gradient leaves\|passed\|failed\|Error\|error" | cut -c1-400
python bench.py --workload visual-cube-single --steps 20 --warmup 5 --no-cpu-baseline --no-fp32-leg --no-scaling-configs --precision bf16 > gpurun_out/r2_pix_bench_bf16.json 2> gpurun_out/r2_pix_bench_bf16.err
for f in bf16; do python -c "
import json
d=json.load(open('gpurun_out/r2_pix_bench_$f.json')); print('$f', d['ms_per_step'], d['value'], d['dtype'], d['gpu_launches_per_step'], d['e2e'])
" || tail -3 gpurun_out/r2_pix_bench_$f.err; done
