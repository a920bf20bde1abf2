mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q -x > gpurun_out/r2_conv_tc.log 2>&1; tail -12 gpurun_out/r2_conv_tc.log | cut -c1-400
python profiles/micro/conv_bench.py 256 2>&1 | tail -5
timeout 900 python -m pytest tests/test_pixels_gpu.py -m gpu -q -x -k "tensor_core and (b8 or 16px or 20px)" > gpurun_out/r2_pix_tc.log 2>&1; tail -3 gpurun_out/r2_pix_tc.log | cut -c1-300
python bench.py --workload visual-cube-single --steps 30 --warmup 5 --no-cpu-baseline --no-fp32-leg --no-scaling-configs --precision bf16 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pixels: ms', round(d['ms_per_step'],4), round(d['value']), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches_per_step'])"
