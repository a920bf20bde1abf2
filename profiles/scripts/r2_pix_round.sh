mkdir -p gpurun_out
T=${TAG:-a}
timeout 300 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 120 python profiles/dbg_timeline_pix.py 2>&1 | grep -v "^\*\|OMP_NUM" | tee gpurun_out/r2_tl_pix_$T.log | tail -14
timeout 200 python bench.py --workload visual-cube-single --steps 30 --warmup 5 --no-cpu-baseline --no-fp32-leg --no-scaling-configs 2> gpurun_out/r2_pix_bench_$T.err | tee gpurun_out/r2_pix_bench_$T.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pixel ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches_per_step'])"
