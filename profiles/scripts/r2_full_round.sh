mkdir -p gpurun_out
T=${TAG:-a}
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r2_gputest_$T.log
timeout 200 python bench.py --workload visual-cube-single --steps 30 --warmup 5 --no-cpu-baseline --no-fp32-leg --no-scaling-configs 2> gpurun_out/r2_pix_bench_$T.err | tee gpurun_out/r2_pix_bench_$T.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pixel ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'])"
timeout 200 python bench.py --workload puzzle-4x4 --seeds 8 --steps 30 --warmup 5 --no-cpu-baseline --no-fp32-leg --no-scaling-configs 2> gpurun_out/r2_p8_$T.err | tee gpurun_out/r2_p8_$T.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('puzzle 8 seeds ms/step', d['ms_per_step'])"
