mkdir -p gpurun_out
T=${TAG:-a}
timeout 400 python -m pytest tests/test_tc_large_gpu.py -m gpu -q -x 2>&1 | tail -4
for B in 16384 8192; do
timeout 200 python bench.py --workload humanoidmaze-medium --batch $B --steps 20 --warmup 5 --no-cpu-baseline --no-fp32-leg --no-scaling-configs 2> gpurun_out/r2_h${B}_$T.err | tee gpurun_out/r2_h${B}_$T.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('humanoid B=$B ms/step', d['ms_per_step'], 'launches', d['gpu_launches_per_step'])"
done
timeout 200 python bench.py --workload puzzle-4x4 --seeds 64 --steps 10 --warmup 3 --no-cpu-baseline --no-fp32-leg --no-scaling-configs 2> gpurun_out/r2_p64_$T.err | tee gpurun_out/r2_p64_$T.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('puzzle 64 seeds ms/step', d['ms_per_step'])"
