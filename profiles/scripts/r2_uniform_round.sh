mkdir -p gpurun_out
T=${TAG:-a}
if [ -n "$DIAG" ]; then timeout 150 python profiles/dbg_chain2_stages.py 2>&1 | grep -v "^\*\|OMP_NUM" | tee gpurun_out/r2_chain2_stages_$T.txt | tail -40; fi
timeout 500 python -m pytest tests/test_tc_gpu.py -m gpu -q -x 2>&1 | tail -4
timeout 200 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-fp32-leg --no-scaling-configs 2> gpurun_out/r2_b256_$T.err | tee gpurun_out/r2_b256_$T.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B=256 ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'euler us', d['roofline']['us_per_launch'])"
if [ -n "$LARGE" ]; then timeout 200 python bench.py --workload humanoidmaze-medium --batch 8192 --steps 20 --warmup 5 --no-cpu-baseline --no-fp32-leg --no-scaling-configs 2> gpurun_out/r2_h8192_$T.err | tee gpurun_out/r2_h8192_$T.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('humanoid B=8192 ms/step', d['ms_per_step'])"; fi
