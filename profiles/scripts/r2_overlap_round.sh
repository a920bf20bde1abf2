mkdir -p gpurun_out
T=${TAG:-a}
timeout 500 python -m pytest tests/test_overlap_gpu.py tests/test_step_gpu.py tests/test_checkpoint_gpu.py -m gpu -q -x 2>&1 | tail -6
for OV in 0 1; do
FQL_B200_OVERLAP_H2D=$OV timeout 200 python bench.py --steps 300 --warmup 30 --no-cpu-baseline --no-fp32-leg --no-scaling-configs 2> gpurun_out/r2_b256_ov${OV}_$T.err | tee gpurun_out/r2_b256_ov${OV}_$T.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('overlap=$OV B=256 ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'e2e ms', 256e3/d['e2e']['value'])"
FQL_B200_OVERLAP_H2D=$OV timeout 200 python bench.py --workload visual-cube-single --steps 40 --warmup 8 --no-cpu-baseline --no-fp32-leg --no-scaling-configs 2> gpurun_out/r2_pix_ov${OV}_$T.err | tee gpurun_out/r2_pix_ov${OV}_$T.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('overlap=$OV pixel ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'e2e ms', 256e3/d['e2e']['value'])"
done
