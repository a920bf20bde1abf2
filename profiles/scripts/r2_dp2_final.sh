mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_dp_gpu.py -m gpu -q -x -k "pixel or nccl" 2>&1 | tail -6 | cut -c1-300 | tee gpurun_out/r2_dp_pixel_tests.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 60 --warmup 10 --no-cpu-baseline --no-fp32-leg 2> gpurun_out/r2_bench_n2_final.err | tail -1 > gpurun_out/r2_bench_n2_final.json
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_n2_final.json').read().strip().splitlines()[-1])
print('N=2 ms/step', d['ms_per_step'], 'value', d['value'], 'transport', d.get('dp_transport'))
for c in d.get('scaling_configs', []):
    print(' ', c['workload'], c['mode'], 'ms/step', c['ms_per_step'], 'finite', c['finite'])
PY
