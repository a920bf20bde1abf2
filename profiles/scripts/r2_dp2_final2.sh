mkdir -p gpurun_out
timeout 170 python -m pytest tests/test_dp_gpu.py -m gpu -q -x -k "data_parallel" 2>&1 | tail -5 | cut -c1-300 | tee gpurun_out/r2_dp_tests_final2.log
timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29537 bench.py --gpus 2 --steps 100 --warmup 10 --no-cpu-baseline --no-fp32-leg --no-scaling-configs 2> gpurun_out/r2_bench_n2_final2.err | tail -1 > gpurun_out/r2_bench_n2_final2.json
python -c "
import json
d = json.loads(open('gpurun_out/r2_bench_n2_final2.json').read().strip().splitlines()[-1])
print('N=2 ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'transport', d.get('dp_transport'))
"
