N=${N:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
echo rc=$?; tail -5 gpurun_out/r2_bench_n$N.err | cut -c1-300
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_n$N.json').read().strip().splitlines()[-1])
print('ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches_per_step'], d.get('dp_transport'))
for s in d['scaling_configs']: print(s)
"
timeout 300 python -m pytest tests/test_dp_gpu.py -m gpu -q -x 2>&1 | tail -3
