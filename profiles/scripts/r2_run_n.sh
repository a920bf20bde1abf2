N=${N:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
echo rc=$?; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2_bench_n$N.err | tail -5 | cut -c1-300
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_n$N.json').read().strip().splitlines()[-1])
print('ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches_per_step'], d.get('dp_transport'))
for s in d['scaling_configs']: print({k: s.get(k) for k in ('workload','mode','ms_per_step','value','tensor_frac','transport','error','finite')})
"
