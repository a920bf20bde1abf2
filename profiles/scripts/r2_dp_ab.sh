mkdir -p gpurun_out
N=${N:-2}
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 200 --warmup 20 --no-scaling-configs > gpurun_out/r2_dpab_${name}_n$N.json 2> gpurun_out/r2_dpab_${name}_n$N.err
  python -c "
import json
d=json.loads(open('gpurun_out/r2_dpab_${name}_n$N.json').read().strip().splitlines()[-1])
print('$name', 'ms', round(d['ms_per_step'],4), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches_per_step'], d.get('dp_transport'))
" || tail -5 gpurun_out/r2_dpab_${name}_n$N.err
}
run default X=1
run peer FQL_DP_MULTICAST=0
run bclate FQL_B200_DP_BC_LATE=1
