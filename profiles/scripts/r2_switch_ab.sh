# A/B of the schedule switches on the headline configuration after the warp-uniform issue change
mkdir -p gpurun_out
run() { env "$@" timeout 200 python bench.py --steps 300 --warmup 30 --no-cpu-baseline --no-fp32-leg --no-scaling-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', 'ms/step', round(d['ms_per_step'],5), 'e2e', round(d['e2e']['value']))"; }
run A=default
run FQL_B200_CRITIC_CHAIN=1
run FQL_B200_SPLIT_ADAM=1
run FQL_B200_SPLIT_ADAM=2
run FQL_B200_FUSED_PREP=1
