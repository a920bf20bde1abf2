python -m pytest tests -m gpu -q -x > gpurun_out/r2_gputest_d.log 2>&1; tail -5 gpurun_out/r2_gputest_d.log | cut -c1-300
python bench.py > gpurun_out/r2_d_bench.json 2> gpurun_out/r2_d_bench.err; tail -3 gpurun_out/r2_d_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2_d_bench.json'))
print('ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches_per_step'])
print('fp32', d.get('fp32_parity_mode'))
print('cpu', d.get('cpu_baseline'))
print('lat', d.get('sample_actions_latency'))
for s in d['scaling_configs']: print(s)
"
B=256 python profiles/dbg_timeline.py 2>&1 | head -30 > gpurun_out/r2_tl_256_d.log; cat gpurun_out/r2_tl_256_d.log
