mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r2_gputest_f.log 2>&1; tail -4 gpurun_out/r2_gputest_f.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -3 gpurun_out/r2_smoke.log | cut -c1-300
python bench.py > gpurun_out/r2_f_bench.json 2> gpurun_out/r2_f_bench.err; tail -3 gpurun_out/r2_f_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2_f_bench.json'))
print('ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches_per_step'])
print('roofline', {k: d['roofline'][k] for k in ('bound','achieved','peak','frac','traffic')})
print('fp32', d.get('fp32_parity_mode',{}).get('ms_per_step'))
print('cpu', d.get('cpu_baseline'))
print('lat', d.get('sample_actions_latency'))
print('clocks', d.get('clocks'))
for s in d['scaling_configs']: print({k: s.get(k) for k in ('workload','mode','ms_per_step','value','tensor_frac','error')})
"
python bench.py --impl reference --steps 20 --warmup 3 2>/dev/null | cut -c1-400
