mkdir -p gpurun_out
# plain run first (must exit 0 without ncu)
python bench.py --workload humanoidmaze-medium --batch 16384 --steps 3 --warmup 3 --no-cpu-baseline --no-fp32-leg --no-scaling-configs > gpurun_out/r2_plain_h16384.log 2>&1 || { tail -5 gpurun_out/r2_plain_h16384.log; exit 1; }
FQL_B200_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_chain2 -s 30 -c 10 -o gpurun_out/r2_chain2_full -f python bench.py --workload humanoidmaze-medium --batch 16384 --steps 3 --warmup 3 --no-cpu-baseline --no-fp32-leg --no-scaling-configs > gpurun_out/r2_ncu_chain2.log 2>&1
tail -3 gpurun_out/r2_ncu_chain2.log
ls -la gpurun_out/
