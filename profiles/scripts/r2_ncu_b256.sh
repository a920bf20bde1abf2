mkdir -p gpurun_out
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-fp32-leg --no-scaling-configs > gpurun_out/r2_plain_b256.log 2>&1 &&
FQL_B200_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 500 --csv --log-file gpurun_out/r2_launches_b256.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-fp32-leg --no-scaling-configs > gpurun_out/r2_ncu_b256.log 2>&1
tail -1 gpurun_out/r2_ncu_b256.log | cut -c1-200
