mkdir -p gpurun_out
python bench.py --workload puzzle-4x4 --batch 256 --seeds 8 --steps 4 --warmup 3 --no-cpu-baseline --no-fp32-leg --no-scaling-configs > gpurun_out/r2_plain_p8.log 2>&1 &&
FQL_B200_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 400 --csv --log-file gpurun_out/r2_launches_p8.csv python bench.py --workload puzzle-4x4 --batch 256 --seeds 8 --steps 4 --warmup 3 --no-cpu-baseline --no-fp32-leg --no-scaling-configs > gpurun_out/r2_ncu_p8.log 2>&1
tail -1 gpurun_out/r2_ncu_p8.log | cut -c1-200
