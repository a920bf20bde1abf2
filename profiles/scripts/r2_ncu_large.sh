# launch lists (ncu, serialised + cold: compare shares) of one steady-state step at the two large configurations
set -x
python bench.py --workload humanoidmaze-medium --batch 16384 --steps 4 --warmup 3 --no-cpu-baseline --no-fp32-leg --no-scaling-configs > gpurun_out/r2_plain_h16384.log 2>&1 &&
FQL_B200_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 260 --csv --log-file gpurun_out/r2_launches_h16384.csv python bench.py --workload humanoidmaze-medium --batch 16384 --steps 4 --warmup 3 --no-cpu-baseline --no-fp32-leg --no-scaling-configs > gpurun_out/r2_ncu_h16384.log 2>&1
python bench.py --workload puzzle-4x4 --batch 256 --seeds 64 --steps 4 --warmup 3 --no-cpu-baseline --no-fp32-leg --no-scaling-configs > gpurun_out/r2_plain_p64.log 2>&1 &&
FQL_B200_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 260 --csv --log-file gpurun_out/r2_launches_p64.csv python bench.py --workload puzzle-4x4 --batch 256 --seeds 64 --steps 4 --warmup 3 --no-cpu-baseline --no-fp32-leg --no-scaling-configs > gpurun_out/r2_ncu_p64.log 2>&1
tail -3 gpurun_out/r2_ncu_h16384.log gpurun_out/r2_ncu_p64.log
