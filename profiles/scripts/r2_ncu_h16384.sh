# launch list (ncu, serialised + cold: compare shares) of one steady-state step at humanoidmaze-shaped batch 16384
python bench.py --workload humanoidmaze-medium --batch 16384 --steps 4 --warmup 3 --no-cpu-baseline --no-fp32-leg --no-scaling-configs > gpurun_out/r2_plain_h16384.log 2>&1 &&
FQL_B200_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 260 --csv --log-file gpurun_out/r2_launches_h16384_${TAG:-a}.csv python bench.py --workload humanoidmaze-medium --batch 16384 --steps 4 --warmup 3 --no-cpu-baseline --no-fp32-leg --no-scaling-configs > gpurun_out/r2_ncu_h16384.log 2>&1
tail -2 gpurun_out/r2_ncu_h16384.log | cut -c1-300
