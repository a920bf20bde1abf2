mkdir -p gpurun_out
python bench.py --workload visual-cube-single --steps 3 --warmup 3 --no-cpu-baseline --no-fp32-leg --no-scaling-configs > gpurun_out/r2_plain_pix.log 2>&1 &&
FQL_B200_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 2200 -c 900 --csv --log-file gpurun_out/r2_launches_pix.csv python bench.py --workload visual-cube-single --steps 3 --warmup 3 --no-cpu-baseline --no-fp32-leg --no-scaling-configs > gpurun_out/r2_ncu_pix.log 2>&1
tail -2 gpurun_out/r2_ncu_pix.log | cut -c1-200
