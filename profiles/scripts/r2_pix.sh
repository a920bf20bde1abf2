mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pixels_gpu.py -m gpu -q -s -k tensor_core > gpurun_out/r2_pix_tc.log 2>&1; grep -v "^$" gpurun_out/r2_pix_tc.log | grep "gradient leaves\|passed\|failed\|Error" | cut -c1-400
python bench.py --workload visual-cube-single --steps 4 --warmup 3 --no-cpu-baseline --no-fp32-leg --no-scaling-configs > gpurun_out/r2_plain_pix.log 2>&1 &&
FQL_B200_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 400 --csv --log-file gpurun_out/r2_launches_pix.csv python bench.py --workload visual-cube-single --steps 4 --warmup 3 --no-cpu-baseline --no-fp32-leg --no-scaling-configs > gpurun_out/r2_ncu_pix.log 2>&1
tail -2 gpurun_out/r2_ncu_pix.log | cut -c1-300
