mkdir -p gpurun_out
# large-batch chain kernels (B=16384 humanoidmaze): plain run first, then one --set full capture of each chain2 instantiation
python bench.py --workload humanoidmaze-medium --batch 16384 --steps 3 --warmup 3 --no-cpu-baseline --no-fp32-leg --no-scaling-configs > gpurun_out/r2_plain_h16384.log 2>&1 || { tail -5 gpurun_out/r2_plain_h16384.log; exit 1; }
FQL_B200_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_chain2 -s 16 -c 8 -o gpurun_out/r2_chain2_full -f python bench.py --workload humanoidmaze-medium --batch 16384 --steps 3 --warmup 3 --no-cpu-baseline --no-fp32-leg --no-scaling-configs > gpurun_out/r2_ncu_chain2.log 2>&1
tail -2 gpurun_out/r2_ncu_chain2.log | cut -c1-200
# tensor-core convolutions (config 5, B=256)
python bench.py --workload visual-cube-single --steps 3 --warmup 3 --no-cpu-baseline --no-fp32-leg --no-scaling-configs > gpurun_out/r2_plain_pix.log 2>&1 || { tail -5 gpurun_out/r2_plain_pix.log; exit 1; }
FQL_B200_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 130 -c 26 -o gpurun_out/r2_conv_full -f python bench.py --workload visual-cube-single --steps 3 --warmup 3 --no-cpu-baseline --no-fp32-leg --no-scaling-configs > gpurun_out/r2_ncu_conv.log 2>&1
tail -2 gpurun_out/r2_ncu_conv.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
