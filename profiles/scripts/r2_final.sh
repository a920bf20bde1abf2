# final state of round 2: GPU tests, smoke, the default bench line, the batch-256 launch list and one --set full capture of the cluster kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r2_final_gputest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "^\*\|OMP_NUM" | tail -3 | tee gpurun_out/r2_final_smoke.log
timeout 600 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; tail -c 600 gpurun_out/r2_final_bench.json
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-fp32-leg --no-scaling-configs > gpurun_out/r2_plain_b256.log 2>&1 &&
FQL_B200_GRAPH=0 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 500 --csv --log-file gpurun_out/r2f_launches_b256.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-fp32-leg --no-scaling-configs > gpurun_out/r2_ncu_b256.log 2>&1
FQL_B200_GRAPH=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:euler_cluster_kernel -s 6 -c 3 -o gpurun_out/r2f_cluster_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fp32-leg --no-scaling-configs > gpurun_out/r2_ncu_cluster.log 2>&1
tail -1 gpurun_out/r2_ncu_cluster.log | cut -c1-200; ls -la gpurun_out/r2f_*
