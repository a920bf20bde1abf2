mkdir -p gpurun_out
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 profiles/dbg_timeline_dp.py > gpurun_out/r2_tl_dp2.log 2>&1
grep -v "^\*\|OMP_NUM" gpurun_out/r2_tl_dp2.log | tail -50
timeout 300 python -m pytest tests/test_dp_gpu.py -m gpu -q -x > gpurun_out/r2_dp_tests.log 2>&1; tail -5 gpurun_out/r2_dp_tests.log | cut -c1-400
