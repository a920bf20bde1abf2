mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dp_gpu.py -m gpu -q -x > gpurun_out/r2_dp_tests.log 2>&1; tail -25 gpurun_out/r2_dp_tests.log | cut -c1-400
