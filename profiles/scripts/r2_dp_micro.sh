mkdir -p gpurun_out
N=${N:-2}
for cfg in "96 4 512" "96 8 512" "148 4 512" "32 8 512" "32 4 256" "16 8 256" "64 1 512" "148 8 1024"; do
  set -- $cfg
  FQL_DP_CTAS=$1 FQL_DP_UNROLL=$2 FQL_DP_THREADS=$3 timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 profiles/micro/dp_reduce_bench.py 2>&1 | grep "ctas="
done
FQL_DP_MULTICAST=0 timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 profiles/micro/dp_reduce_bench.py 2>&1 | grep "ctas="
