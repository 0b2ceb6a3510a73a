# NCCL channel-count probe at N=$1: fewer p2p channels = fewer SMs taken from the concurrent GEMM / FFT kernels
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29530
for cfg in "" "NCCL_MAX_P2P_NCHANNELS=2" "NCCL_MAX_P2P_NCHANNELS=4" "NCCL_MAX_P2P_NCHANNELS=8" "NCCL_MAX_P2P_NCHANNELS=2 NCCL_MAX_NCHANNELS=2"; do
  port=$((port+1))
  echo "=== $cfg"
  env $cfg timeout 300 $TR --master-port $port bench.py --gpus $N --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/nccl_probe.json 2> gpurun_out/nccl_probe.err
  python tools/show_bench.py < gpurun_out/nccl_probe.json || tail -3 gpurun_out/nccl_probe.err
done
