set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r02l_bench_n8.json 2> gpurun_out/r02l_bench_n8.err
python tools/show_bench.py < gpurun_out/r02l_bench_n8.json
python -c "
import json
d=json.loads(open('gpurun_out/r02l_bench_n8.json').read().strip().splitlines()[-1]); print(d['e2e']['variant'], d['e2e']['variants'])"
MAGIC_TRANSP_CE=1 timeout 300 $TR --master-port 29522 bench.py --gpus 8 --steps 3 --warmup 3 --no-e2e > gpurun_out/r02l_bench_n8_ce.json 2> gpurun_out/r02l_bench_n8_ce.err
python tools/show_bench.py < gpurun_out/r02l_bench_n8_ce.json
tail -3 gpurun_out/r02l_bench_n8_ce.err
