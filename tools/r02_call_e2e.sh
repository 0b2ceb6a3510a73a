set -x
mkdir -p gpurun_out
for w in dynamo_benchmark hydro_bench_anel full_sphere_l511; do
  timeout 300 python bench.py --workload $w --steps 5 --no-cpu > gpurun_out/r02k_bench_${w}.json 2> gpurun_out/r02k_bench_${w}.err
  tail -3 gpurun_out/r02k_bench_${w}.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02k_bench_${w}.json").read().strip().splitlines()[-1])
print("$w", round(d["ms_per_step"],3), d["e2e"]["variant"], d["e2e"]["variants"])
PY
done
timeout 500 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r02k_bench_n1.json 2> gpurun_out/r02k_bench_n1.err
tail -3 gpurun_out/r02k_bench_n1.err
python tools/show_bench.py < gpurun_out/r02k_bench_n1.json
python -c "
import json
d=json.loads(open('gpurun_out/r02k_bench_n1.json').read().strip().splitlines()[-1]); print(d['e2e'])"
if [ "$1" = "2" ]; then
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02k_bench_n2.json 2> gpurun_out/r02k_bench_n2.err
tail -3 gpurun_out/r02k_bench_n2.err
python tools/show_bench.py < gpurun_out/r02k_bench_n2.json
python -c "
import json
d=json.loads(open('gpurun_out/r02k_bench_n2.json').read().strip().splitlines()[-1]); print(d['e2e'])"
fi
