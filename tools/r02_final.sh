# Round-2 closing bundle (one GPU): the driver's GPU suite, smoke, the default bench line and the reference arm.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r02_final_gpu_suite.log; cat gpurun_out/r02_final_gpu_suite.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1; tail -1 gpurun_out/r02_final_smoke.log
python bench.py > gpurun_out/r02_final_bench_n1.json 2> gpurun_out/r02_final_bench_n1.err; python tools/show_bench.py < gpurun_out/r02_final_bench_n1.json
python -c "
import json
d=json.loads(open('gpurun_out/r02_final_bench_n1.json').read().strip().splitlines()[-1])
print('e2e', d['e2e']['variant'], round(d['e2e']['ms_per_step'],1), 'traffic', d['roofline']['traffic'], 'frac', round(d['roofline']['frac'],3), 'cpu', d['cpu_baseline'])"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_final_reference_arm.json 2> gpurun_out/r02_final_reference_arm.err; tail -c 400 gpurun_out/r02_final_reference_arm.json
