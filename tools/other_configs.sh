# bench lines of the other BASELINE configs (parity-test cases; the bench contract line is dynamo_l1023)
for w in dynamo_benchmark hydro_bench_anel bouss_dynamo_l255 full_sphere_l511; do
  timeout 600 python bench.py --workload $w --no-cpu > gpurun_out/bench_${w}_r01c.json 2> gpurun_out/bench_${w}_r01c.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_${w}_r01c.json"))
print("$w", "ms/step", round(d["ms_per_step"],3), "TF/s", round(d["value"]/1e3,3), "e2e ms", round(d["e2e"]["ms_per_step"],3), "gemm frac", round(d["roofline"]["frac"],3), "chunk", d["config"]["level_chunk"], {k: round(v,2) for k,v in d["stages_ms"].items()})
PY
done
