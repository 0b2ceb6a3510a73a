# Round-2 measurement bundle (one GPU): bench line of every BASELINE config, the reference arm, the ncu launch list of the
# default bench command and one full capture per kernel of a level chunk at l_max=1023.
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_l1023_n1.json 2> gpurun_out/r02_bench_l1023_n1.err
tail -c 1500 gpurun_out/r02_bench_l1023_n1.json
python bench.py --level-chunk 32 --no-e2e --no-cpu > gpurun_out/r02_bench_l1023_n1_chunk32.json 2> gpurun_out/r02_bench_l1023_n1_chunk32.err
for w in dynamo_benchmark hydro_bench_anel hydro_bench_anel_l85 bouss_dynamo_l255 full_sphere_l511; do
  timeout 600 python bench.py --workload $w --steps 10 > gpurun_out/r02_bench_${w}_n1.json 2> gpurun_out/r02_bench_${w}_n1.err
  tail -c 400 gpurun_out/r02_bench_${w}_n1.json; tail -3 gpurun_out/r02_bench_${w}_n1.err
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_l1023_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r02_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"legendre|fft_|get_nl|synth_prep|extract_td" -c 8 -f \
    -o gpurun_out/r02_prof_all python tools/quick_rloop.py 1023 16 16 > gpurun_out/r02_ncu_all.log 2>&1
ls -la gpurun_out | tail -15
