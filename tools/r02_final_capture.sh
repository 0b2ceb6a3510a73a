# Round-2 final measurement bundle (one GPU): golden diagnostics leg, default bench line, launch list of the same command, one
# full capture per kernel of a 32-level chunk at l_max=1023 (the bench's chunk size).
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_testOutputs.py tests/test_diagnostics.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r02r_diag_gpu.log; cat gpurun_out/r02r_diag_gpu.log
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"legendre|fft_|get_nl|synth_prep|extract_td" -c 7 -f \
    -o gpurun_out/r02r_prof_chunk32 python tools/quick_rloop.py 1023 32 32 > gpurun_out/r02r_ncu_chunk32.log 2>&1
tail -3 gpurun_out/r02r_ncu_chunk32.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02r_launches_l1023_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r02r_bench_under_ncu.log 2>&1
python bench.py > gpurun_out/r02r_bench_n1.json 2> gpurun_out/r02r_bench_n1.err
python tools/show_bench.py < gpurun_out/r02r_bench_n1.json
