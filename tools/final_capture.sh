# Round-end measurement bundle (run under gpurun): bench line, ncu launch list of the same command, one full capture per kernel.
set -x
python bench.py > gpurun_out/bench_n1_r01c.json 2> gpurun_out/bench_n1_r01c.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/launches_l1023_r01c.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"legendre|fft_|get_nl|synth_prep|extract_td" -c 8 -f \
    -o gpurun_out/prof_all_r01c python tools/quick_rloop.py 1023 16 > gpurun_out/ncu_all_r01c.log 2>&1
tail -c 600 gpurun_out/bench_n1_r01c.json; tail -3 gpurun_out/ncu_all_r01c.log
