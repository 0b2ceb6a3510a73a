# Round-end measurement bundle (run under gpurun on one GPU): GPU test suite, smoke, bench line, ncu launch list of the same
# command, one full capture per kernel of a 16-level chunk at l_max=1023.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/smoke_final.log 2>&1
python bench.py > gpurun_out/bench_n1_final2.json 2> gpurun_out/bench_n1_final2.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/launches_l1023_final2.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"legendre|fft_|get_nl|synth_prep|extract_td" -c 8 -f \
    -o gpurun_out/prof_all_final2 python tools/quick_rloop.py 1023 16 > gpurun_out/ncu_all_final2.log 2>&1
cat gpurun_out/pytest_gpu_final.log; tail -2 gpurun_out/smoke_final.log; tail -1 gpurun_out/bench_n1_final2.json | python tools/show_bench.py
