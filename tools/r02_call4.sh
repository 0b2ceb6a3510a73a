set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_rloop_gpu.py tests/test_sht_gpu.py tests/test_full_size_gpu.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r02s_tests.log; cat gpurun_out/r02s_tests.log
bash tools/variant_probe.sh "" > gpurun_out/r02s_variants.log 2>&1; cat gpurun_out/r02s_variants.log
timeout 300 python bench.py --workload boussBenchSat_ckpt --steps 20 --no-cpu > gpurun_out/r02s_bench_boussBenchSat_ckpt.json 2> gpurun_out/r02s_bench_boussBenchSat_ckpt.err
tail -3 gpurun_out/r02s_bench_boussBenchSat_ckpt.err; python tools/show_bench.py < gpurun_out/r02s_bench_boussBenchSat_ckpt.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"legendre|fft_|get_nl|synth_prep|extract_td|lmside|rside|courant" -c 400 --csv \
    --log-file gpurun_out/r02s_launches_l1023_bench.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r02s_bench_under_ncu.log 2>&1
tail -2 gpurun_out/r02s_bench_under_ncu.log | cut -c1-300
