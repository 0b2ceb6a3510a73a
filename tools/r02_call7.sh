set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_truncations.py tests/test_sht_gpu.py tests/test_rloop_gpu.py tests/test_full_size_gpu.py tests/test_hydro_bench_anel.py tests/test_full_sphere.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r02z_tests.log; cat gpurun_out/r02z_tests.log
(bash tools/variant_probe.sh ""; for w in dynamo_benchmark hydro_bench_anel hydro_bench_anel_l85 bouss_dynamo_l255 full_sphere_l511 boussBenchSat_ckpt; do echo "== $w"; python bench.py --workload $w --steps 10 --no-cpu --no-e2e 2>/dev/null | python tools/show_bench.py; done) > gpurun_out/r02z_fft_rows_rule.log 2>&1; cat gpurun_out/r02z_fft_rows_rule.log
