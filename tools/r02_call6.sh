set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rloop_gpu.py tests/test_sht_gpu.py tests/test_full_size_gpu.py tests/test_lm_side_gpu.py tests/test_truncations.py tests/test_hydro_bench_anel.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r02w_tests.log; cat gpurun_out/r02w_tests.log
(bash tools/variant_probe.sh ""; for w in bouss_dynamo_l255 full_sphere_l511 hydro_bench_anel; do for aw in 0 1; do echo "== $w MAGIC_GEMM_AN_WIDE=$aw"; MAGIC_GEMM_AN_WIDE=$aw python bench.py --workload $w --steps 10 --no-cpu --no-e2e 2>/dev/null | python tools/show_bench.py; done; done) > gpurun_out/r02w_variants_an.log 2>&1; cat gpurun_out/r02w_variants_an.log
