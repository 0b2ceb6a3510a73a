set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rloop_gpu.py tests/test_sht_gpu.py tests/test_full_size_gpu.py tests/test_lm_side_gpu.py tests/test_diagnostics.py tests/test_dynamo_benchmark.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r02v_tests.log; cat gpurun_out/r02v_tests.log
(bash tools/variant_probe.sh "" _an128; for v in "" _an128; do echo "=== 32-level chunk, variant '$v'"; MAGIC_B200_LIB=$PWD/magic_b200/libmagic_b200$v.so timeout 300 python tools/quick_rloop.py 1023 32 32 2>&1 | tail -2; done) > gpurun_out/r02v_variants_an.log 2>&1; cat gpurun_out/r02v_variants_an.log
for w in bouss_dynamo_l255 full_sphere_l511; do for v in "" _an128; do echo "== $w '$v'"; MAGIC_B200_LIB=$PWD/magic_b200/libmagic_b200$v.so python bench.py --workload $w --steps 10 --no-cpu --no-e2e 2>/dev/null | python tools/show_bench.py; done; done 2>&1 | tee -a gpurun_out/r02v_variants_an.log
