set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sht_gpu.py tests/test_rloop_gpu.py tests/test_truncations.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02g_tests.log
cat gpurun_out/r02g_tests.log
(bash tools/variant_probe.sh "" _nostw; MAGIC_FFT_TPC=16 bash tools/variant_probe.sh ""; MAGIC_FFT_TPC=4 bash tools/variant_probe.sh "") > gpurun_out/r02g_variants.log 2>&1
cat gpurun_out/r02g_variants.log
