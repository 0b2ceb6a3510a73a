# Round-2 call 2 (one GPU): parity of the changed kernels, A/B of the FFT cache-hint / twiddle-hoist variants, full ncu capture.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sht_gpu.py tests/test_rloop_gpu.py tests/test_truncations.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02f_tests.log
cat gpurun_out/r02f_tests.log
bash tools/variant_probe.sh "" _nohint _nohoist > gpurun_out/r02f_variants.log 2>&1
cat gpurun_out/r02f_variants.log
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"legendre|fft_|get_nl|synth_prep|extract_td" -c 7 -f \
    -o gpurun_out/r02f_prof_all python tools/quick_rloop.py 1023 16 16 > gpurun_out/r02f_ncu_all.log 2>&1
tail -3 gpurun_out/r02f_ncu_all.log
