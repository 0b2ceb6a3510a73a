#!/bin/bash
# First GPU call of the next round: run the GPU tests that were written after round 1's GPU budget was spent (marker
# gpu_unverified: precession and condICrotIC golden runs with the CUDA loop, full-size property tests), then the regular
# suite, then the bench line.  Usage:  gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
set -u
mkdir -p gpurun_out
export MAGIC_UNVERIFIED_GPU=1
timeout 900 python -m pytest tests -m gpu -q -k "precession or condICrotIC or varCond or varProps or doubleDiffusion or boussBenchSat or full_size" -x --durations=15 > gpurun_out/pytest_unverified.log 2>&1
echo "unverified rc=$?" | tee -a gpurun_out/pytest_unverified.log
tail -25 gpurun_out/pytest_unverified.log
unset MAGIC_UNVERIFIED_GPU
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
echo "regular rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 600 gpurun_out/bench_n1.json
