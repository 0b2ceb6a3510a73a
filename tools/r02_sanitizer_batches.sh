#!/bin/bash
# memcheck / initcheck / synccheck over the GPU tests of the log-step batches (diagnostics, dtB, TO, RMS, shared workspace)
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck initcheck synccheck; do
  log=gpurun_out/sanitizer_batches_$tool.log
  ( time timeout 400 $CS --tool $tool --print-limit 20 python -m pytest tests/test_to.py tests/test_rms.py tests/test_diagnostics.py -m gpu -q -x -p no:cacheprovider ) > $log 2>&1
  echo "== $tool: $(grep -h 'ERROR SUMMARY\|passed\|failed' $log | tr '\n' ' ')"
  grep -h -m 12 "========= Invalid\|========= Uninitialized\|========= Error\|========= Barrier\|=========     at " $log | cut -c1-200 | sort | uniq -c | sort -rn | head -8
done
