#!/bin/bash
# racecheck of one pytest selection for the default library and variants (names as for tools/variant_probe.sh); hazards per kernel
# usage: tools/r02_racecheck.sh "<pytest -k expression>" "" _variant ...
mkdir -p gpurun_out
sel=$1; shift
for v in "$@"; do
  log=gpurun_out/racecheck$v.log
  ( time MAGIC_B200_LIB=$PWD/magic_b200/libmagic_b200$v.so timeout 400 /usr/local/cuda/bin/compute-sanitizer --tool racecheck --racecheck-report all --print-limit 5000 \
      python -m pytest tests/test_rloop_gpu.py -m gpu -q -x -k "$sel" -p no:cacheprovider ) > $log 2>&1
  echo "=== racecheck variant '$v': $(grep -c 'hazard detected' $log) hazard records; $(grep 'RACECHECK SUMMARY\|passed\|failed\|^real' $log | tr '\n' ' ')"
  grep -A2 "hazard detected" $log | grep "Read Thread\|Write Thread" | sed 's/Thread ([0-9,]*)/Thread/; s/+0x[0-9a-f]*//' | cut -c1-200 | sort | uniq -c | sort -rn | head -12
  grep "hazard detected" $log | sed 's/at __shared__ 0x[0-9a-f]*//' | sort | uniq -c | sort -rn | head -20
done
