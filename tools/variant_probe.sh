#!/bin/bash
# A/B timing of experimental builds of the library (run under gpurun).  Build variants next to the default library with
#   python magic_b200/build.py --out=$PWD/magic_b200/libmagic_b200_<name>.so -D<MACRO>=<value>
# and pass their names: tools/variant_probe.sh "" _name1 _name2   ("" = the default library).  Prints the per-stage device
# times of one 16-level chunk at l_max=1023 and a hash of all outputs (variants must not change results).
for v in "$@"; do
  echo "=== variant '$v'"
  MAGIC_B200_LIB=$PWD/magic_b200/libmagic_b200$v.so timeout 300 python tools/quick_rloop.py 1023 16 2>&1 | tail -2
done
