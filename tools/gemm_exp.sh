for v in "" _r1 _r4; do
  echo "=== variant '$v'"
  MAGIC_B200_LIB=$PWD/magic_b200/libmagic_b200$v.so timeout 300 python tools/quick_rloop.py 1023 16 2>&1 | tail -2
done
