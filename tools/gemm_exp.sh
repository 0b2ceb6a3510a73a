echo "=== frag skipping on"; timeout 300 python tools/quick_rloop.py 1023 16 2>&1 | tail -2
echo "=== frag skipping off"; MAGIC_POLAR_FRAG=0 timeout 300 python tools/quick_rloop.py 1023 16 2>&1 | tail -2
echo "=== 33 levels, chunk 16 (16 + 17)"; timeout 300 python tools/quick_rloop.py 1023 33 16 2>&1 | tail -2
