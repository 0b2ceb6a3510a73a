timeout 300 python tools/quick_rloop.py 1023 16 2>&1 | tail -2
