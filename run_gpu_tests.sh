timeout 300 python tools/_cmp_tmp.py 2>&1 | tail -30
