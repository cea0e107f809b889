timeout 300 python scripts/_dbg_thin.py 2>&1 | tail -8
