timeout 300 python tools/e2e_probe.py 2>&1 | tail -14
