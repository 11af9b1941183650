for plan in "r:1:25:0" "r:37:25:1"; do echo "== $plan"; timeout 20 python tools/repro_hang.py "$plan"; echo "rc=$?"; done > gpurun_out/t12_repro.log 2>&1
( time timeout 1000 python -m pytest tests/ -m gpu -x -q --durations=12 -o faulthandler_timeout=300 ) > gpurun_out/t12_pytest_full.log 2>&1
