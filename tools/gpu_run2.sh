timeout 900 python -m pytest tests/ -q -m gpu -x -k "cross or stream or config5 or tcs or envelope or export" -o faulthandler_timeout=300 2>&1 | tail -3
timeout 300 python tools/gpu_tune.py cross_merge_env 8192 200 2>&1 | tail -1 | cut -c1-180
TSP_PHASE_TIMING=1 timeout 300 python tools/gpu_tune.py cross_merge_env 8192 200 2>&1 | tail -2 | head -1 | grep -o "'stream'.*"
timeout 300 python tools/gpu_tune.py cross_merge_env 65536 200 2>&1 | tail -1 | cut -c1-180
