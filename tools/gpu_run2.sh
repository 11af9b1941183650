# developer scratch: bring-up of the streamed tcgen05 kernel (cross-merge)
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cross_merge or two_players" -o faulthandler_timeout=120 2>&1 | tail -30
timeout 120 python tools/gpu_tune.py cross_merge_env 8192 200
TSP_PHASE_TIMING=1 timeout 120 python tools/gpu_tune.py cross_merge_env 8192 200
timeout 120 python tools/gpu_tune.py cross_merge_env 4096 50
TSP_NO_TCS=1 timeout 120 python tools/gpu_tune.py cross_merge_env 8192 200
