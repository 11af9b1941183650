# developer scratch: what the last gpurun call of the session ran (search-kernel time per config + the GPU test suite)
timeout 120 python tools/gpu_tune.py roundabout_env_ttc_flat 4096 50
timeout 120 python tools/gpu_tune.py highway_env_ttc_flat 65536 50
timeout 600 python -m pytest tests/ -m gpu -x -q -o faulthandler_timeout=200
