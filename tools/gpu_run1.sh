( time timeout 600 python -m pytest tests/ -m gpu -x -q -o faulthandler_timeout=200 ) > gpurun_out/t14_pytest_full.log 2>&1
timeout 120 python tests/gpu_tune.py roundabout_env_ttc_flat 4096 50 > gpurun_out/t14_tune.log 2>&1
timeout 120 python tests/gpu_tune.py highway_env_ttc_flat 65536 50 >> gpurun_out/t14_tune.log 2>&1
timeout 120 python tests/gpu_tune.py highway_env_ttc_flat 8192 200 >> gpurun_out/t14_tune.log 2>&1
