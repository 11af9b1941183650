TSP_PHASE_TIMING=1 timeout 120 python tests/gpu_tune.py roundabout_env_ttc_flat 4096 50 > gpurun_out/t18_tune.log 2>&1
timeout 120 python tests/gpu_tune.py roundabout_env_ttc_flat 4096 50 >> gpurun_out/t18_tune.log 2>&1
timeout 120 python tests/gpu_tune.py highway_env_ttc_flat 65536 50 >> gpurun_out/t18_tune.log 2>&1
timeout 120 python tests/gpu_tune.py roundabout_env_ttc_flat_stochastic_concat 16384 25 >> gpurun_out/t18_tune.log 2>&1
( time timeout 600 python -m pytest tests/ -m gpu -x -q -o faulthandler_timeout=200 ) > gpurun_out/t18_pytest_full.log 2>&1
