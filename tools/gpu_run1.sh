python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/f4_pytest.log
for g in 1 2; do TSP_FAST_GROUPS=$g python tests/gpu_tune.py roundabout_env_ttc_flat 4096 50; done > gpurun_out/f4_tune.log 2>&1
TSP_PHASE_TIMING=1 TSP_FAST_GROUPS=2 python tests/gpu_tune.py roundabout_env_ttc_flat 4096 50 >> gpurun_out/f4_tune.log 2>&1
for g in 1 2; do TSP_FAST_GROUPS=$g python tests/gpu_tune.py highway_env_ttc_flat 65536 50; done >> gpurun_out/f4_tune.log 2>&1
TSP_FAST_GROUPS=2 python tests/gpu_tune.py highway_env_ttc_flat 8192 200 >> gpurun_out/f4_tune.log 2>&1
