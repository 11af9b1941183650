timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config2" 2>&1 | tail -5 > gpurun_out/t3_pytest.log
TSP_PHASE_TIMING=1 timeout 120 python tests/gpu_tune.py roundabout_env_ttc_flat 4096 50 > gpurun_out/t3_tune.log 2>&1
timeout 120 python tests/gpu_tune.py roundabout_env_ttc_flat 4096 50 >> gpurun_out/t3_tune.log 2>&1
TSP_PHASE_TIMING=1 TSP_FAST_LPT=8 timeout 120 python tests/gpu_tune.py roundabout_env_ttc_flat 8192 50 >> gpurun_out/t3_tune.log 2>&1
for tc in 0 1; do TSP_NO_TC=$tc timeout 120 python tests/gpu_tune.py highway_env_ttc_flat 65536 50; done >> gpurun_out/t3_tune.log 2>&1
