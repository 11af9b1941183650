TSP_FAST_LPT=8 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/f21_pytest.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 >> gpurun_out/f21_pytest.log
for l in 16 8; do TSP_FAST_LPT=$l python tests/gpu_tune.py roundabout_env_ttc_flat 4096 50; done > gpurun_out/f21_tune.log 2>&1
for l in 16 8; do TSP_FAST_LPT=$l python tests/gpu_tune.py highway_env_ttc_flat 65536 50; done >> gpurun_out/f21_tune.log 2>&1
for l in 16 8; do TSP_FAST_LPT=$l python tests/gpu_tune.py highway_env_ttc_flat 8192 200; done >> gpurun_out/f21_tune.log 2>&1
for l in 16 8; do TSP_FAST_LPT=$l python tests/gpu_tune.py highway_env_ttc_flat 16384 50; done >> gpurun_out/f21_tune.log 2>&1
TSP_FAST_LPT=8 TSP_FAST_GROUPS=2 python tests/gpu_tune.py highway_env_ttc_flat 65536 50 >> gpurun_out/f21_tune.log 2>&1
