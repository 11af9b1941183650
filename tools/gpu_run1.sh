python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cross_merge or two_players or config5" 2>&1 | tail -15 > gpurun_out/f25_pytest.log
for l in 16 8; do TSP_FAST_LPT=$l python tests/gpu_tune.py cross_merge_env 8192 200; done > gpurun_out/f25_tune.log 2>&1
TSP_NO_WGLOBAL=1 python tests/gpu_tune.py cross_merge_env 8192 200 >> gpurun_out/f25_tune.log 2>&1
TSP_FAST_LPT=16 python tests/gpu_tune.py cross_merge_env 4096 50 >> gpurun_out/f25_tune.log 2>&1
TSP_NO_WGLOBAL=1 python tests/gpu_tune.py cross_merge_env 4096 50 >> gpurun_out/f25_tune.log 2>&1
