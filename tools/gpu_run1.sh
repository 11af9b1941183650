for n in 8 16 32 64; do timeout 60 ./tools/tc_stage_test $n | grep -E "MMA N|iters=200" | cut -c1-30,150-300; done > gpurun_out/tc5.log 2>&1
TSP_PHASE_TIMING=1 timeout 120 python tests/gpu_tune.py roundabout_env_ttc_flat 4096 50 > gpurun_out/t5_tune.log 2>&1
timeout 120 python tests/gpu_tune.py roundabout_env_ttc_flat 4096 50 >> gpurun_out/t5_tune.log 2>&1
TSP_TC_NOFOLD=1 timeout 120 python tests/gpu_tune.py roundabout_env_ttc_flat 4096 50 >> gpurun_out/t5_tune.log 2>&1
timeout 120 python tests/gpu_tune.py highway_env_ttc_flat 65536 50 >> gpurun_out/t5_tune.log 2>&1
timeout 900 python -m pytest tests/test_tc.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/t5_pytest.log
