for f in 0 1 2 4 5; do TSP_DBG_FLAGS=$f TSP_PHASE_TIMING=1 python tests/gpu_tune.py roundabout_env_ttc_flat 4096 50; done > gpurun_out/f11_tune.log 2>&1
