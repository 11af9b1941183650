TSP_PHASE_TIMING=1 timeout 120 python tools/gpu_tune.py roundabout_env_ttc_flat_stochastic_concat 16384 25 | tail -2
TSP_PHASE_TIMING=1 timeout 120 python tools/gpu_tune.py roundabout_env_ttc_flat_risk_sensitive 16384 25 | tail -2
TSP_PHASE_TIMING=1 timeout 120 python tools/gpu_tune.py roundabout_env_ttc_flat 4096 50 | tail -2
