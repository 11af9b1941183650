for nb in 0 1 0 1; do
TSP_NO_BALANCE=$nb timeout 300 python tools/gpu_tune.py roundabout_env_ttc_flat_stochastic_concat 16384 25 2>&1 | tail -1 | cut -c1-190
done
TSP_NO_BALANCE=1 timeout 300 python tools/gpu_tune.py roundabout_env_ttc_flat_risk_sensitive 16384 25 2>&1 | tail -1 | cut -c1-190
TSP_NO_BALANCE=0 timeout 300 python tools/gpu_tune.py roundabout_env_ttc_flat_risk_sensitive 16384 25 2>&1 | tail -1 | cut -c1-190
