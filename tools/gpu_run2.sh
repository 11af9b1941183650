timeout 900 python -m pytest tests/ -q -m gpu -x -k "not stochastic and not risk and not multi" -o faulthandler_timeout=300 2>&1 | tail -5
for nn in 0 1; do
for cfg in "roundabout_env_ttc_flat 4096 50" "roundabout_env_ttc_flat 65536 50" "highway_env_ttc_flat 8192 200"; do
TSP_TC_NONORM=$nn timeout 300 python tools/gpu_tune.py $cfg 2>&1 | tail -1 | cut -c1-175
done
done
TSP_PHASE_TIMING=1 timeout 300 python tools/gpu_tune.py roundabout_env_ttc_flat 4096 50 2>&1 | tail -2 | head -1
