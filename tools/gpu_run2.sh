timeout 900 python -m pytest tests/ -q -m gpu -x -k "stochastic or risk or variant or chance or futures or golden or config" -o faulthandler_timeout=300 2>&1 | tail -3
for g in roundabout_env_ttc_flat_stochastic_concat roundabout_env_ttc_flat_risk_sensitive; do
  timeout 300 python tools/gpu_tune.py $g 16384 25 2>&1 | tail -1
  TSP_PHASE_TIMING=1 timeout 300 python tools/gpu_tune.py $g 16384 25 2>&1 | tail -2 | head -1 | cut -c1-200
done
