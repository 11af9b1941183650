for env in "" "TSP_NO_BALANCE=1"; do
  echo "== $env"
  env $env timeout 120 python tools/gpu_tune.py roundabout_env_ttc_flat_stochastic_concat 16384 25 | tail -1
  env $env timeout 120 python tools/gpu_tune.py roundabout_env_ttc_flat_risk_sensitive 16384 25 | tail -1
  env $env timeout 120 python tools/gpu_tune.py highway_env_ttc_flat 32768 200 | tail -1
  env $env timeout 120 python tools/gpu_tune.py highway_env_ttc_flat 16384 200 | tail -1
done
timeout 900 python -m pytest tests/ -q -m gpu -x -o faulthandler_timeout=300 2>&1 | tail -3
