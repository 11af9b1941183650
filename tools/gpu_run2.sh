timeout 1200 python -m pytest tests/ -q -m gpu -o faulthandler_timeout=300 2>&1 | tail -8
for i in 1 2; do
timeout 300 python tools/gpu_tune.py roundabout_env_ttc_flat 4096 50 2>&1 | tail -1
done
timeout 300 python tools/gpu_tune.py roundabout_env_ttc_flat 65536 50 2>&1 | tail -1
timeout 300 python tools/gpu_tune.py roundabout_env_ttc_flat_stochastic_concat 16384 25 2>&1 | tail -1
timeout 300 python tools/gpu_tune.py roundabout_env_ttc_flat_risk_sensitive 16384 25 2>&1 | tail -1
timeout 300 python tools/gpu_tune.py cross_merge_env 8192 200 2>&1 | tail -1
