python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/f22_pytest.log
for g in roundabout_env_ttc_flat_stochastic_concat roundabout_env_ttc_flat_risk_sensitive; do
  python tests/gpu_tune.py $g 16384 25; TSP_NO_FAST=1 python tests/gpu_tune.py $g 16384 25; TSP_FAST_LPT=16 python tests/gpu_tune.py $g 16384 25
done > gpurun_out/f22_tune.log 2>&1
