timeout 1200 python -m pytest tests/test_tc.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/t6_pytest.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stochastic or risk or golden or config3 or config4 or ragged or continued or mcts_run" 2>&1 | tail -15 >> gpurun_out/t6_pytest.log
for g in roundabout_env_ttc_flat_stochastic_concat roundabout_env_ttc_flat_risk_sensitive; do for tc in 0 1; do TSP_NO_TC=$tc timeout 120 python tests/gpu_tune.py $g 16384 25; done; done > gpurun_out/t6_tune.log 2>&1
