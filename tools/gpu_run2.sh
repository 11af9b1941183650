TSP_PHASE_TIMING=1 timeout 120 python tools/gpu_tune.py cross_merge_env 8192 200 2>&1 | tail -2
timeout 120 python tools/gpu_tune.py cross_merge_env 8192 200
timeout 120 python tools/gpu_tune.py cross_merge_env 65536 200
timeout 600 python -m pytest tests/test_tree_export.py tests/test_gpu_parity.py tests/test_tc.py -x -q -m gpu -k "cross_merge or expansion or two_players or envelope or config5 or network_parity" -o faulthandler_timeout=200 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
