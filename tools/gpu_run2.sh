timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "continued_search_large" -o faulthandler_timeout=300 2>&1 | tail -15
