timeout 900 python -m pytest tests/test_tc.py -q -m gpu -x -k "other_layer_widths" -o faulthandler_timeout=300 2>&1 | tail -15
