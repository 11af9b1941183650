timeout 900 python -m pytest tests/test_tc.py -q -m gpu -x -k "other_layer_widths or root_kernel_shapes or envelope" -o faulthandler_timeout=300 2>&1 | grep -v "^$" | tail -25
