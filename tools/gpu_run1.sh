python bench.py --steps 300 --warmup 10 --no-cpu-baseline > gpurun_out/t13_bench_overlap.json 2> gpurun_out/t13_bench_overlap.err
TSP_NO_OVERLAP=1 python bench.py --steps 300 --warmup 10 --no-cpu-baseline > gpurun_out/t13_bench_nooverlap.json 2>> gpurun_out/t13_bench_overlap.err
( time timeout 600 python -m pytest tests/ -m gpu -x -q -o faulthandler_timeout=200 ) > gpurun_out/t13_pytest_full.log 2>&1
