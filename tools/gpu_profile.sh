set -x
timeout 900 python -m pytest tests/test_tc.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r1c_pytest_tc.log
python bench.py --sweep > gpurun_out/r1c_bench_1gpu.json 2> gpurun_out/r1c_bench_1gpu.err
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r1c_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r1c_launches_bench_steps4_warmup3.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r1c_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:search_tc -s 4 -c 2 -o gpurun_out/r1c_search_tc python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r1c_ncu2.log 2>&1
for n in 8 16 32 64; do timeout 60 ./tools/tc_stage_test $n | grep -E "MMA N|iters=200" | cut -c1-30,150-300; done > gpurun_out/tc5.log 2>&1
