set -x
python bench.py --sweep > gpurun_out/r1d_bench_1gpu.json 2> gpurun_out/r1d_bench_1gpu.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r1d_bench_reference_arm.json 2> gpurun_out/r1d_ref.err
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r1d_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r1d_launches_bench_steps4_warmup3.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r1d_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:search_tc -s 4 -c 2 -o gpurun_out/r1d_search_tc python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r1d_ncu2.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1d_smoke.log 2>&1
