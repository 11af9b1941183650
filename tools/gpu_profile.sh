# round 2 profiling run (one GPU): every ncu command follows the same command line having exited 0 without ncu
set -x
O=gpurun_out
python bench.py --steps 200 --warmup 5 > $O/r2_bench_1gpu.json 2> $O/r2_bench_1gpu.err
python bench.py --impl reference --steps 5 --warmup 1 > $O/r2_bench_reference_arm.json 2> $O/r2_ref.err
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-configs > $O/r2_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/r2_launches_bench_steps4_warmup3.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-configs > $O/r2_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:search_tc_kernel -s 4 -c 1 -o $O/r2_search_tc python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-configs > $O/r2_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:root_tcs -s 4 -c 1 -o $O/r2_root_tcs python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-configs > $O/r2_ncu3.log 2>&1
# the streamed kernel: cross-merge 8,192 roots x 200 simulations
python tools/gpu_tune.py cross_merge_env 8192 200 > $O/r2_tcs_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:search_tcs -s 3 -c 1 -o $O/r2_search_tcs python tools/gpu_tune.py cross_merge_env 8192 200 > $O/r2_ncu4.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/r2_launches_cross_merge_8192x200.csv python tools/gpu_tune.py cross_merge_env 8192 200 > $O/r2_ncu5.log 2>&1
# the chance-node kernel (both futures per pass): stochastic search, 16,384 roots x 25 simulations
python tools/gpu_tune.py roundabout_env_ttc_flat_stochastic_concat 16384 25 > $O/r2_variant_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:search_tc_variant -s 3 -c 1 -o $O/r2_search_tc_variant python tools/gpu_tune.py roundabout_env_ttc_flat_stochastic_concat 16384 25 > $O/r2_ncu6.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke.txt 2>&1
for f in r2_search_tc r2_root_tcs r2_search_tcs r2_search_tc_variant; do
  ncu -i $O/$f.ncu-rep --page raw --csv > $O/$f.raw.csv 2>/dev/null
  python tools/ncu_summary.py $O/$f.raw.csv > $O/${f}_ncu_full_summary.json
  ncu -i $O/$f.ncu-rep --page source --csv --print-source sass > $O/$f.src.csv 2>/dev/null
  python tools/ncu_regions.py $O/$f.src.csv 100 > $O/${f}_stall_regions.txt 2>&1
  rm -f $O/$f.ncu-rep $O/$f.raw.csv   # (gpurun_out travels back only below 64 MiB; the summaries are what is kept)
  gzip -f $O/$f.src.csv
done
timeout 900 python -m pytest tests/ -q -m gpu -s -o faulthandler_timeout=300 2>&1 | grep -v "^$" > $O/r2_gpu_tests.log; tail -2 $O/r2_gpu_tests.log
