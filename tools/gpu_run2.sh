for lib in "" "TSP_LIBRARY=tools/libtsp_prev.so"; do
  echo "== $lib"
  env $lib timeout 120 python tools/gpu_tune.py cross_merge_env 8192 200 | tail -1 | cut -c1-200
  env $lib timeout 120 python tools/gpu_tune.py cross_merge_env 65536 200 | tail -1 | cut -c1-200
  env $lib timeout 120 python tools/gpu_tune.py highway_env_ttc_flat 65536 50 | tail -1 | cut -c1-200
  env $lib timeout 120 python tools/gpu_tune.py roundabout_env_ttc_flat 4096 50 | tail -1 | cut -c1-200
done
