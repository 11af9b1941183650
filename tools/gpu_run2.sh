for pf in 0 1; do
for cfg in "roundabout_env_ttc_flat 65536 50" "highway_env_ttc_flat 8192 200" "highway_env_ttc_flat 32768 200" "roundabout_env_ttc_flat 4096 50" "roundabout_env_ttc_flat_stochastic_concat 16384 25"; do
TSP_PREFETCH_L2=$pf timeout 300 python tools/gpu_tune.py $cfg 2>&1 | tail -1 | cut -c1-175
done
done
