for i in 1 2; do
timeout 300 python tools/gpu_tune.py roundabout_env_ttc_flat 4096 50 2>&1 | tail -1 | cut -c1-175
timeout 300 python tools/gpu_tune.py roundabout_env_ttc_flat 65536 50 2>&1 | tail -1 | cut -c1-175
done
