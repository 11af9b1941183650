"""Developer script: search-kernel time for a config under the current TSP_* env overrides."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
import tree_search_planning_b200 as tsp
from tree_search_planning_b200.synthetic import random_state_dict, synthetic_inputs
game, B, S = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
cfg = tsp.get_config(game, num_simulations=S)
eng = tsp.Engine(cfg, max_roots=B, max_simulations=S, state_dict=random_state_dict(cfg, 0))
obs, kw = synthetic_inputs(cfg, B, S, seed=1)
for _ in range(3):
    out = eng.run(S, observations=obs, **kw)
eng.reset_timing()
eng.phase_cycles()
for _ in range(10):
    out = eng.run(S, observations=obs, **kw)
st = eng.stats()
ms = st["sum_search_ms"] / st["timed_runs"]
if os.environ.get("TSP_PHASE_TIMING"):
    print("phase cycles per iteration (thread 0 of each CTA):", eng.phase_cycles())
print(json.dumps(dict(game=game, B=B, S=S, tm=st["trees_per_cta"], thr=st["threads_per_cta"], smem=st["search_smem_bytes"],
                      ctas=st["search_ctas"], search_ms=round(ms, 4), root_ms=round(st["sum_root_ms"] / st["timed_runs"], 4),
                      sims_per_s=round(B * S / ms * 1e3), us_per_iter=round(ms * 1e3 / S, 2),
                      env={k: v for k, v in os.environ.items() if k.startswith("TSP_")})))
