"""Developer tool: a few small searches covering every search kernel, for compute-sanitizer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy
import tree_search_planning_b200 as tsp
from tree_search_planning_b200.synthetic import random_state_dict, synthetic_inputs

for game, B, S, over in [("roundabout_env_ttc_flat", 70, 12, {}), ("roundabout_env_ttc_flat", 40, 12, {"players": [0, 1], "mask_absorbing_states": True}),
                         ("roundabout_env_ttc_flat_stochastic_concat", 50, 8, {}), ("roundabout_env_ttc_flat_risk_sensitive", 50, 8, {}),
                         ("cross_merge_env", 20, 6, {})]:
    for lpt in ("16", "8"):
        os.environ["TSP_FAST_LPT"] = lpt
        cfg = tsp.get_config(game, num_simulations=S, **over)
        eng = tsp.Engine(cfg, max_roots=B, max_simulations=2 * S, state_dict=random_state_dict(cfg, 0))
        obs, kw = synthetic_inputs(cfg, B, S, seed=1, ragged_legal=True)
        out = eng.run(S, observations=obs, trace=True, **kw)
        assert (out["visit_counts"].sum(1) == S).all()
        if tsp.variant_of(cfg) != 1:
            out = eng.run(S, resume=True, num_roots=B, **kw)
            assert (out["visit_counts"].sum(1) == 2 * S).all()
        eng.select_action(out["visit_counts"], 1.0, numpy.random.default_rng(0).random(B))
        eng.search_statistics(out["visit_counts"])
        print("ok", game, lpt, eng.stats()["search_kernel"], flush=True)
        eng.close()
