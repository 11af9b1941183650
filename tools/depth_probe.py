"""Developer probe: how unevenly is the selection depth spread over the CTAs (32 consecutive trees each) of config 2?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy
import tree_search_planning_b200 as tsp
from tree_search_planning_b200.synthetic import random_state_dict, synthetic_inputs
S, B = 50, 4096
cfg = tsp.get_config("roundabout_env_ttc_flat", num_simulations=S)
eng = tsp.Engine(cfg, max_roots=B, max_simulations=S, state_dict=random_state_dict(cfg, 0))
obs, kw = synthetic_inputs(cfg, B, S, seed=1)
out = eng.run(S, observations=obs, want=("visit_counts", "total_select_depth", "max_tree_depth"), **kw)
d = out["total_select_depth"].astype(numpy.float64) / S
print("mean depth per tree: mean %.3f min %.3f max %.3f std %.3f" % (d.mean(), d.min(), d.max(), d.std()))
for trees in (16, 32):
    g = d.reshape(-1, trees)
    print("%d-tree blocks: mean of block means %.3f, min %.3f, max %.3f (max / mean %.3f); block max-of-tree-means: mean %.3f max %.3f" % (
        trees, g.mean(1).mean(), g.mean(1).min(), g.mean(1).max(), g.mean(1).max() / g.mean(1).mean(), g.max(1).mean(), g.max(1).max()))
