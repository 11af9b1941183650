"""Developer script (not a test): first contact with the GPU, prints details instead of asserting."""
import sys, time, os
import numpy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import tree_search_planning_b200 as tsp
from golden_util import Golden, golden_names
from oracle.oracle import Oracle

for name in golden_names():
    g = Golden(name)
    eng = tsp.Engine(g.config, max_roots=64, max_simulations=g.S, state_dict=g.state_dict)
    got = eng.run(g.S, fed_root=g.out["root_record"], fed_records=g.out["records"], **g.run_kwargs())
    ok = {k: bool((numpy.asarray(got[k]) == numpy.asarray(g.out[k])).all()) for k in
          ("visit_counts", "root_value", "child_value", "child_prior", "child_reward", "max_tree_depth")}
    print("FED", name, ok, flush=True)
    if not ok["visit_counts"]:
        print("   got", got["visit_counts"][:3].tolist(), "want", g.out["visit_counts"][:3].tolist())
    got = eng.run(g.S, observations=g.obs, trace=True, **g.run_kwargs())
    same = (got["visit_counts"] == g.out["visit_counts"]).all(axis=1)
    print("E2E", name, "visit agreement", same.mean(), "max |root_value diff|",
          numpy.abs(got["root_value"] - g.out["root_value"]).max(), "root rec diff",
          numpy.abs(got["trace_root"] - g.out["root_record"]).max(), flush=True)
    print("    stats", eng.stats(), flush=True)
    eng.close()
