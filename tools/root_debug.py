"""Developer script: root network (tsp_initial_inference) vs the oracle, error statistics; TSP_NO_RTCS=1 for the mma.sync kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy
import tree_search_planning_b200 as tsp
from tree_search_planning_b200.synthetic import random_state_dict, synthetic_inputs
from oracle.oracle import Oracle

for game in ("highway_env_ttc_flat", "cross_merge_env"):
    cfg = tsp.get_config(game)
    B = 2000
    sd = random_state_dict(cfg, 7)
    eng = tsp.Engine(cfg, max_roots=B, max_simulations=5, state_dict=sd)
    obs, kw = synthetic_inputs(cfg, B, 5, 7)
    got = eng.initial_inference(obs)
    want = Oracle(cfg, sd).initial_inference(obs)
    st = eng.stats()
    eh = numpy.abs(got["hidden"] - want["hidden"])
    print(game, "root_kernel", st.get("root_kernel"), "hidden max err %.3g  mean %.3g  rows with err>1e-5: %d; policy logits max err %.3g; value logits %.3g" % (
        eh.max(), eh.mean(), (eh.max(axis=1) > 1e-5).sum(), numpy.abs(got["policy_logits"] - want["policy_logits"]).max(),
        numpy.abs(got["value_logits"] - want["value_logits"]).max()))
    worst = numpy.argsort(-eh.max(axis=1))[:5]
    print("   worst rows", worst.tolist(), "errs", eh.max(axis=1)[worst])
    # error by row position inside the CTA (32 rows per CTA)
    pos = numpy.arange(B) % 32
    print("   mean err by row-in-CTA:", numpy.array2string(numpy.array([eh[pos == i].mean() for i in range(32)]), precision=2))
    eng.close()
