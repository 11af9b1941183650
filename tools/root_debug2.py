import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy
import tree_search_planning_b200 as tsp
from tree_search_planning_b200.synthetic import random_state_dict, synthetic_inputs
from oracle.oracle import Oracle
for name, over in [("orig", {}), ("rep64x3", dict(fc_representation_layers=[64, 64, 64])), ("rep128x3", dict(fc_representation_layers=[128, 128, 128])),
                   ("rep64", dict(fc_representation_layers=[64])), ("H128", dict(encoding_size=128)), ("rep32", dict(fc_representation_layers=[32])),
                   ("H128rep128", dict(encoding_size=128, fc_representation_layers=[128])), ("H256", dict(encoding_size=256, fc_representation_layers=[64]))]:
    cfg = tsp.get_config("highway_env_ttc_flat", **over)
    B = 256
    sd = random_state_dict(cfg, 7)
    eng = tsp.Engine(cfg, max_roots=B, max_simulations=5, state_dict=sd)
    obs, kw = synthetic_inputs(cfg, B, 5, 7)
    got = eng.initial_inference(obs)
    want = Oracle(cfg, sd).initial_inference(obs)
    eh = numpy.abs(got["hidden"] - want["hidden"])
    pos = numpy.arange(B) % 32
    by = numpy.array([eh[pos == i].max() for i in range(32)])
    print("%-12s root_kernel %d: max err rows 0-7 %.2e | 8-15 %.2e | 16-23 %.2e | 24-31 %.2e ; policy %.2e" % (
        name, eng.stats()["root_kernel"], by[:8].max(), by[8:16].max(), by[16:24].max(), by[24:].max(),
        numpy.abs(got["policy_logits"] - want["policy_logits"]).max()))
    eng.close()
