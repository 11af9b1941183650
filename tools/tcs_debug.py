"""Developer script: streamed tcgen05 kernel vs the oracle's float32 network, error statistics per record column."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy
import tree_search_planning_b200 as tsp
from tree_search_planning_b200.synthetic import random_state_dict, synthetic_inputs
from oracle.oracle import Oracle

game = sys.argv[1] if len(sys.argv) > 1 else "cross_merge_env"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
for S in [int(x) for x in (sys.argv[3:] or ["1", "4", "40", "200"])]:
    cfg = tsp.get_config(game, num_simulations=S)
    sd = random_state_dict(cfg, 14)
    eng = tsp.Engine(cfg, max_roots=B, max_simulations=S, state_dict=sd)
    orc = Oracle(cfg, sd)
    obs, kw = synthetic_inputs(cfg, B, S, 14)
    got = eng.run(S, observations=obs, trace=True, **kw)
    want = orc.run(S, observations=obs, trace=True, **kw)
    a = got["trace_records"].astype(numpy.float64); w = want["trace_records"].astype(numpy.float64)
    err = numpy.abs(a - w) / numpy.maximum(1.0, numpy.abs(w))
    err[:, :, 3] = 0  # terminal logit: evaluated by the oracle only
    same = (got["visit_counts"] == want["visit_counts"]).all(axis=1)
    tree_max = err.max(axis=(1, 2))
    # first record at which a tree exceeds 1e-5 in a prior column
    pri = err[:, :, 4:9].max(axis=2)
    first_bad = numpy.where((pri > 1e-5).any(axis=1), (pri > 1e-5).argmax(axis=1), -1)
    print("S=%d kernel=%d visit agreement %.4f; trees with max err >1e-5: %d, >2e-3: %d; col max (all trees) %s" % (
        S, eng.stats()["search_kernel"], same.mean(), (tree_max > 1e-5).sum(), (tree_max > 2e-3).sum(),
        numpy.array2string(err.max(axis=(0, 1))[:9], precision=2)))
    ok = tree_max <= 2e-3
    print("   non-diverged trees: col max %s; prior err quantiles (all records) 50%%=%.2e 99%%=%.2e 99.99%%=%.2e" % (
        numpy.array2string(err[ok].max(axis=(0, 1))[:9], precision=2), numpy.quantile(pri, 0.5), numpy.quantile(pri, 0.99), numpy.quantile(pri, 0.9999)))
    bad = numpy.argwhere((pri > 1e-5) & ok[:, None])
    print("   records with prior err > 1e-5 in non-diverged trees: %d; first few (tree, rec, err): %s" % (
        len(bad), [(int(t), int(r), float(pri[t, r])) for t, r in bad[:8]]))
    print("   first bad record index per tree: hist", numpy.bincount(first_bad[first_bad >= 0], minlength=1)[:20], "value col err max", err[ok][:, :, 1].max(), "reward", err[ok][:, :, 2].max())
    eng.close()
