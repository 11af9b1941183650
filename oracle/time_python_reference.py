"""TEST INFRASTRUCTURE ONLY -- times the UNMODIFIED Python reference (MCTS.run of muzero-general/self_play.py and its
stochastic / risk-sensitive siblings, torch-CPU network) in THIS container, where /root/reference exists, and writes
profiles/python_reference_rate.json. bench.py copies that record into cpu_baseline.python_reference (the GPU box has no
/root/reference, so the figure cannot be measured there); it is a reported baseline, not a target.

    python oracle/time_python_reference.py
"""
import json
import os
import platform
import sys
import time

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import reference_live as live  # noqa: E402
from tree_search_planning_b200.configs import get_config, observation_dim  # noqa: E402


def rate(game, sims, roots, seed=0):
    cfg = get_config(game, num_simulations=sims)
    ref = live.load_reference()
    ref.torch.set_num_threads(1)
    model = live.build_model(cfg, seed)
    rng = numpy.random.default_rng(seed)
    A = len(cfg.action_space)
    obs = rng.random((roots, 1, 1, observation_dim(cfg)), dtype=numpy.float32)
    live.run_reference(cfg, model, obs[0], noise=rng.dirichlet([0.25] * A), record=False)  # warm-up
    t0 = time.perf_counter()
    for b in range(roots):
        out = live.run_reference(cfg, model, obs[b], noise=rng.dirichlet([0.25] * A), record=False)
        assert int(out["visit_counts"].sum()) == sims
    dt = time.perf_counter() - t0
    return roots * sims / dt, dt


def main():
    rows = {}
    for game, sims, roots in [("roundabout_env_ttc_flat", 50, 40), ("highway_env_ttc_flat", 200, 10),
                              ("roundabout_env_ttc_flat_stochastic_concat", 25, 40),
                              ("roundabout_env_ttc_flat_risk_sensitive", 25, 40), ("cross_merge_env", 200, 10)]:
        r, dt = rate(game, sims, roots)
        rows["%s_S%d" % (game, sims)] = {"sims_per_s": r, "roots": roots, "seconds": dt}
        print(game, sims, "%.1f sims/s (%d roots, %.1f s)" % (r, roots, dt), flush=True)
    cpu = ""
    try:
        cpu = [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:
        pass
    head = rows["roundabout_env_ttc_flat_S50"]
    rec = {"value": head["sims_per_s"], "unit": "sims/s", "cores": 1, "kind": "reference",
           "what": "unmodified muzero-general MCTS.run (Python + torch-CPU, one root at a time, one thread), config 2 shapes "
                   "(roundabout_env_ttc_flat, 50 simulations per root), random-init weights",
           "host": "build container (%s, %s, %d logical cores; NOT the GPU box)" % (platform.node(), cpu, os.cpu_count()),
           "sample": "%d roots x 50 sims, %.1f s" % (head["roots"], head["seconds"]),
           "other_configs": rows, "script": "oracle/time_python_reference.py"}
    out = os.path.join(ROOT, "profiles", "python_reference_rate.json")
    json.dump(rec, open(out, "w"), indent=1)
    print("wrote", out)


if __name__ == "__main__":
    main()
