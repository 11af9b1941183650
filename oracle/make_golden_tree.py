"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/trees.npz from the LIVE reference.

Run in the build container (needs /root/reference):  python -m oracle.make_golden_tree

For a few roots of existing fixtures (tests/golden/<name>.npz: same weights, observations and random streams) the
UNMODIFIED reference `MCTS.run` is run again and its whole Node graph -- not only the root's children -- is flattened in
depth-first order, children in insertion order (the order diagnose_model.py:156-185 walks them). The GPU test feeds the
fixture's recorded network outputs to the engine and compares the tree exported by tsp_export_tree node by node.

Per node: depth, key (action, or future z under a TransitionNode), visit_count, prior, reward, value (value_sum /
visit_count, or the `.value` attribute of the risk-sensitive nodes), expanded (has children).
"""
import os
import sys

import numpy

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))

from oracle import reference_live as RL  # noqa: E402
from golden_util import Golden  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "trees.npz")
CASES = [("highway_s25", 4), ("roundabout_s50_legal", 3), ("roundabout_s25_mask", 3), ("stochastic_s25", 3), ("risk_s25", 3),
         ("roundabout_s25_twoplayer", 2)]
NODE_DTYPE = numpy.dtype([("depth", numpy.int32), ("key", numpy.int32), ("visit_count", numpy.int32), ("prior", numpy.float64),
                          ("reward", numpy.float64), ("value", numpy.float64), ("expanded", numpy.int8)])


def node_value(n):
    v = n.value
    return float(v() if callable(v) else v)


def flatten(root):
    rows = []

    def walk(node, depth, key):
        rows.append((depth, key, int(node.visit_count), float(node.prior), float(node.reward), node_value(node),
                     1 if len(node.children) > 0 else 0))
        for k, ch in node.children.items():
            walk(ch, depth + 1, int(k))

    walk(root, 0, -1)
    return numpy.array(rows, dtype=NODE_DTYPE)


def main():
    ref = RL.load_reference()
    out = {}
    for name, n_roots in CASES:
        g = Golden(name)
        model = ref.models.MuZeroNetwork(g.config)
        model.load_state_dict({k: ref.torch.from_numpy(numpy.asarray(v)) for k, v in g.state_dict.items()}, strict=False)
        model.eval()
        kw = g.run_kwargs()
        for i in range(n_roots):
            la = [int(a) for a in g.legal[i][:g.n_legal[i]]]
            r = RL.run_reference(g.config, model, g.obs[i].reshape(g.config.stacked_observations * 2 + 1, 1, -1), legal_actions=la,
                                 noise=kw["noise"][i][:len(la)], tie=kw["tie"][i], chance=kw["chance"][i],
                                 add_exploration_noise=kw["add_exploration_noise"], uniform_policy=kw["uniform_policy"],
                                 to_play=int(g.to_play[i]) if g.to_play is not None else 0, keep_root=True)
            # the fixture and this run are the same search
            assert (r["visit_counts"] == g.out["visit_counts"][i]).all(), (name, i)
            assert r["root_value"] == g.out["root_value"][i], (name, i)
            out["%s:%d" % (name, i)] = flatten(r["root"])
            print(name, i, len(out["%s:%d" % (name, i)]), "nodes")
    numpy.savez_compressed(OUT, **out)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
