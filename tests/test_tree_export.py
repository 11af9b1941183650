"""The whole search tree below the root, as the reference's consumers walk it (diagnose_model.py:156-185 recurses through
node.children): tsp_export_tree + the lazy Node views of mcts.py against the Node graphs of the UNMODIFIED reference
(tests/golden/trees.npz, written by oracle/make_golden_tree.py; depth-first, children in insertion order).

GPU: the fixture's recorded network outputs are fed to the engine, so every node must match bit for bit.
CPU: the view logic alone, on a hand-built block array."""
import os

import numpy
import pytest

from golden_util import Golden, GOLDEN_DIR
import tree_search_planning_b200 as tsp
from tree_search_planning_b200 import mcts as tmcts
from tree_search_planning_b200.configs import variant_of
from tree_search_planning_b200.engine import TreeInfo, tree_block_dtype

TREES = numpy.load(os.path.join(GOLDEN_DIR, "trees.npz"))


def flatten_view(root):
    rows = []

    def walk(node, depth, key):
        v = node.value
        rows.append((depth, key, node.visit_count, node.prior, node.reward, float(v() if callable(v) else v), int(node.expanded())))
        for k, ch in node.children.items():
            walk(ch, depth + 1, int(k))

    walk(root, 0, -1)
    return rows


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted({k.split(":")[0] for k in TREES.files}))
def test_exported_tree_equals_the_reference_node_graph(name):
    g = Golden(name)
    eng = tsp.Engine(g.config, max_roots=g.N, max_simulations=g.S, state_dict=g.state_dict)
    res = eng.run(g.S, fed_root=g.out["root_record"], fed_records=g.out["records"], **g.run_kwargs())
    legal = [[int(a) for a in g.legal[i][:g.n_legal[i]]] for i in range(g.N)]
    batch = tmcts.BatchResult(res, legal, variant_of(g.config), engine=eng)
    for key in sorted(k for k in TREES.files if k.split(":")[0] == name):
        i = int(key.split(":")[1])
        want = TREES[key]
        got = flatten_view(batch.root(i))
        assert len(got) == len(want), (key, len(got), len(want))
        for n, (r, w) in enumerate(zip(got, want)):
            assert r[0] == w["depth"] and r[1] == w["key"] and r[2] == w["visit_count"] and r[6] == w["expanded"], (key, n, r, w)
            if n > 0:  # (the root's own prior / reward are never read: Node(0) in self_play.py:335)
                assert numpy.float64(r[3]) == w["prior"] and numpy.float64(r[4]) == w["reward"], (key, n, r, w)
            assert numpy.float64(r[5]) == w["value"] or (r[5] == 0 and w["value"] == 0), (key, n, r, w)
    # a later search on the engine invalidates the device tree: deeper levels then read as unexpanded, depth 1 stays
    # (a tree exported BEFORE the next search is a host copy and stays readable: root 0 was walked above)
    last = g.N - 1
    root, root0 = batch.root(last), batch.root(0)
    eng.run(g.S, fed_root=g.out["root_record"], fed_records=g.out["records"], **g.run_kwargs())
    assert [c.visit_count for c in root.children.values()] == [int(g.out["visit_counts"][last][a]) for a in legal[last]]
    assert all(c.children == {} for c in root.children.values())
    assert any(len(c.children) > 0 for c in root0.children.values())
    eng.close()


@pytest.mark.gpu
def test_exported_tree_carries_hidden_states():
    cfg = tsp.get_config("highway_env_ttc_flat")
    from tree_search_planning_b200.synthetic import random_state_dict, synthetic_inputs
    sd = random_state_dict(cfg, 3)
    eng = tsp.Engine(cfg, max_roots=8, max_simulations=cfg.num_simulations, state_dict=sd)
    obs, kw = synthetic_inputs(cfg, 8, cfg.num_simulations, seed=4)
    res = eng.run(cfg.num_simulations, observations=obs, want_root_hidden=True, **kw)
    batch = tmcts.BatchResult(res, [list(cfg.action_space)] * 8, 0, engine=eng)
    root = batch.root(2)
    child = max(root.children.values(), key=lambda c: c.visit_count)
    assert child.expanded() and len(child.children) == len(cfg.action_space)
    tree = batch.device_tree(2)
    assert numpy.array_equal(tree.hidden[0], res["root_hidden"][2])  # block 0 == the root
    # the hidden state of an expanded node is what the network returns for (parent state, action) (models.py:233-257)
    a = [k for k, c in root.children.items() if c is child][0]
    want = eng.recurrent_inference(res["root_hidden"][2][None], numpy.array([a], numpy.int32))["hidden"]
    grand = child.children  # noqa: F841  (forces the lazy build)
    hs = tree.hidden[int(tree.blocks[0]["child"][list(root.children).index(a)])]
    assert numpy.abs(hs - want[0]).max() <= 1e-5
    eng.close()


def test_device_tree_views_on_a_hand_built_tree():
    """CPU: block arrays -> Node views (deterministic layout: block id == hidden slot)."""
    dt = tree_block_dtype(5)
    assert dt.itemsize == 128
    blocks = numpy.zeros(2, dt)
    blocks["child"][:] = -1
    # root: two legal actions (3, 1); action 3 expanded into block 1 and visited twice
    blocks[0]["visit"][:2] = (2, 0)
    blocks[0]["child"][0] = 1
    blocks[0]["reward"][0] = 0.5
    blocks[0]["value_sum"][0] = 1.5
    blocks[1]["visit"][:5] = (1, 0, 0, 0, 0)
    blocks[1]["prior"][:5] = (0.1, 0.2, 0.3, 0.25, 0.15)
    blocks[1]["value_sum"][0] = 0.25
    blocks[1]["parent"] = 0
    blocks[1]["aux"] = 0 | (1 << 8)
    info = TreeInfo()
    info.n_blocks, info.block_bytes, info.num_children, info.n_root, info.root_visit, info.n_hidden = 2, 128, 5, 2, 2, 2
    info.root_prior[0], info.root_prior[1] = 0.75, 0.25
    info.root_action[0], info.root_action[1] = 3, 1
    hidden = numpy.arange(2 * 4, dtype=numpy.float32).reshape(2, 4)
    tree = tmcts.DeviceTree(dict(info=info, blocks=blocks, hidden=hidden), 0, 5, 1)
    res = dict(visit_counts=numpy.array([[0, 0, 0, 2, 0]]), root_value=numpy.array([0.7]),
               child_prior=numpy.array([[0, 0.25, 0, 0.75, 0]]), child_reward=numpy.array([[0, 0, 0, 0.5, 0]]),
               child_value=numpy.array([[0, 0, 0, 0.75, 0]]))
    root = tmcts.RootView(res, 0, [3, 1], 0, tree=lambda: tree)
    assert list(root.children) == [3, 1]
    c3, c1 = root.children[3], root.children[1]
    assert c3.expanded() and not c1.expanded() and c1.children == {}
    assert list(c3.children) == [0, 1, 2, 3, 4]
    g = c3.children[0]
    assert g.visit_count == 1 and g.value() == 0.25 and abs(g.prior - 0.1) < 1e-7 and not g.expanded() and g.to_play == -1
    # through the tree alone the root's children carry their hidden states
    top = tree.children_of(0, 0)
    assert numpy.array_equal(top[3].hidden_state, hidden[1][None]) and top[1].hidden_state is None
    assert top[3].prior == 0.75 and top[3].value() == 0.75


@pytest.mark.gpu
@pytest.mark.parametrize("game,B,S,kernel", [("cross_merge_env", 200, 120, 3), ("highway_env_ttc_flat", 4096, 50, 2),
                                             ("highway_env_ttc_flat", 20000, 30, 2)])
def test_every_expansion_of_the_search_kernel_against_the_oracle_network(game, B, S, kernel):
    """Path-independent check of the network INSIDE the search kernels (tcgen05 with weights in tensor memory / streamed
    by TMA): for every node the search expanded, the stored hidden state, reward and child priors must be what the
    oracle's float32 network (models.py:233-257, 283-287) returns for (the parent's stored hidden state, the action) --
    whatever order the two searches evaluated their leaves in. Hidden states and priors within 1e-5, the reward within
    the quantum of the reference's float32 inverse transform."""
    from oracle.oracle import Oracle
    from tree_search_planning_b200.synthetic import random_state_dict, synthetic_inputs
    cfg = tsp.get_config(game, num_simulations=S)
    sd = random_state_dict(cfg, 21)
    eng = tsp.Engine(cfg, max_roots=B, max_simulations=S, state_dict=sd)
    obs, kw = synthetic_inputs(cfg, B, S, seed=22)
    eng.run(S, observations=obs, **kw)
    assert eng.stats()["search_kernel"] == kernel
    A = len(cfg.action_space)
    ph, act, ch, rew, pri = [], [], [], [], []
    rng = numpy.random.default_rng(5)
    for i in sorted(rng.choice(B, size=12, replace=False)):
        t = eng.export_tree(int(i))
        blocks, hidden, info = t["blocks"], t["hidden"], t["info"]
        assert info.n_blocks == S + 1 and len(hidden) == S + 1  # one expansion per simulation (no terminal head)
        for k in range(1, info.n_blocks):
            parent, slot = int(blocks[k]["parent"]), int(blocks[k]["aux"]) & 0xFF
            assert int(blocks[parent]["child"][slot]) == k
            ph.append(hidden[parent]); ch.append(hidden[k])
            act.append(int(info.root_action[slot]) if parent == 0 else slot)
            rew.append(float(blocks[parent]["reward"][slot])); pri.append(blocks[k]["prior"][:A].copy())
    want = Oracle(cfg, sd).recurrent_inference(numpy.stack(ph), numpy.array(act, numpy.int32))
    lg = want["policy_logits"].astype(numpy.float32)
    e = numpy.exp(lg - lg.max(axis=1, keepdims=True))
    want_prior = e / e.sum(axis=1, keepdims=True)
    err_h = numpy.abs(numpy.stack(ch) - want["hidden"]).max()
    err_p = numpy.abs(numpy.stack(pri) - want_prior).max()
    err_r = (numpy.abs(numpy.array(rew) - want["reward"]) / (1 + numpy.abs(want["reward"]))).max()
    print("EXPANSIONS %s B=%d S=%d: %d nodes, max |hidden err| %.2e, max |prior err| %.2e, max reward err %.2e" % (
        game, B, S, len(ph), err_h, err_p, err_r))
    assert err_h <= 1e-5 and err_p <= 1e-5 and err_r <= 2e-4
    eng.close()
