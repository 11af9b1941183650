"""Boundary proof with the reference's OWN objects (CPU; runs only where /root/reference exists).

The drop-in claim of SURVEY 8b is that what `MCTS.run` returns can be handed to the reference's consumers unchanged, and
that the reference's own network object can be handed to `MCTS.run`. Both directions are exercised here without a GPU:
  * `mcts.BatchResult` / `RootView` are built from search statistics (those of the golden fixtures, i.e. of the
    unmodified reference) and passed to the imported `self_play.SelfPlay.select_action`
    (self_play.py:277-299) and `GameHistory.store_search_statistics` (self_play.py:570-585; risk-sensitive:
    `root.value` is an attribute, self_play_risk_sensitive.py:612) next to the reference's own Node graph of the same
    search: the two must agree bit for bit;
  * a real `models.MuZeroNetwork(config)` goes through `weights.extract_mlps` (the path `MCTS.run(model, ...)` takes) and
    the resulting network must reproduce `model.initial_inference` / `recurrent_inference`.
"""
import numpy
import pytest

from oracle import reference_live as rl
from oracle.oracle import Oracle
from golden_util import Golden
from tree_search_planning_b200 import mcts as tmcts
from tree_search_planning_b200.configs import get_config, variant_of
from tree_search_planning_b200.weights import extract_mlps

pytestmark = pytest.mark.skipif(not rl.reference_available(), reason="needs the reference tree at /root/reference")


def _views(g):
    """RootView per fixture root, built the way MCTS.run_batch builds them from the engine's output arrays."""
    legal = [[int(a) for a in g.legal[i][:g.n_legal[i]]] for i in range(g.N)]
    res = {k: numpy.asarray(v) for k, v in g.out.items()}
    res.setdefault("root_predicted_value", numpy.zeros(g.N))
    batch = tmcts.BatchResult(res, legal, variant_of(g.config))
    return batch, legal


@pytest.mark.parametrize("name", ["highway_s25", "roundabout_s50_legal", "stochastic_s25", "risk_s25", "roundabout_s25_twoplayer"])
def test_root_view_feeds_the_reference_consumers(name):
    g = Golden(name)
    ref = rl.load_reference()
    variant = variant_of(g.config)
    mod = (ref.self_play, ref.self_play_stochastic, ref.self_play_risk_sensitive)[variant]
    batch, legal = _views(g)
    action_space = list(g.config.action_space)
    for i in range(min(g.N, 8)):
        root = batch.root(i)
        # what the reference itself would hold: children in the insertion order of legal_actions
        assert list(root.children.keys()) == legal[i]
        visits = [int(g.out["visit_counts"][i][a]) for a in legal[i]]
        # --- select_action (static method of the reference's SelfPlay), temperature 0 and sampled ---
        assert mod.SelfPlay.select_action(root, 0) == legal[i][int(numpy.argmax(visits))]
        numpy.random.seed(123 + i)
        a_view = mod.SelfPlay.select_action(root, 1.0)
        numpy.random.seed(123 + i)
        p = numpy.array(visits, dtype="int32") ** 1.0
        a_want = numpy.random.choice(legal[i], p=p / sum(p))
        assert a_view == a_want
        # --- GameHistory.store_search_statistics ---
        hist = mod.GameHistory()
        hist.store_search_statistics(root, action_space)
        total = max(sum(visits), 1)
        want = [int(g.out["visit_counts"][i][a]) / total if a in legal[i] else 0 for a in action_space]
        assert hist.child_visits[-1] == want
        rv = hist.root_values[-1]
        assert numpy.float64(rv) == numpy.float64(g.out["root_value"][i])
        # extra_info keys as the callers read them (self_play.py:429-433 vs self_play_stochastic.py:396-399)
        info = batch.extra_info(i)
        assert ("n_env_interactions" in info) == (variant == 0)
        assert info["max_tree_depth"] == int(g.out["max_tree_depth"][i])


def test_root_view_matches_a_live_reference_root():
    """The same consumers on the reference's own Node graph and on the view of the same search."""
    g = Golden("highway_s25")
    ref = rl.load_reference()
    model = ref.models.MuZeroNetwork(g.config)
    # (the fixtures carry the weights of the networks the planner evaluates: the reconstruction head is unused by MCTS)
    model.load_state_dict({k: ref.torch.from_numpy(numpy.asarray(v)) for k, v in g.state_dict.items()}, strict=False)
    model.eval()
    batch, legal = _views(g)
    i = 0
    kw = g.run_kwargs()
    streams = rl.ExternalStreams(noise=kw["noise"][i][:len(legal[i])], tie=kw["tie"][i], chance=None)
    with rl.patched_rng(streams), ref.torch.no_grad():
        root, info = ref.self_play.MCTS(g.config).run(model, g.obs[i].reshape(g.config.stacked_observations * 2 + 1, 1, -1),
                                                      legal[i], 0, bool(kw["add_exploration_noise"]))
    view = batch.root(i)
    assert [c.visit_count for c in root.children.values()] == [c.visit_count for c in view.children.values()]
    assert numpy.float64(root.value()) == numpy.float64(view.value())
    for a in legal[i]:
        assert numpy.float64(root.children[a].prior) == numpy.float64(view.children[a].prior)
        assert numpy.float64(root.children[a].value()) == numpy.float64(view.children[a].value())
        assert numpy.float64(root.children[a].reward) == numpy.float64(view.children[a].reward)
    h1, h2 = ref.self_play.GameHistory(), ref.self_play.GameHistory()
    h1.store_search_statistics(root, g.config.action_space)
    h2.store_search_statistics(view, g.config.action_space)
    assert h1.child_visits == h2.child_visits and h1.root_values == h2.root_values
    assert info["max_tree_depth"] == batch.extra_info(i)["max_tree_depth"]


@pytest.mark.parametrize("game", ["highway_env_ttc_flat", "cross_merge_env", "roundabout_env_ttc_flat_stochastic_concat"])
def test_reference_model_object_through_weight_ingestion(game):
    """models.MuZeroNetwork(config).state_dict() -> weights.extract_mlps -> network == the model's own inference."""
    ref = rl.load_reference()
    cfg = get_config(game)
    model = rl.build_model(cfg, seed=5)
    sd = model.state_dict()  # torch tensors with the .module. infixes of DataParallel
    mlps = extract_mlps(sd)
    for name in ("representation", "dynamics", "reward", "policy", "value"):
        assert name in mlps and len(mlps[name]) >= 1
    orc = Oracle(cfg, sd)
    rng = numpy.random.default_rng(3)
    from tree_search_planning_b200.configs import observation_dim
    B = 16
    obs = rng.random((B, observation_dim(cfg)), dtype=numpy.float32)
    got = orc.initial_inference(obs)
    with ref.torch.no_grad():
        t_obs = ref.torch.from_numpy(obs).reshape(B, cfg.stacked_observations * 2 + 1, 1, -1)
        value, reward, _term, logits, _recon, hidden = model.initial_inference(t_obs)
        want_value = ref.models.support_to_scalar(value, cfg.support_size).numpy().reshape(-1)
    numpy.testing.assert_allclose(got["hidden"], hidden.numpy(), rtol=0, atol=1e-5)
    numpy.testing.assert_allclose(got["policy_logits"], logits.numpy(), rtol=0, atol=1e-5)
    tol = 2e-3 * (1 + numpy.abs(want_value))  # float32 inverse transform quantum (golden_util.assert_close_quantised)
    assert (numpy.abs(got["value"].reshape(-1) - want_value) <= tol).all()
