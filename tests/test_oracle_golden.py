"""CPU: the C oracle (oracle/tsp_oracle.c) against the golden fixtures produced by the LIVE
reference (oracle/make_golden.py). This is what pins the oracle.

* tree arithmetic (select / ucb / expand / noise / ties / chance / backup / min-max / risk
  update_value): BIT-EXACT when the reference's recorded network outputs are fed back;
* the oracle's own float32 network: within 1e-5 of torch (relative to max(1, |x|));
* end to end with its own network: visit counts identical on these fixtures.
"""
import numpy
import pytest

from golden_util import Golden, golden_names, assert_tree_outputs_bit_exact, assert_close_quantised
from oracle.oracle import Oracle

NAMES = golden_names()


def test_fixtures_present():
    assert len(NAMES) >= 16


@pytest.mark.parametrize("name", NAMES)
def test_tree_logic_bit_exact_with_fed_network_outputs(name):
    g = Golden(name)
    orc = Oracle(g.config)  # no weights: network outputs come from the reference's records
    got = orc.run(g.S, fed_root=g.out["root_record"], fed_records=g.out["records"], B=g.N, **g.oracle_kwargs())
    assert_tree_outputs_bit_exact(got, g.out, what=name)
    assert (got["n_nodes"] == g.out["n_nodes"]).all()
    assert (got["n_records"] == g.out["n_records"]).all()
    assert (got["visit_counts"].sum(1) == g.S * g.P).all()  # self_play_stochastic.py:333


@pytest.mark.parametrize("name", NAMES)
def test_end_to_end_own_network(name):
    g = Golden(name)
    orc = Oracle(g.config, g.state_dict)
    got = orc.run(g.S, observations=g.obs, trace=True, want_root_hidden=True, **g.oracle_kwargs())
    # visit counts / depth identical on every fixture root (values are float32-network dependent)
    assert (got["visit_counts"] == g.out["visit_counts"]).all()
    assert (got["max_tree_depth"] == g.out["max_tree_depth"]).all()
    assert_close_quantised(got["root_value"], g.out["root_value"], min_fraction=0.85, what="root_value")
    assert_close_quantised(got["child_value"], g.out["child_value"], min_fraction=0.85, what="child_value")
    numpy.testing.assert_allclose(got["child_prior"], g.out["child_prior"], rtol=1e-5, atol=1e-6)
    numpy.testing.assert_allclose(got["root_hidden"], g.out["root_hidden"], rtol=1e-5, atol=1e-6)
    assert_close_quantised(got["trace_root"], g.out["root_record"], what="root record")
    for b in range(g.N):
        n = g.out["n_records"][b]  # (continued fixtures: of the last search)
        tr, ref = (got["trace_records"], g.out["records"]) if g.P == 1 else (got["trace_records"][-1], g.out["records"][-1])
        assert_close_quantised(tr[b, :n], ref[b, :n], what="records")


def _close(a, b, tol=1e-5):
    a, b = numpy.asarray(a, numpy.float64), numpy.asarray(b, numpy.float64)
    return numpy.abs(a - b) <= tol * numpy.maximum(1.0, numpy.abs(b))


@pytest.mark.parametrize("name", ["highway_s25", "cross_merge_s200", "stochastic_s25", "stochastic_sep_s25", "stochastic_sep_s25_legal"])
def test_network_parity_batched(name):
    g = Golden(name)
    orc = Oracle(g.config, g.state_dict)
    M = g.net["init_hidden"].shape[0]
    ini = orc.initial_inference(g.obs[:M])
    assert _close(ini["hidden"], g.net["init_hidden"]).all()
    assert _close(ini["policy_logits"], g.net["init_policy_logits"]).all()
    assert _close(ini["value_logits"], g.net["init_value_logits"]).all()
    assert _close(ini["value"], g.net["init_value"], 2e-4).all()
    h, a = g.net["rec_in_hidden"], g.net["rec_in_action"]
    if "rec_hidden" in g.net:
        rec = orc.recurrent_inference(h, a)
        for k_o, k_g in (("hidden", "rec_hidden"), ("policy_logits", "rec_policy_logits"),
                         ("value_logits", "rec_value_logits"), ("reward_logits", "rec_reward_logits"),
                         ("terminal", "rec_terminal")):
            assert _close(rec[k_o], g.net[k_g]).all(), k_o
        # the reference's float32 inverse transform quantises its output
        assert _close(rec["value"], g.net["rec_value"], 2e-4).all()
        assert _close(rec["reward"], g.net["rec_reward"], 2e-4).all()
    else:
        dyn = orc.stochastic_dynamics(h, a)
        assert _close(dyn["transition_logits"], g.net["dyn_transition_logits"]).all()
        assert _close(dyn["hidden"], g.net["dyn_hidden"]).all()
        assert _close(dyn["reward"], g.net["dyn_reward"], 2e-4).all()
        pred = orc.prediction(h)
        assert _close(pred["policy_logits"], g.net["pred_policy_logits"]).all()
        assert _close(pred["value"], g.net["pred_value"], 2e-4).all()


def test_oracle_threads_do_not_change_results():
    g = Golden("roundabout_s50")
    orc = Oracle(g.config, g.state_dict)
    rng = numpy.random.default_rng(7)
    B = 96
    obs = rng.random((B, g.obs.shape[1]), dtype=numpy.float32)
    noise = rng.dirichlet([0.25] * 5, size=B)
    a = orc.run(20, observations=obs, noise=noise, n_threads=1)
    b = orc.run(20, observations=obs, noise=noise, n_threads=4)
    assert_tree_outputs_bit_exact(a, b)
