"""GPU (-m gpu): the tcgen05 / tensor-memory search kernel (tsp::search_tc_kernel, csrc/tsp_tc.cuh).

The kernel replaces only the MLP of tsp::search_fast_kernel; selection / expansion / backup are the same code.
Checked here, through the C ABI:
  * the dispatcher really runs it (stats.search_kernel == 2) for the fully-connected MuZero nets of the
    highway / roundabout game files in both group shapes, and keeps the mma.sync kernels where the network is
    outside its envelope (terminal head, cross-merge); the stochastic and risk-sensitive searches run their two
    networks (one future of a transition expansion, prediction) on it as well;
  * the engine's own network outputs replayed through the oracle's tree code give bit-identical root statistics,
    and the traced network outputs agree with the oracle's float32 network within 1e-5 (check_against_oracle);
  * it agrees with the mma.sync kernel on the same inputs (both are 3xTF32 float32-accurate networks);
  * other layer widths (tables for replication / lane packing are built per network).
"""
import numpy
import pytest

torch = pytest.importorskip("torch")

import tree_search_planning_b200 as tsp  # noqa: E402
from tree_search_planning_b200.synthetic import random_state_dict, synthetic_inputs  # noqa: E402
from test_gpu_parity import check_against_oracle  # noqa: E402
from golden_util import assert_records_close, undisturbed_trees  # noqa: E402

pytestmark = pytest.mark.gpu


def search_kernel_of(cfg, B, S, monkeypatch=None, **env):
    if monkeypatch:
        for k, v in env.items():
            monkeypatch.setenv(k, v)
    eng = tsp.Engine(cfg, max_roots=B, max_simulations=S, state_dict=random_state_dict(cfg, 3))
    obs, kw = synthetic_inputs(cfg, B, S, seed=3)
    out = eng.run(S, observations=obs, trace=True, **kw)
    st = eng.stats()
    eng.close()
    return st, out


@pytest.mark.parametrize("B,shape", [(4096, (32, 512)), (300, (32, 512)), (16384, (64, 512))])
def test_dispatch_runs_the_tcgen05_kernel(B, shape):
    cfg = tsp.get_config("roundabout_env_ttc_flat", num_simulations=20)
    st, out = search_kernel_of(cfg, B, 20)
    assert st["search_kernel"] == 2
    assert (st["trees_per_cta"], st["threads_per_cta"]) == shape
    assert (out["visit_counts"].sum(1) == 20).all()


def test_outside_the_envelope_keeps_the_mma_sync_kernels(monkeypatch):
    for name, opts in (("roundabout_env_ttc_flat", dict(mask_absorbing_states=True)),
                       ("cross_merge_env", dict(mask_absorbing_states=True)),
                       ("roundabout_env_ttc_flat_stochastic", {})):  # terminal head; one dynamics MLP per future
        cfg = tsp.get_config(name, num_simulations=10, **opts)
        st, _ = search_kernel_of(cfg, 256, 10)
        assert st["search_kernel"] in (0, 1), name
    # cross-merge (weights beyond tensor memory): tcgen05 with the weights streamed by TMA, unless switched off
    cfg = tsp.get_config("cross_merge_env", num_simulations=10)
    st, out = search_kernel_of(cfg, 256, 10)
    assert st["search_kernel"] == 3 and (st["trees_per_cta"], st["threads_per_cta"]) == (64, 576)
    assert (out["visit_counts"].sum(1) == 10).all()
    st, _ = search_kernel_of(cfg, 256, 10, monkeypatch, TSP_NO_TCS="1")
    assert st["search_kernel"] in (0, 1)
    monkeypatch.delenv("TSP_NO_TCS")
    cfg = tsp.get_config("roundabout_env_ttc_flat", num_simulations=10)
    st, _ = search_kernel_of(cfg, 256, 10, monkeypatch, TSP_NO_TC="1")
    assert st["search_kernel"] == 1


@pytest.mark.parametrize("B,S", [(4096, 50), (8192, 25)])
def test_agrees_with_the_mma_sync_kernel(B, S, monkeypatch):
    cfg = tsp.get_config("highway_env_ttc_flat", num_simulations=S)
    st_tc, tc = search_kernel_of(cfg, B, S)
    st_mm, mm = search_kernel_of(cfg, B, S, monkeypatch, TSP_NO_TC="1")
    assert st_tc["search_kernel"] == 2 and st_mm["search_kernel"] == 1
    same = (tc["visit_counts"] == mm["visit_counts"]).all(axis=1)
    assert same.mean() >= 0.985, same.mean()
    a = tc["trace_records"][same].astype(numpy.float64)
    b = mm["trace_records"][same].astype(numpy.float64)
    err = numpy.abs(a - b) / numpy.maximum(1.0, numpy.abs(b))
    assert numpy.median(err) <= 1e-6
    # priors within 1e-5 on every entry of every tree whose evaluation order was not disturbed (golden_util.undisturbed_trees),
    # and those are nearly all; only value / reward get the quantisation allowance
    keep = undisturbed_trees(a, b, 1)
    assert keep.mean() >= 0.98, keep.mean()
    assert_records_close(a[keep], b[keep], 1, what="tcgen05 vs mma.sync records")


@pytest.mark.parametrize("B", [2048, 8192])
def test_trace_replays_bit_exactly_and_network_matches_oracle(B):
    cfg = tsp.get_config("highway_env_ttc_flat", num_simulations=30)
    check_against_oracle(cfg, B, 30, seed=21)


@pytest.mark.parametrize("widths", [dict(encoding_size=32, fc_dynamics_layers=[16], fc_reward_layers=[16], fc_value_layers=[24],
                                         fc_policy_layers=[8], support_size=5),
                                    dict(encoding_size=48, fc_dynamics_layers=[24], fc_reward_layers=[32], fc_value_layers=[16],
                                         fc_policy_layers=[32], support_size=15),
                                    dict(encoding_size=64, fc_dynamics_layers=[32], fc_reward_layers=[8], fc_value_layers=[8],
                                         fc_policy_layers=[8], support_size=1)])
@pytest.mark.parametrize("B", [700, 9000])
def test_other_layer_widths(widths, B):
    cfg = tsp.get_config("roundabout_env_ttc_flat", num_simulations=20, **widths)
    eng = tsp.Engine(cfg, max_roots=B, max_simulations=20, state_dict=random_state_dict(cfg, 5))
    obs, kw = synthetic_inputs(cfg, 16, 20, seed=5)
    eng.run(20, observations=obs, **kw)
    assert eng.stats()["search_kernel"] == 2
    eng.close()
    check_against_oracle(cfg, B, 20, seed=22, legal=True)


def test_ragged_uniform_and_two_players_on_tcgen05():
    cfg = tsp.get_config("roundabout_env_ttc_flat", num_simulations=25)
    check_against_oracle(cfg, 1000, 25, seed=23, legal=True)
    check_against_oracle(cfg, 1000, 25, seed=24, uniform_policy=True, min_agree=0.9)
    check_against_oracle(cfg.replace(players=[0, 1]), 1000, 25, seed=25, min_agree=0.9)


@pytest.mark.parametrize("game", ["roundabout_env_ttc_flat_stochastic_concat", "roundabout_env_ttc_flat_risk_sensitive"])
@pytest.mark.parametrize("B", [1500, 9000])
def test_chance_node_searches_on_tcgen05(game, B):
    """Transition expansions (nf futures through dynamics + reward head + prior head) and state expansions
    (prediction) on the tensor-memory path: trace replay bit-exact, network within 1e-5 of the oracle's."""
    cfg = tsp.get_config(game)
    st, out = search_kernel_of(cfg, B, 25)
    assert st["search_kernel"] == 2
    assert (st["trees_per_cta"], st["threads_per_cta"]) == ((32, 512) if B <= 16 * 2 * st["sm_count"] else (64, 512))
    assert (out["visit_counts"].sum(1) == 25).all()
    check_against_oracle(cfg, B, 25, seed=31, legal=(B == 1500))


@pytest.mark.parametrize("game", ["roundabout_env_ttc_flat_stochastic_concat", "roundabout_env_ttc_flat_risk_sensitive"])
def test_chance_node_searches_agree_with_the_mma_sync_kernel(game, monkeypatch):
    cfg = tsp.get_config(game)
    st_tc, tc = search_kernel_of(cfg, 4096, 25)
    st_mm, mm = search_kernel_of(cfg, 4096, 25, monkeypatch, TSP_NO_TC="1")
    assert st_tc["search_kernel"] == 2 and st_mm["search_kernel"] == 1
    same = (tc["visit_counts"] == mm["visit_counts"]).all(axis=1)
    assert same.mean() >= 0.97, same.mean()


@pytest.mark.parametrize("game", ["roundabout_env_ttc_flat_stochastic_concat", "roundabout_env_ttc_flat_risk_sensitive"])
@pytest.mark.parametrize("B", [600, 9000])
def test_both_futures_in_one_pass_match_one_future_per_pass(game, B, monkeypatch):
    """Two futures: the transition expansion evaluates both side by side (MMA columns 16..31) together with the prediction of the
    state expansions, three stages per iteration; TSP_TC_NOMERGE=1 keeps one future per pass. Same arithmetic per column:
    the searches must agree."""
    cfg = tsp.get_config(game)
    st_m, m = search_kernel_of(cfg, B, 25)
    st_s, s = search_kernel_of(cfg, B, 25, monkeypatch, TSP_TC_NOMERGE="1")
    assert st_m["search_kernel"] == 2 and st_s["search_kernel"] == 2
    same = (m["visit_counts"] == s["visit_counts"]).all(axis=1)
    assert same.mean() >= 0.995, same.mean()
    a = m["trace_records"][same].astype(numpy.float64)
    b = s["trace_records"][same].astype(numpy.float64)
    assert numpy.median(numpy.abs(a - b)) <= 1e-7
    # the accumulators may differ in the last bit between the two MMA shapes: priors / probabilities within 1e-5 everywhere,
    # value / reward with the one-step allowance of the float32 inverse value transform (golden_util.quantised_columns)
    keep = undisturbed_trees(a, b, cfg.n_futures)
    assert keep.mean() >= 0.98, keep.mean()
    assert_records_close(a[keep], b[keep], cfg.n_futures, min_fraction=0.995, what="both futures per pass vs one future per pass")


@pytest.mark.parametrize("widths", [dict(encoding_size=32, fc_dynamics_layers=[16], fc_reward_layers=[16], fc_value_layers=[24],
                                         fc_policy_layers=[8], fc_prior_layers=[16], support_size=5),
                                    dict(encoding_size=32, fc_dynamics_layers=[8], fc_reward_layers=[8], fc_value_layers=[8],
                                         fc_policy_layers=[8], fc_prior_layers=[8], support_size=2),
                                    dict(encoding_size=48, fc_dynamics_layers=[24], fc_reward_layers=[32], fc_value_layers=[16],
                                         fc_policy_layers=[32], fc_prior_layers=[24], support_size=15),
                                    dict(encoding_size=64, fc_dynamics_layers=[16], fc_reward_layers=[32], fc_value_layers=[32],
                                         fc_policy_layers=[32], fc_prior_layers=[32], support_size=10)])
@pytest.mark.parametrize("game", ["roundabout_env_ttc_flat_stochastic_concat", "roundabout_env_ttc_flat_risk_sensitive"])
@pytest.mark.parametrize("B", [300, 9000])
def test_chance_node_searches_other_layer_widths(widths, game, B):
    """The merged schedule's tables depend on the widths: one or two lane quarters of next state, the un-normalised states
    aliased onto the dead gathered-state + hidden-layer buffers or in a buffer of their own (H > 2 x dynamics width)."""
    cfg = tsp.get_config(game, **widths)
    st, out = search_kernel_of(cfg, B, 25)
    assert st["search_kernel"] == 2
    assert (out["visit_counts"].sum(1) == 25).all()
    check_against_oracle(cfg, B, 25, seed=34, legal=(B == 300))


def test_three_futures_keep_one_future_per_pass():
    cfg = tsp.get_config("roundabout_env_ttc_flat_stochastic_concat", n_futures=3)
    st, out = search_kernel_of(cfg, 700, 25)
    assert st["search_kernel"] == 2
    assert (out["visit_counts"].sum(1) == 25).all()
    check_against_oracle(cfg, 700, 25, seed=33)


@pytest.mark.parametrize("game", ["roundabout_env_ttc_flat", "roundabout_env_ttc_flat_stochastic_concat",
                                  "roundabout_env_ttc_flat_risk_sensitive"])
@pytest.mark.parametrize("B", [1, 17, 33])
def test_odd_number_of_roots(game, B):
    """A warp holds two trees: with an odd batch one warp owns a valid and an invalid tree. The lanes of the invalid
    tree must reach the CTA barrier in front of tcgen05.dealloc together with the others (regression: a __syncwarp
    inside `if (valid)` deadlocked the kernel for every odd batch size)."""
    cfg = tsp.get_config(game)
    st, out = search_kernel_of(cfg, B, 25)
    assert st["search_kernel"] == 2
    assert (out["visit_counts"].sum(1) == 25).all()
    check_against_oracle(cfg, B, 25, seed=40 + B, min_agree=0.9)


def test_streamed_kernel_envelope_is_decided_at_weight_load():
    """The streamed tcgen05 kernels split operands into fp16 pairs: a network whose activations COULD leave fp16's range
    (worst case over all normalised states, from the weights alone) is searched by the mma.sync kernels instead, so the
    overflow flag of the streamed search kernel can never fire."""
    cfg = tsp.get_config("cross_merge_env", num_simulations=10)
    sd = random_state_dict(cfg, 3)
    key = "dynamics_encoded_state_network.module.2.weight"
    big = dict(sd)
    big[key] = sd[key] * 3.0e3  # un-normalised next states of magnitude ~1e5
    obs, kw = synthetic_inputs(cfg, 128, 10, seed=3)
    for weights, want_streamed in ((sd, True), (big, False)):
        eng = tsp.Engine(cfg, max_roots=128, max_simulations=10, state_dict=weights)
        out = eng.run(10, observations=obs, **kw)
        st = eng.stats()
        eng.close()
        assert (st["search_kernel"] == 3) == want_streamed, st["search_kernel"]
        assert (out["visit_counts"].sum(1) == 10).all()


@pytest.mark.parametrize("widths", [dict(encoding_size=128, fc_dynamics_layers=[64], fc_reward_layers=[16], fc_value_layers=[16],
                                         fc_policy_layers=[32], support_size=5),
                                    dict(encoding_size=192, fc_dynamics_layers=[128], fc_reward_layers=[32], fc_value_layers=[32],
                                         fc_policy_layers=[16], support_size=10),
                                    dict(encoding_size=256, fc_dynamics_layers=[64], fc_reward_layers=[32], fc_value_layers=[16],
                                         fc_policy_layers=[16], support_size=1)])
def test_streamed_kernel_other_layer_widths(widths):
    """The envelope of search_tcs_kernel beyond the cross-merge shapes: 64 < H <= 256 in steps of 64 (one or two
    next-state tiles, a partial second tile at H = 192), dynamics hidden width 64 / 128, head widths 16 / 32."""
    cfg = tsp.get_config("cross_merge_env", num_simulations=20, **widths)
    eng = tsp.Engine(cfg, max_roots=300, max_simulations=20, state_dict=random_state_dict(cfg, 5))
    obs, kw = synthetic_inputs(cfg, 16, 20, seed=5)
    eng.run(20, observations=obs, **kw)
    st = eng.stats()
    eng.close()
    assert st["search_kernel"] == 3 and st["root_kernel"] == 1
    check_against_oracle(cfg, 300, 20, seed=31, legal=True)


@pytest.mark.parametrize("over", [dict(fc_representation_layers=[64, 64, 32]), dict(fc_representation_layers=[128]),
                                  dict(fc_representation_layers=[32]), dict(encoding_size=128, fc_representation_layers=[64, 64, 32]),
                                  dict(encoding_size=256, fc_representation_layers=[256, 128, 128], fc_value_layers=[16]),
                                  dict(fc_representation_layers=[]), dict(encoding_size=48, fc_representation_layers=[80, 16])])
def test_root_kernel_shapes_against_the_oracle(over):
    """root_tcs_kernel (initial_inference on the streamed tcgen05 machinery) for representation networks of other depths /
    widths, rows that do not fill a CTA, against the oracle's float32 network."""
    from oracle.oracle import Oracle
    cfg = tsp.get_config("highway_env_ttc_flat", **over)
    sd = random_state_dict(cfg, 7)
    B = 333
    eng = tsp.Engine(cfg, max_roots=B, max_simulations=5, state_dict=sd)
    obs, _ = synthetic_inputs(cfg, B, 5, seed=7)
    got = eng.initial_inference(obs)
    st = eng.stats()
    eng.close()
    assert st["root_kernel"] == 1, over
    want = Oracle(cfg, sd).initial_inference(obs)
    for k in ("hidden", "policy_logits", "value_logits"):
        err = numpy.abs(got[k] - want[k]).max()
        assert err <= 1e-5, (over, k, err)
