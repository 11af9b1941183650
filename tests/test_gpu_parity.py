"""GPU (-m gpu): parity of the CUDA engine, called through the C ABI, against
  (1) the golden fixtures written by the LIVE reference (bit-exact tree arithmetic with the
      reference's network outputs fed identically; float32 network within 1e-5), and
  (2) the C oracle on seeded inputs up to the BASELINE.json sizes: the engine's own network
      outputs (trace) are replayed through the oracle's tree code and every root statistic must
      match BIT-EXACTLY, while the traced network outputs must agree with the oracle's float32
      network within tolerance.
"""
import numpy
import pytest

torch = pytest.importorskip("torch")

from golden_util import (Golden, golden_names, assert_tree_outputs_bit_exact, assert_close_quantised,  # noqa: E402
                         assert_records_close, undisturbed_trees)
import tree_search_planning_b200 as tsp  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402  (the checker)

pytestmark = pytest.mark.gpu
NAMES = golden_names()

_engines = {}


def engine_for(g, max_roots=64):
    key = (g.name, max_roots)
    if key not in _engines:
        _engines[key] = tsp.Engine(g.config, max_roots=max_roots, max_simulations=g.S * g.P, state_dict=g.state_dict)
    return _engines[key]


def engine_run(eng, g, fed, **opts):
    """One search per phase of the fixture: phase 0 from scratch, the others continue the device-resident
    trees (tsp_run_args.resume == MCTS.run's override_root_with). Returns the last run's outputs and the
    per-phase trace records."""
    traces = []
    for ph in range(g.P):
        kw = dict(g.run_kwargs(ph if g.P > 1 else None), **opts)
        recs = g.out["records"] if g.P == 1 else g.out["records"][ph]
        if ph == 0:
            src = dict(fed_root=g.out["root_record"], fed_records=recs) if fed else dict(observations=g.obs)
            got = eng.run(g.S, **src, **kw)
            first = got
        else:
            src = dict(fed_records=recs) if fed else {}
            got = eng.run(g.S, resume=True, num_roots=g.N, **src, **kw)
        if "trace_records" in got:
            traces.append(got["trace_records"])
    for k in ("trace_root", "root_hidden", "root_predicted_value"):  # written by the first search only
        if k in first:
            got[k] = first[k]
    return got, traces


def drop_terminal(records, cfg):
    """The engine evaluates the terminal head only under mask_absorbing_states (the reference
    computes it always and ignores it otherwise, self_play.py:408-417): blank that column."""
    r = numpy.array(records, copy=True)
    if not getattr(cfg, "mask_absorbing_states", False):
        r[..., 3] = 0
    return r


# Floors of the end-to-end visit-count agreement with the reference fixtures (fraction of the fixture's roots), pinned
# just under what the B200 run measured (see the AGREE lines of profiles/r2_gpu_tests.log) so that a regression shows.
# Measured: every fixture root agrees (1.0000 on all 18 fixtures); the kernels are deterministic, so the floor is "all".
GOLDEN_FLOOR = {}
GOLDEN_FLOOR_DEFAULT = 0.995


def close(a, b, tol=1e-5):
    a, b = numpy.asarray(a, numpy.float64), numpy.asarray(b, numpy.float64)
    return numpy.abs(a - b) <= tol * numpy.maximum(1.0, numpy.abs(b))


# ---------------------------------------------------------------- golden fixtures ----
@pytest.mark.parametrize("name", NAMES)
def test_tree_arithmetic_bit_exact_with_reference_outputs_fed(name):
    """select / ucb / expand / noise / ties / chance / backup / min-max on the GPU ==
    the reference, when the reference's own network outputs are fed in call order."""
    g = Golden(name)
    eng = engine_for(g)
    got, _ = engine_run(eng, g, fed=True)
    assert_tree_outputs_bit_exact(got, g.out, what=name)
    assert (got["num_records"] == g.out["n_records"]).all()
    assert (got["visit_counts"].sum(1) == g.S * g.P).all()


@pytest.mark.parametrize("name", NAMES)
def test_end_to_end_against_reference_golden(name):
    """Own float32 network + own tree: visit counts equal the reference's on the fixture roots."""
    g = Golden(name)
    eng = engine_for(g)
    got, traces = engine_run(eng, g, fed=False, trace=True, want_root_hidden=True)
    got["trace_records"] = traces[-1]
    same = (got["visit_counts"] == g.out["visit_counts"]).all(axis=1)
    assert same.mean() >= GOLDEN_FLOOR.get(name, GOLDEN_FLOOR_DEFAULT), "visit counts differ on %d of %d roots" % ((~same).sum(), g.N)
    st = eng.stats()
    print("AGREE golden %s %.4f kernel %d" % (name, same.mean(), st["search_kernel"]))
    # the fixtures inside the tcgen05 envelope must really run on the tcgen05 kernels (2: weights in tensor memory,
    # 3: weights streamed by TMA) -- this test must not silently stop covering them
    # (outside: the terminal head of mask_absorbing_states, and MuZeroStochastic's one dynamics MLP per future)
    if not getattr(g.config, "mask_absorbing_states", False) and getattr(g.config, "network", "") != "stochastic":
        assert st["search_kernel"] in (2, 3), (name, st["search_kernel"])
    assert close(got["root_hidden"], g.out["root_hidden"]).all()
    assert_close_quantised(got["trace_root"], g.out["root_record"], what="root record")
    assert_close_quantised(got["root_value"][same], g.out["root_value"][same], min_fraction=0.85, what="root value")
    assert_close_quantised(got["child_value"][same], g.out["child_value"][same], min_fraction=0.85, what="child value")
    assert close(got["child_prior"][same], g.out["child_prior"][same]).all()
    # every record of every agreeing tree at once: priors / probabilities within 1e-5 on each entry, value / reward entries
    # within the quantisation step; the SHARE of quantised entries off by a step is judged over the whole fixture (a tree
    # of 25 records has 50 of them: two steps -- ~1.3 % of the evaluations take one -- would already read as 4 %)
    idx = numpy.nonzero(same)[0]
    assert (got["num_records"][idx] == g.out["n_records"][idx]).all()
    ref_rec = g.out["records"] if g.P == 1 else g.out["records"][-1]
    m = min(got["trace_records"].shape[1], ref_rec.shape[1])
    live = numpy.arange(m)[None, :] < g.out["n_records"][idx][:, None]
    assert_records_close(drop_terminal(got["trace_records"][idx, :m], g.config), drop_terminal(ref_rec[idx, :m], g.config),
                         getattr(g.config, "n_futures", 1), live=live, what="records")


@pytest.mark.parametrize("name", ["highway_s25", "cross_merge_s200", "stochastic_s25", "stochastic_sep_s25", "stochastic_sep_s25_legal"])
def test_network_parity_against_torch_golden(name):
    g = Golden(name)
    eng = engine_for(g)
    M = g.net["init_hidden"].shape[0]
    ini = eng.initial_inference(g.obs[:M])
    assert close(ini["hidden"], g.net["init_hidden"]).all()
    assert close(ini["policy_logits"], g.net["init_policy_logits"]).all()
    assert close(ini["value_logits"], g.net["init_value_logits"]).all()
    assert_close_quantised(ini["value"], g.net["init_value"], min_fraction=0.9)
    h, a = g.net["rec_in_hidden"], g.net["rec_in_action"]
    if "rec_hidden" in g.net:
        rec = eng.recurrent_inference(h, a)
        for k_o, k_g in (("hidden", "rec_hidden"), ("policy_logits", "rec_policy_logits"),
                         ("value_logits", "rec_value_logits"), ("reward_logits", "rec_reward_logits")):
            assert close(rec[k_o], g.net[k_g]).all(), k_o
        assert_close_quantised(rec["value"], g.net["rec_value"], min_fraction=0.9)
        assert_close_quantised(rec["reward"], g.net["rec_reward"], min_fraction=0.9)
    else:
        dyn = eng.stochastic_dynamics(h, a)
        assert close(dyn["transition_logits"], g.net["dyn_transition_logits"]).all()
        assert close(dyn["hidden"], g.net["dyn_hidden"]).all()
        assert_close_quantised(dyn["reward"], g.net["dyn_reward"], min_fraction=0.9)
        pred = eng.prediction(h)
        assert close(pred["policy_logits"], g.net["pred_policy_logits"]).all()
        assert_close_quantised(pred["value"], g.net["pred_value"], min_fraction=0.9)


@pytest.mark.parametrize("game", ["roundabout_env_ttc_flat", "cross_merge_env",
                                  "roundabout_env_ttc_flat_stochastic_concat"])
def test_network_parity_large_batch_vs_oracle(game):
    """Trajectory-independent: every row of a 5,000-row batch through each fused MLP kernel."""
    cfg = tsp.get_config(game)
    sd = random_state_dict(cfg, 31)
    B = 5000
    eng = tsp.Engine(cfg, max_roots=B, state_dict=sd)
    orc = Oracle(cfg, sd)
    rng = numpy.random.default_rng(32)
    obs = rng.random((B, tsp.observation_dim(cfg)), dtype=numpy.float32)
    g, w = eng.initial_inference(obs), orc.initial_inference(obs)
    for k in ("hidden", "policy_logits", "value_logits"):
        assert close(g[k], w[k]).all(), k
    assert_close_quantised(g["value"], w["value"], min_fraction=0.97)
    h = w["hidden"]
    a = rng.integers(0, 5, B)
    if tsp.variant_of(cfg) == 0:
        g, w = eng.recurrent_inference(h, a), orc.recurrent_inference(h, a)
        for k in ("hidden", "policy_logits", "value_logits", "reward_logits"):
            assert close(g[k], w[k]).all(), k
        assert_close_quantised(g["value"], w["value"], min_fraction=0.97)
        assert_close_quantised(g["reward"], w["reward"], min_fraction=0.97)
    else:
        g, w = eng.stochastic_dynamics(h, a), orc.stochastic_dynamics(h, a)
        assert close(g["hidden"], w["hidden"]).all()
        assert close(g["transition_logits"], w["transition_logits"]).all()
        assert_close_quantised(g["reward"], w["reward"], min_fraction=0.97)
        g, w = eng.prediction(h), orc.prediction(h)
        assert close(g["policy_logits"], w["policy_logits"]).all()
        assert_close_quantised(g["value"], w["value"], min_fraction=0.97)
    eng.close()


def test_terminal_head_parity():
    g = Golden("roundabout_s25_mask")
    eng = engine_for(g)
    orc = Oracle(g.config, g.state_dict)
    rng = numpy.random.default_rng(3)
    h = rng.random((100, 64), dtype=numpy.float32)
    a = rng.integers(0, 5, 100)
    got, want = eng.recurrent_inference(h, a), orc.recurrent_inference(h, a)
    assert close(got["terminal"], want["terminal"]).all()


# ------------------------------------------------------------ seeded, vs the oracle ----
from tree_search_planning_b200.synthetic import random_state_dict, synthetic_inputs  # noqa: E402


def seeded_inputs(cfg, B, S, seed, legal=False):
    return synthetic_inputs(cfg, B, S, seed, ragged_legal=legal)


def check_against_oracle(cfg, B, S, seed, legal=False, min_agree=0.985, **run_opts):
    """`min_agree`: floor of the visit-count agreement AND 1 - ceiling of the share of disturbed trees. The BASELINE-size
    tests pin it half a point under what the B200 run measured (AGREE lines of profiles/r2_gpu_tests.log)."""
    sd = random_state_dict(cfg, seed)
    eng = tsp.Engine(cfg, max_roots=B, max_simulations=S, state_dict=sd)
    orc = Oracle(cfg, sd)
    obs, kw = seeded_inputs(cfg, B, S, seed, legal)
    kw.update(run_opts)
    got = eng.run(S, observations=obs, trace=True, **kw)
    # size-independent invariants (self_play_stochastic.py:333; root.visit_count == S)
    assert (got["visit_counts"].sum(1) == S).all()
    assert (got["max_tree_depth"] >= 1).all() and (got["max_tree_depth"] <= 2 * S + 1).all()
    # (a) the engine's tree arithmetic is bit-exact given ITS OWN network outputs
    replay = orc.run(S, fed_root=got["trace_root"], fed_records=got["trace_records"], B=B, **kw)
    assert_tree_outputs_bit_exact(got, replay, what="replay of the engine's trace through the oracle")
    assert (replay["n_records"] == got["num_records"]).all()
    # (b) end to end against the oracle's own float32 network. Two float32 networks that differ in
    # the last ulp (and hence by one quantum of the reference's float32 inverse value transform on
    # ~1 % of the evaluations, see golden_util.assert_close_quantised) can flip a near-tie deep in a
    # tree -- a few % of the trees at S = 200 --, after which that tree's evaluation sequence is
    # a different (equally valid) one; such trees are counted, bounded, and excluded from the
    # record-by-record comparison.
    want = orc.run(S, observations=obs, trace=True, **kw)
    same = (got["visit_counts"] == want["visit_counts"]).all(axis=1)
    assert same.mean() >= min_agree, "visit counts agree on only %.4f of %d roots" % (same.mean(), B)
    assert_close_quantised(got["trace_root"], want["trace_root"], min_fraction=0.97, what="root records")
    a = drop_terminal(got["trace_records"], cfg).astype(numpy.float64)
    w = drop_terminal(want["trace_records"], cfg).astype(numpy.float64)
    err = numpy.abs(a - w) / numpy.maximum(1.0, numpy.abs(w))
    live = numpy.arange(a.shape[1])[None, :] < numpy.minimum(want["n_records"], got["num_records"])[:, None]
    err = err * live[:, :, None]
    # trees whose evaluation order was disturbed by a flipped near-tie (golden_util.undisturbed_trees) are bounded, the
    # others hold 1e-5 on every prior / probability entry
    diverged = ~undisturbed_trees(a, w, getattr(cfg, "n_futures", 1), live=live) | (want["n_records"] != got["num_records"])
    assert diverged.mean() <= 1.0 - min_agree, "%.4f of the trees diverged" % diverged.mean()
    keep = ~diverged
    if keep.any():
        # priors / transition probabilities within 1e-5 on EVERY entry of the trees that did not diverge; only value /
        # reward entries get the quantisation allowance
        assert_records_close(a[keep], w[keep], getattr(cfg, "n_futures", 1), live=live[keep], what="trace records")
        assert_close_quantised(got["root_value"][keep], want["root_value"][keep], min_fraction=0.85, what="root value")
    print("AGREE oracle %s B=%d S=%d visit-count agreement %.4f diverged %.4f kernel %d" % (
        cfg.name, B, S, same.mean(), diverged.mean(), eng.stats()["search_kernel"]))
    eng.close()
    return float(same.mean())


def test_config2_roundabout_4096_roots_50_sims():
    cfg = tsp.get_config("roundabout_env_ttc_flat", num_simulations=50)
    check_against_oracle(cfg, 4096, 50, seed=11, min_agree=0.989)  # measured 0.9990 / 0.59 % disturbed


def test_config3_stochastic_16384_roots():
    cfg = tsp.get_config("roundabout_env_ttc_flat_stochastic_concat")
    check_against_oracle(cfg, 16384, 25, seed=12, min_agree=0.994)  # measured 0.9998 / 0.05 % disturbed


def test_config4_risk_sensitive_16384_roots():
    cfg = tsp.get_config("roundabout_env_ttc_flat_risk_sensitive")
    check_against_oracle(cfg, 16384, 25, seed=13, min_agree=0.994)  # measured 0.9999 / 0.02 % disturbed


def test_separate_heads_stochastic_network():
    """models.MuZeroStochastic (one dynamics MLP per future, models.py:298-462; 4 futures in
    roundabout_env_ttc_flat_stochastic.py:66, 2 in highway_env_ttc_flat_stochastic.py)."""
    cfg = tsp.get_config("roundabout_env_ttc_flat_stochastic")
    check_against_oracle(cfg, 2048, 25, seed=41)
    cfg = tsp.get_config("highway_env_ttc_flat_stochastic")
    check_against_oracle(cfg, 777, 25, seed=42, legal=True, min_agree=0.97)


def test_config5_cross_merge_200_sims_slice():
    cfg = tsp.get_config("cross_merge_env", num_simulations=200)
    check_against_oracle(cfg, 1024, 200, seed=14, min_agree=0.96)  # measured 0.9980 / 3.1 % disturbed (200 simulations deep)


def _full_size_properties(game, B, S, seed, monkeypatch, slice_roots=192):
    """BASELINE.json's full batch sizes, through size-independent properties: every root completes (sum of the
    root's child visits == S, self_play_stochastic.py:333), depths are in range, the search is deterministic
    (bit-identical on a second run), and a root's result does not depend on the batch it is in -- a slice of the
    roots searched on its own, with its own trace replayed through the oracle, reproduces the big batch bit for bit."""
    cfg = tsp.get_config(game, num_simulations=S)
    sd = random_state_dict(cfg, seed)
    eng = tsp.Engine(cfg, max_roots=B, max_simulations=S, state_dict=sd)
    obs, kw = seeded_inputs(cfg, B, S, seed)
    big = eng.run(S, observations=obs, **kw)
    assert (big["visit_counts"].sum(1) == S).all()
    assert (big["max_tree_depth"] >= 1).all() and (big["max_tree_depth"] <= S).all()
    assert numpy.isfinite(big["root_value"]).all() and numpy.isfinite(big["child_value"]).all()
    again = eng.run(S, observations=obs, **kw)
    assert_tree_outputs_bit_exact(again, big, what="second run")
    rng = numpy.random.default_rng(seed)
    pick = numpy.sort(rng.choice(B, size=slice_roots, replace=False))
    sub_kw = {k: (v[pick] if isinstance(v, numpy.ndarray) and v.shape[:1] == (B,) else v) for k, v in kw.items()}
    # the dispatcher picks the kernel shape from the batch size, and float32 reductions of different shapes round
    # differently: the slice is searched by the kernel shape the big batch used
    st = eng.stats()
    eng.close()
    if st["search_kernel"]:
        monkeypatch.setenv("TSP_FAST_LPT", str(st["threads_per_cta"] // st["trees_per_cta"]))
    else:
        monkeypatch.setenv("TSP_NO_WGLOBAL", "1")
    eng = tsp.Engine(cfg, max_roots=slice_roots, max_simulations=S, state_dict=sd)
    small = eng.run(S, observations=obs[pick], trace=True, **sub_kw)
    assert_tree_outputs_bit_exact(small, {k: big[k][pick] for k in big if k in small and k.startswith(("visit", "root_value", "child", "max"))},
                                  what="slice searched on its own")
    replay = Oracle(cfg, sd).run(S, fed_root=small["trace_root"], fed_records=small["trace_records"], B=slice_roots, **sub_kw)
    assert_tree_outputs_bit_exact(small, replay, what="slice replayed through the oracle")
    eng.close()


def test_config5_cross_merge_65536_roots_200_sims_full_size(monkeypatch):
    _full_size_properties("cross_merge_env", 65536, 200, seed=31, monkeypatch=monkeypatch)


def test_north_star_highway_65536_roots_full_size(monkeypatch):
    _full_size_properties("highway_env_ttc_flat", 65536, 50, seed=32, monkeypatch=monkeypatch)


def test_highway_200_sims():
    cfg = tsp.get_config("highway_env_ttc_flat", num_simulations=200)
    check_against_oracle(cfg, 1024, 200, seed=15, min_agree=0.9)


@pytest.mark.parametrize("game", ["roundabout_env_ttc_flat", "roundabout_env_ttc_flat_stochastic_concat",
                                  "roundabout_env_ttc_flat_risk_sensitive"])
def test_ragged_legal_actions_and_odd_batch(game):
    cfg = tsp.get_config(game)
    check_against_oracle(cfg, 37, 25, seed=21, legal=True, min_agree=0.9)


def test_uniform_policy_many_ties():
    cfg = tsp.get_config("roundabout_env_ttc_flat")
    check_against_oracle(cfg, 200, 25, seed=22, uniform_policy=True, min_agree=0.9)


def test_no_exploration_noise_and_no_tie_stream():
    cfg = tsp.get_config("roundabout_env_ttc_flat")
    check_against_oracle(cfg, 100, 25, seed=23, add_exploration_noise=False, tie=None, min_agree=0.9)


def test_mask_absorbing_states():
    cfg = tsp.get_config("roundabout_env_ttc_flat", mask_absorbing_states=True)
    check_against_oracle(cfg, 300, 25, seed=24, min_agree=0.9)


def test_two_players():
    """players == [0, 1]: -child.value() in ucb_score and the sign-alternating backup (self_play.py:469-472, 492-500)."""
    cfg = tsp.get_config("roundabout_env_ttc_flat", players=[0, 1])
    check_against_oracle(cfg, 512, 25, seed=26, min_agree=0.9)
    cfg = tsp.get_config("roundabout_env_ttc_flat", players=[0, 1], mask_absorbing_states=True)
    check_against_oracle(cfg, 300, 25, seed=27, legal=True, min_agree=0.9)
    cfg = tsp.get_config("cross_merge_env", players=[0, 1], num_simulations=60)  # weights too large for the fast kernel
    check_against_oracle(cfg, 64, 60, seed=28, min_agree=0.9)


def test_streamed_weights_fast_kernel_large_network(monkeypatch):
    """cross-merge (365 KB of search weights) on the fast kernel with the weights streamed from L2, both group
    shapes, forced also at a batch size where the dispatcher would pick the generic kernel."""
    monkeypatch.setenv("TSP_FORCE_WGLOBAL", "1")
    monkeypatch.setenv("TSP_NO_TCS", "1")  # (the dispatcher's choice for this network is the streamed tcgen05 kernel)
    cfg = tsp.get_config("cross_merge_env", num_simulations=40)
    for lpt in ("16", "8"):
        monkeypatch.setenv("TSP_FAST_LPT", lpt)
        check_against_oracle(cfg, 700, 40, seed=29, min_agree=0.9)


def test_two_players_rejected_by_the_chance_node_searches():
    cfg = tsp.get_config("roundabout_env_ttc_flat_stochastic_concat", players=[0, 1])  # self_play_stochastic.py:407
    with pytest.raises(tsp.TspError):
        tsp.Engine(cfg, max_roots=1)


@pytest.mark.parametrize("name", [n for n in NAMES if not n.startswith("cross_merge")])
def test_generic_kernel_fallback_bit_exact(name, monkeypatch):
    """The schedule-driven generic search kernel (what runs when a network does not fit the fast kernel's
    shared-memory layout) against the reference fixtures, with the fast kernel switched off."""
    monkeypatch.setenv("TSP_NO_FAST", "1")
    g = Golden(name)
    eng = tsp.Engine(g.config, max_roots=64, max_simulations=g.S * g.P, state_dict=g.state_dict)
    got, _ = engine_run(eng, g, fed=True)
    assert_tree_outputs_bit_exact(got, g.out, what=name)
    got, traces = engine_run(eng, g, fed=False, trace=True)
    replay = Oracle(g.config).run(g.S, fed_root=got["trace_root"], fed_records=numpy.stack(traces) if g.P > 1 else traces[0],
                                  B=g.N, **g.oracle_kwargs())
    assert_tree_outputs_bit_exact(got, replay, what=name + " (own network, replayed)")
    eng.close()


@pytest.mark.parametrize("game,B,kernel", [("roundabout_env_ttc_flat", 2048, 2), ("cross_merge_env", 333, 3),
                                           ("roundabout_env_ttc_flat_risk_sensitive", 1024, 2)])
def test_continued_search_large_batch_vs_oracle(game, B, kernel):
    """resume (override_root_with) at batch scale: 3 x 20 simulations on 2,048 roots (tensor-memory kernel) / 333 roots
    (streamed kernel), every phase with fresh noise and streams; the engine's traces replayed through the oracle must
    give bit-identical root statistics."""
    cfg = tsp.get_config(game, num_simulations=20)
    sd = random_state_dict(cfg, 41)
    S, P = 20, 3
    eng = tsp.Engine(cfg, max_roots=B, max_simulations=S * P, state_dict=sd)
    obs, kw0 = seeded_inputs(cfg, B, S, 42)
    phases = [kw0] + [seeded_inputs(cfg, B, S, 43 + i)[1] for i in range(P - 1)]
    traces = []
    for ph, kw in enumerate(phases):
        got = eng.run(S, observations=obs, trace=True, **kw) if ph == 0 else \
            eng.run(S, resume=True, num_roots=B, trace=True, **kw)
        traces.append(got["trace_records"])
        if ph == 0:
            troot = got["trace_root"]
    assert (got["visit_counts"].sum(1) == S * P).all()
    assert eng.stats()["search_kernel"] == kernel
    stack = lambda k: numpy.stack([kw[k] for kw in phases])
    streams = dict(noise=stack("noise"), tie=stack("tie"))
    if "chance" in phases[0]:  # the risk-sensitive search continues a tree too (self_play_risk_sensitive.py has no sims assert)
        streams["chance"] = stack("chance")
    replay = Oracle(cfg, sd).run(S, fed_root=troot, fed_records=numpy.stack(traces), B=B, n_phases=P, **streams)
    assert_tree_outputs_bit_exact(got, replay, what="continued search replayed through the oracle")
    with pytest.raises(tsp.TspError, match="exceed max_simulations"):
        eng.run(S, resume=True, num_roots=B, **phases[0])
    with pytest.raises(tsp.TspError, match="held"):
        eng.run(S, resume=True, num_roots=B - 1, **{k: v[:B - 1] for k, v in phases[0].items()})
    eng.close()
    cfg = tsp.get_config("roundabout_env_ttc_flat_stochastic_concat")
    eng = tsp.Engine(cfg, max_roots=4, max_simulations=50, state_dict=random_state_dict(cfg, 1))
    obs, kw = seeded_inputs(cfg, 4, 25, 1)
    eng.run(25, observations=obs, **kw)
    with pytest.raises(tsp.TspError, match="cannot be continued"):  # self_play_stochastic.py:333
        eng.run(25, resume=True, num_roots=4, **kw)
    eng.close()


def test_single_root_and_policy_only():
    cfg = tsp.get_config("highway_env_ttc_flat")
    check_against_oracle(cfg, 1, 25, seed=25, min_agree=1.0)
    sd = random_state_dict(cfg, 1)
    eng = tsp.Engine(cfg, max_roots=4, state_dict=sd)
    obs, kw = seeded_inputs(cfg, 4, 25, 1)
    got = eng.run(0, observations=obs, **kw)  # policy_only: S == 0 (self_play.py:381)
    assert (got["visit_counts"] == 0).all() and (got["root_value"] == 0).all()
    want = Oracle(cfg, sd).run(0, observations=obs, **kw)
    assert close(got["child_prior"], want["child_prior"]).all()
    eng.close()


def test_error_behaviour():
    cfg = tsp.get_config("highway_env_ttc_flat")
    eng = tsp.Engine(cfg, max_roots=4, state_dict=random_state_dict(cfg, 1))
    obs, kw = seeded_inputs(cfg, 4, 25, 1)
    bad = dict(kw, legal_actions=numpy.full((4, 5), 7, numpy.int32), num_legal=numpy.full(4, 2, numpy.int32))
    with pytest.raises(tsp.TspError, match="subset of the action space"):
        eng.run(25, observations=obs, **bad)
    empty = dict(kw, legal_actions=numpy.zeros((4, 5), numpy.int32), num_legal=numpy.zeros(4, numpy.int32))
    with pytest.raises(tsp.TspError, match="should not be an empty array"):
        eng.run(25, observations=obs, **empty)
    with pytest.raises(tsp.TspError, match="outside"):
        eng.run(26, observations=obs, **kw)
    with pytest.raises((tsp.TspError, ValueError)):
        eng.run(25, observations=numpy.zeros((5, obs.shape[1]), numpy.float32), **kw)
    eng.close()


# ----------------------------------------------------------- the reference surface ----
class _FakeModel:
    def __init__(self, sd):
        self._sd = sd

    def state_dict(self):
        return self._sd

    def parameters(self):
        return []


@pytest.mark.parametrize("game", ["highway_env_ttc_flat", "roundabout_env_ttc_flat_stochastic_concat",
                                  "roundabout_env_ttc_flat_risk_sensitive"])
def test_mcts_run_surface(game):
    cfg = tsp.get_config(game)
    model = _FakeModel(random_state_dict(cfg, 5))
    obs = numpy.random.default_rng(0).random((3, 1, 155))
    numpy.random.seed(0)
    root, extra = tsp.MCTS(cfg).run(model, obs, [3, 1, 4], 0, True)
    assert list(root.children.keys()) == [3, 1, 4]  # insertion order of legal_actions
    assert sum(c.visit_count for c in root.children.values()) == cfg.num_simulations
    assert root.visit_count == cfg.num_simulations
    assert root.hidden_state.shape == (1, cfg.encoding_size)
    assert "max_tree_depth" in extra and "root_predicted_value" in extra
    assert ("n_env_interactions" in extra) == (tsp.variant_of(cfg) == 0)
    v = root.value if tsp.variant_of(cfg) == 2 else root.value()
    assert numpy.isfinite(v)
    for c in root.children.values():
        assert 0.0 <= c.prior <= 1.0
        assert numpy.isfinite(c.value if tsp.variant_of(cfg) == 2 else c.value())
    with pytest.raises(AssertionError):
        tsp.MCTS(cfg).run(model, obs, [], 0, True)
    with pytest.raises(AssertionError):
        tsp.MCTS(cfg).run(model, obs, [0, 9], 0, True)
    # override_root_with (self_play.py:331-333): the same root searched further
    mcts = tsp.MCTS(cfg, max_search_multiple=2)
    root, _ = mcts.run(model, obs, [3, 1, 4], 0, True)
    if tsp.variant_of(cfg) == 1:
        with pytest.raises(AssertionError):  # the reference's own assert, self_play_stochastic.py:333
            mcts.run(model, obs, [3, 1, 4], 0, True, override_root_with=root)
    else:
        before = {a: c.visit_count for a, c in root.children.items()}
        root2, extra2 = mcts.run(model, obs, [3, 1, 4], 0, True, override_root_with=root)
        assert extra2["root_predicted_value"] is None
        assert root2.visit_count == 2 * cfg.num_simulations
        assert all(root2.children[a].visit_count >= before[a] for a in before)
        assert root2.hidden_state is not None
        stale = root  # a root whose tree has been searched over since
        with pytest.raises(ValueError):
            mcts.run(model, obs, [3, 1, 4], 0, True, override_root_with=stale)
