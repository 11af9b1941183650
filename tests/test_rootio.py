"""Root I/O formats either side of the search: GameHistory.get_stacked_observations,
SelfPlay.select_action, GameHistory.store_search_statistics.

CPU: the numpy restatement (oracle/rootio.py) against tests/golden/rootio.npz, written by the LIVE reference
functions (oracle/make_golden_rootio.py).  GPU (-m gpu): the CUDA kernels, through the C ABI, against the same
golden file and against the restatement at batch scale; the batched self-play driver against a step-by-step
host loop built from the restatement."""
import os

import numpy
import pytest

from oracle import rootio

G = numpy.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rootio.npz"))


def test_select_action_restatement_matches_reference():
    for i in range(len(G["sa_action"])):
        n = int(G["sa_nlegal"][i])
        la = [int(a) for a in G["sa_legal"][i, :n]]
        v = [int(G["sa_visits"][i, a]) for a in la]
        got = rootio.select_action(v, la, float(G["sa_temperature"][i]), float(G["sa_uniform"][i]))
        assert got == int(G["sa_action"][i]), i


def test_child_visits_restatement_matches_reference():
    for i in range(len(G["sa_action"])):
        n = int(G["sa_nlegal"][i])
        d = {int(a): int(G["sa_visits"][i, a]) for a in G["sa_legal"][i, :n]}
        assert rootio.child_visits(d, list(range(5))) == list(G["ss_child_visits"][i])


def _stack_cases():
    for ci, (C, H, W, so, L) in enumerate(G["so_cases"]):
        frames, actions = G["so_case%d_frames" % ci], G["so_case%d_actions" % ci]
        for index in list(range(L)) + [-1]:
            yield ci, (int(C), int(H), int(W), int(so), int(L)), frames, actions, index, G["so_case%d_index%d" % (ci, index)]


def test_stacked_observations_restatement_matches_reference():
    for ci, (C, H, W, so, L), frames, actions, index, want in _stack_cases():
        got = rootio.stacked_observations(list(frames), list(actions), index, so)
        assert got.shape == want.shape and (numpy.asarray(got, numpy.float32) == want).all(), (ci, index)


# ------------------------------------------------------------------------------------ GPU ----
def _engine(game="roundabout_env_ttc_flat", B=512, **over):
    import tree_search_planning_b200 as tsp
    from tree_search_planning_b200.synthetic import random_state_dict
    cfg = tsp.get_config(game, **over)
    return cfg, tsp.Engine(cfg, max_roots=B, state_dict=random_state_dict(cfg, 3))


@pytest.mark.gpu
def test_select_action_and_statistics_kernels_match_reference_golden():
    cfg, eng = _engine()
    for t in (0.0, 0.25, 0.5, 1.0):
        m = G["sa_temperature"] == t
        got = eng.select_action(G["sa_visits"][m], t, G["sa_uniform"][m], G["sa_legal"][m], G["sa_nlegal"][m])
        assert (got == G["sa_action"][m]).all(), t
    got = eng.search_statistics(G["sa_visits"])
    assert (got == G["ss_child_visits"]).all()  # bit-exact float64
    eng.close()


@pytest.mark.gpu
def test_select_action_kernel_at_batch_scale():
    cfg, eng = _engine(B=65536)
    rng = numpy.random.default_rng(5)
    B, A = 65536, 5
    visits = rng.multinomial(50, rng.dirichlet([0.3] * A, size=B)).astype(numpy.int32)
    u = rng.random(B)
    for t in (0.0, 0.25, 1.0, float("inf")):
        got = eng.select_action(visits, t, u)
        want = numpy.array([rootio.select_action(visits[b], list(range(A)), t, u[b]) for b in range(0, B, 37)])
        assert (got[::37] == want).all(), t
    # size-independent properties: the arg-max never has fewer visits than any other action; sampled actions
    # always have visits > 0; frequencies follow the visit distribution
    best = eng.select_action(visits, 0.0, u)
    assert (visits[numpy.arange(B), best] == visits.max(1)).all()
    samp = eng.select_action(visits, 1.0, u)
    assert (visits[numpy.arange(B), samp] > 0).all()
    freq = numpy.bincount(samp, minlength=A) / B
    assert numpy.abs(freq - (visits / 50.0).mean(0)).max() < 0.01
    cv = eng.search_statistics(visits)
    assert numpy.allclose(cv.sum(1), 1.0) and (cv == visits / 50.0).all()
    eng.close()


@pytest.mark.gpu
def test_stack_observations_kernel_matches_reference_golden():
    import torch
    import tree_search_planning_b200 as tsp
    from tree_search_planning_b200.synthetic import random_state_dict
    dev = torch.device("cuda", 0)
    for ci, (C, H, W, so, L) in enumerate(G["so_cases"]):
        C, H, W, so, L = int(C), int(H), int(W), int(so), int(L)
        cfg = tsp.get_config("roundabout_env_ttc_flat", observation_shape=(C, H, W), stacked_observations=so)
        eng = tsp.Engine(cfg, max_roots=8, state_dict=random_state_dict(cfg, 1))
        frames, actions = G["so_case%d_frames" % ci], G["so_case%d_actions" % ci]
        idxs = list(range(L)) + [-1]
        fi = numpy.full((len(idxs), so + 1), -1, numpy.int32)
        av = numpy.zeros((len(idxs), max(so, 1)), numpy.float32)
        for r, index in enumerate(idxs):
            index = index % L
            fi[r, 0] = index
            for i in range(1, so + 1):
                if index - i >= 0:
                    fi[r, i] = index - i
                    av[r, i - 1] = actions[index - i + 1]
        pool = torch.from_numpy(frames.reshape(L, -1)).to(dev)
        out = torch.zeros((len(idxs), tsp.observation_dim(cfg)), dtype=torch.float32, device=dev)
        eng.stack_observations(pool, torch.from_numpy(fi).to(dev), torch.from_numpy(av).to(dev), out, so, C * H * W, H * W)
        eng.synchronize()
        got = out.cpu().numpy()
        for r, index in enumerate(idxs):
            want = G["so_case%d_index%d" % (ci, index)].reshape(-1)
            assert (got[r] == want).all(), (ci, index)
        eng.close()


class _ToyGame:
    """Deterministic stand-in for the reference's Game protocol (gym / highway-env are absent): the observation is
    a fixed function of (game id, time, last action); legal actions vary over time; ends after `length` moves."""

    def __init__(self, gid, frame, length):
        self.gid, self.frame, self.length, self.t, self.last = gid, frame, length, 0, 0

    def _obs(self):
        k = numpy.arange(self.frame, dtype=numpy.float32)
        return (numpy.sin(0.37 * k + self.gid + 0.61 * self.t + 1.3 * self.last) * 0.5 + 0.5).reshape(1, 1, self.frame)

    def reset(self):
        self.t, self.last = 0, 0
        return self._obs()

    def step(self, action):
        self.t += 1
        self.last = int(action)
        return self._obs(), float(action == 2), self.t >= self.length

    def legal_actions(self):
        return [0, 1, 2, 3, 4] if (self.t + self.gid) % 3 else [4, 2, 0]

    def to_play(self):
        return 0


@pytest.mark.gpu
def test_batched_self_play_driver_against_host_loop():
    """Every step of the batched driver == the reference's per-game formats applied to the same search results:
    stacked observations (bit-exact vs the restatement of get_stacked_observations), actions (select_action on the
    driver's visit counts and uniforms), policy targets (store_search_statistics)."""
    import tree_search_planning_b200 as tsp
    from tree_search_planning_b200.selfplay import BatchedSelfPlay
    from tree_search_planning_b200.synthetic import random_state_dict
    cfg = tsp.get_config("roundabout_env_ttc_flat", num_simulations=20)
    B = 48
    games = [_ToyGame(g, 155, 3 + g % 4) for g in range(B)]
    numpy.random.seed(11)
    drv = BatchedSelfPlay(cfg, random_state_dict(cfg, 9), games, temperature=1.0)
    moves = 0
    while not drv.done.all():
        live = ~drv.done
        acts = drv.step()
        moves += 1
        obs = drv.obs.cpu().numpy()
        visits = drv.d_visits.cpu().numpy()
        u = drv.d_uniform.cpu().numpy()
        for b in numpy.nonzero(live)[0]:
            h = drv.histories[b]
            # the observation the search saw: history without the frame appended by this step
            want = rootio.stacked_observations(h.observation_history[:-1], h.action_history[:-1], -1, cfg.stacked_observations)
            assert (obs[b] == numpy.asarray(want, numpy.float32).reshape(-1)).all(), (moves, b)
            la = games[b].legal_actions.__func__(types_ns(games[b], moves - 1))
            v = [int(visits[b, a]) for a in la]
            assert sum(v) == cfg.num_simulations and visits[b].sum() == cfg.num_simulations
            assert int(acts[b]) == rootio.select_action(v, la, 1.0, float(u[b]))
            assert h.child_visits[-1] == rootio.child_visits({a: int(visits[b, a]) for a in la}, list(cfg.action_space))
            assert h.action_history[-1] == int(acts[b])
    assert moves == 6 and all(len(h.root_values) == g.length for h, g in zip(drv.histories, games))
    drv.close()


def types_ns(game, t):
    """The game as it was at move t (legal_actions depends on t and the game id only)."""
    import types
    return types.SimpleNamespace(t=t, gid=game.gid)


@pytest.mark.gpu
def test_reanalyse_values_and_hot_weight_refresh():
    """Reanalyse (replay_buffer.py:386-406) through the batched root kernel, and a weight refresh on a live engine
    (what SelfPlay.continuous_self_play does with set_weights every checkpoint_interval, self_play.py:43)."""
    import tree_search_planning_b200 as tsp
    from tree_search_planning_b200.selfplay import GameHistory, reanalyse_values
    from tree_search_planning_b200.synthetic import random_state_dict, synthetic_inputs
    from oracle.oracle import Oracle
    cfg = tsp.get_config("roundabout_env_ttc_flat", num_simulations=25)
    sd_a, sd_b = random_state_dict(cfg, 1), random_state_dict(cfg, 2)
    eng = tsp.Engine(cfg, max_roots=256, state_dict=sd_a)
    rng = numpy.random.default_rng(3)
    hists = []
    for g in range(5):
        hgame = GameHistory()
        L = 3 + 2 * g
        for t in range(L + 1):
            hgame.observation_history.append(rng.random((1, 1, 155), dtype=numpy.float32))
            hgame.action_history.append(int(rng.integers(0, 5)))
        hgame.root_values = [0.0] * L
        hists.append(hgame)
    got = reanalyse_values(eng, cfg, hists)
    orc = Oracle(cfg, sd_a)
    for hgame, v in zip(hists, got):
        stacked = numpy.stack([rootio.stacked_observations(hgame.observation_history, hgame.action_history, i,
                                                           cfg.stacked_observations).reshape(-1)
                               for i in range(len(hgame.root_values))])
        want = orc.initial_inference(stacked)["value"]
        assert v.shape == want.shape
        assert (numpy.abs(v - want) <= 2e-4 * numpy.maximum(1, numpy.abs(want))).all()
        assert (numpy.abs(v - want) <= 1e-5 * numpy.maximum(1, numpy.abs(want))).mean() >= 0.9
    # hot refresh: the same engine with new weights == a fresh engine with those weights, bit for bit
    obs, kw = synthetic_inputs(cfg, 256, 25, seed=4)
    first = eng.run(25, observations=obs, **kw)
    eng.load_state_dict(sd_b)
    refreshed = eng.run(25, observations=obs, **kw)
    fresh_eng = tsp.Engine(cfg, max_roots=256, state_dict=sd_b)
    fresh = fresh_eng.run(25, observations=obs, **kw)
    for k in ("visit_counts", "root_value", "child_value", "child_prior", "child_reward", "max_tree_depth"):
        assert (refreshed[k] == fresh[k]).all(), k
    assert not (first["root_value"] == refreshed["root_value"]).all()
    eng.close()
    fresh_eng.close()
