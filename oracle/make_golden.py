"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/*.npz from the LIVE reference.

Run in the build container (needs /root/reference):  python -m oracle.make_golden

For every BASELINE config it stores: the seeded random-init reference weights (the repo
checkpoints are Git-LFS stubs), the external inputs (observations, ordered legal-action
lists, Dirichlet noise, tie stream, chance stream) and everything the UNMODIFIED reference
`MCTS.run` produced for them (root visit counts, values, priors, rewards, depth, node count)
plus the per-evaluation network records in call order, and batched network outputs of
`initial_inference` / `recurrent_inference` / `dynamics` / `prediction` for network parity.
"""
import os
import sys

import numpy

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from tree_search_planning_b200.configs import get_config, variant_of, observation_dim  # noqa: E402
from oracle import reference_live as RL  # noqa: E402

OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# (fixture name, game config, S, number of roots, options, seed) -- the seed is part of the case, not its list position
CASES = [
    ("highway_s25", "highway_env_ttc_flat", 25, 16, {}, 0),
    ("roundabout_s50", "roundabout_env_ttc_flat", 50, 16, {}, 1),
    ("roundabout_s50_legal", "roundabout_env_ttc_flat", 50, 12, {"legal": True}, 2),
    ("roundabout_s25_uniform", "roundabout_env_ttc_flat", 25, 8, {"uniform_policy": True}, 3),
    ("roundabout_s25_nonoise", "roundabout_env_ttc_flat", 25, 8, {"no_noise": True}, 4),
    ("roundabout_s25_mask", "roundabout_env_ttc_flat", 25, 8, {"mask": True}, 5),
    ("stochastic_s25", "roundabout_env_ttc_flat_stochastic_concat", 25, 16, {}, 6),
    ("stochastic_s25_legal", "roundabout_env_ttc_flat_stochastic_concat", 25, 8, {"legal": True}, 7),
    ("risk_s25", "roundabout_env_ttc_flat_risk_sensitive", 25, 16, {}, 8),
    ("risk_s25_legal", "roundabout_env_ttc_flat_risk_sensitive", 25, 8, {"legal": True}, 9),
    ("cross_merge_s200", "cross_merge_env", 200, 4, {}, 10),
    # two-player backup / ucb branch (self_play.py:469-472, 492-500); to_play alternates over the roots
    ("roundabout_s25_twoplayer", "roundabout_env_ttc_flat", 25, 12, {"players": [0, 1]}, 11),
    # the search continued twice with override_root_with=root (self_play.py:331-333): 3 x S simulations
    ("roundabout_s25_override", "roundabout_env_ttc_flat", 25, 8, {"phases": 3}, 12),
    # (the stochastic search cannot be continued: its assert at self_play_stochastic.py:333 fails on a root
    #  that already has visits; the risk-sensitive one has no such assert)
    ("roundabout_s25_override_legal", "roundabout_env_ttc_flat", 25, 6, {"phases": 2, "legal": True, "mask": "adaptive"}, 13),
    ("risk_s25_override", "roundabout_env_ttc_flat_risk_sensitive", 25, 6, {"phases": 2, "legal": True}, 14),
    ("roundabout_s25_twoplayer_mask", "roundabout_env_ttc_flat", 25, 8, {"players": [0, 1], "mask": "adaptive", "legal": True}, 12),
    # models.MuZeroStochastic: one dynamics MLP per future (models.py:298-462), 4 and 2 futures as in the game files
    ("stochastic_sep_s25", "roundabout_env_ttc_flat_stochastic", 25, 8, {}, 15),
    ("stochastic_sep_s25_legal", "highway_env_ttc_flat_stochastic", 25, 8, {"legal": True}, 16),
]


def make_case(name, game, S, N, opt, seed):
    cfg = get_config(game, num_simulations=S)
    if opt.get("mask"):
        cfg = cfg.replace(mask_absorbing_states=True)
    if opt.get("players"):
        cfg = cfg.replace(players=list(opt["players"]))
    ref = RL.load_reference()
    torch = ref.torch
    model = RL.build_model(cfg, seed)
    if opt.get("mask"):
        # random-init terminal logits sit just below 0; lift the bias so that the absorbing
        # branch (self_play.py:415-417) actually fires in some evaluations
        sd = model.state_dict()
        key = [k for k in sd if k.startswith("dynamics_terminal_network") and k.endswith("bias")][-1]
        lift = 0.07
        if opt.get("mask") == "adaptive":  # shift so that about a third of the evaluations are terminal
            with torch.no_grad():
                o = torch.rand(512, cfg.stacked_observations * 2 + 1, 1, cfg.observation_shape[2])
                hs = model.initial_inference(o)[5]
                acts = torch.randint(0, len(cfg.action_space), (512, 1))
                term = model.recurrent_inference(hs, acts)[2].reshape(-1)  # depth-1 evaluations
            lift = -float(torch.quantile(term, 0.65))
        sd[key] = sd[key] + lift
        model.load_state_dict(sd)
    A = len(cfg.action_space)
    variant = variant_of(cfg)
    D = observation_dim(cfg)
    rng = numpy.random.default_rng(1000 + seed)
    max_rec = S if variant == 0 else S * (A + 2) + A + 2
    K = 8 * S
    T = 64
    P = int(opt.get("phases", 1))
    obs = rng.random((N, D), dtype=numpy.float32)
    legal = numpy.zeros((N, A), numpy.int32)
    n_legal = numpy.zeros(N, numpy.int32)
    noise = numpy.zeros((N, A), numpy.float64)
    tie = rng.integers(0, 256, size=(N, T)).astype(numpy.uint8)
    chance = rng.integers(0, 256, size=(N, K)).astype(numpy.uint8)
    if P > 1:  # per-phase streams [P][N][...]; phase 0 keeps the arrays above
        noise = numpy.zeros((P, N, A), numpy.float64)
        tie = numpy.concatenate([tie[None], rng.integers(0, 256, size=(P - 1, N, T)).astype(numpy.uint8)])
        chance = numpy.concatenate([chance[None], rng.integers(0, 256, size=(P - 1, N, K)).astype(numpy.uint8)])
    out = dict(
        visit_counts=numpy.zeros((N, A), numpy.int32), root_value=numpy.zeros(N),
        child_value=numpy.zeros((N, A)), child_prior=numpy.zeros((N, A)), child_reward=numpy.zeros((N, A)),
        max_tree_depth=numpy.zeros(N, numpy.int32), root_predicted_value=numpy.zeros(N, numpy.float32),
        n_nodes=numpy.zeros(N, numpy.int32), n_records=numpy.zeros(N, numpy.int32),
        n_tie_draws=numpy.zeros(N, numpy.int32), n_chance_draws=numpy.zeros(N, numpy.int32),
        root_record=numpy.zeros((N, RL.REC_WIDTH), numpy.float32),
        records=numpy.zeros((N, max_rec, RL.REC_WIDTH), numpy.float32) if int(opt.get("phases", 1)) == 1
        else numpy.zeros((int(opt["phases"]), N, max_rec, RL.REC_WIDTH), numpy.float32),
        root_hidden=numpy.zeros((N, cfg.encoding_size), numpy.float32),
    )
    add_noise = not opt.get("no_noise", False)
    uniform = bool(opt.get("uniform_policy", False))
    to_play = numpy.array([i % len(cfg.players) for i in range(N)], numpy.int32)
    for i in range(N):
        if opt.get("legal"):
            k = int(rng.integers(1, A + 1))
            la = [int(a) for a in rng.permutation(A)[:k]]
        else:
            la = list(range(A))
        n_legal[i] = len(la)
        legal[i, :len(la)] = la
        nz = rng.dirichlet([cfg.root_dirichlet_alpha] * len(la)) if len(la) > 1 else numpy.ones(1)
        conts = []
        if P == 1:
            noise[i, :len(la)] = nz
            tie_i, chance_i = tie[i], chance[i]
        else:
            noise[0, i, :len(la)] = nz
            tie_i, chance_i = tie[0, i], chance[0, i]
            for ph in range(1, P):
                nz2 = rng.dirichlet([cfg.root_dirichlet_alpha] * len(la)) if len(la) > 1 else numpy.ones(1)
                noise[ph, i, :len(la)] = nz2
                conts.append(dict(noise=nz2, tie=tie[ph, i], chance=chance[ph, i]))
        r = RL.run_reference(cfg, model, obs[i].reshape(cfg.stacked_observations * 2 + 1, 1, -1), legal_actions=la,
                             noise=nz, tie=tie_i, chance=chance_i, add_exploration_noise=add_noise,
                             uniform_policy=uniform, to_play=int(to_play[i]), continuations=conts)
        for k2 in ("visit_counts", "root_value", "child_value", "child_prior", "child_reward", "max_tree_depth",
                   "root_predicted_value", "n_nodes", "n_tie_draws", "n_chance_draws", "root_record",
                   "root_hidden"):
            out[k2][i] = r[k2]
        if P == 1:
            nr = len(r["records"])
            assert nr <= max_rec
            out["n_records"][i] = nr
            out["records"][i, :nr] = r["records"]
        else:
            for ph in range(P):
                nr = len(r["records"][ph])
                assert nr <= max_rec
                out["records"][ph, i, :nr] = r["records"][ph]
            out["n_records"][i] = nr  # of the last search
        assert int(r["visit_counts"].sum()) == S * P and r["root_visit_count"] == S * P

    # ---- batched network goldens (torch fp32, batch M) ----
    M = 48
    net = {}
    with torch.no_grad():
        o = torch.tensor(obs[:min(N, M)]).reshape(-1, cfg.stacked_observations * 2 + 1, 1, cfg.observation_shape[2])
        value, reward, _t, logits, _r, hidden = model.initial_inference(o)
        net["init_value_logits"] = value.numpy()
        net["init_value"] = ref.models.support_to_scalar(value, cfg.support_size).numpy()[:, 0]
        net["init_policy_logits"] = logits.numpy()
        net["init_hidden"] = hidden.numpy()
        h = torch.tensor(rng.random((M, cfg.encoding_size), dtype=numpy.float32))
        h[: hidden.shape[0]] = hidden[:M]
        act = torch.tensor(rng.integers(0, A, size=(M, 1)))
        net["rec_in_hidden"] = h.numpy()
        net["rec_in_action"] = act.numpy()[:, 0].astype(numpy.int32)
        if variant == 0:
            value, reward, terminal, logits, _recon, nh = model.recurrent_inference(h, act)
            net["rec_value_logits"] = value.numpy()
            net["rec_reward_logits"] = reward.numpy()
            net["rec_value"] = ref.models.support_to_scalar(value, cfg.support_size).numpy()[:, 0]
            net["rec_reward"] = ref.models.support_to_scalar(reward, cfg.support_size).numpy()[:, 0]
            net["rec_terminal"] = terminal.numpy()
            net["rec_policy_logits"] = logits.numpy()
            net["rec_hidden"] = nh.numpy()
        else:
            nf = cfg.n_futures
            tl = numpy.zeros((M, nf), numpy.float32)
            hs = numpy.zeros((M, nf, cfg.encoding_size), numpy.float32)
            rw = numpy.zeros((M, nf), numpy.float32)
            for i in range(M):  # the reference asserts batch 1 here (models.py:538-541)
                tlog, nhs, rws = model.dynamics(h[i:i + 1], act[i:i + 1])
                tl[i] = tlog[:, 0].numpy()
                hs[i] = nhs[:, 0].numpy()
                for z in range(nf):
                    rw[i, z] = ref.models.support_to_scalar(rws[z], cfg.support_size).item()
            net["dyn_transition_logits"] = tl
            net["dyn_hidden"] = hs
            net["dyn_reward"] = rw
            logits, value = model.prediction(h)
            net["pred_policy_logits"] = logits.numpy()
            net["pred_value"] = ref.models.support_to_scalar(value, cfg.support_size).numpy()[:, 0]

    weights = {"w:" + k: v for k, v in RL.state_dict_numpy(model).items()
               if not k.startswith("reconstruction_network") and not k.startswith("transition_posterior")}
    meta = dict(game=game, S=S, N=N, seed=seed, add_noise=add_noise, uniform_policy=uniform,
                mask_absorbing_states=bool(opt.get("mask", False)), players=list(cfg.players), phases=P)
    path = os.path.join(OUT_DIR, name + ".npz")
    numpy.savez_compressed(
        path, obs=obs, legal=legal, n_legal=n_legal, noise=noise, tie=tie, chance=chance, to_play=to_play,
        meta=numpy.array(repr(meta)), **{"out:" + k: v for k, v in out.items()},
        **{"net:" + k: v for k, v in net.items()}, **weights)
    return path, out


def main():
    os.makedirs(OUT_DIR, exist_ok=True)
    only = set(sys.argv[1:])  # optional fixture names: the committed fixtures are not rewritten needlessly
    for name, game, S, N, opt, seed in CASES:
        if only and name not in only:
            continue
        path, out = make_case(name, game, S, N, opt, seed=seed)
        print("%-26s %8.1f KB  visits[0]=%s depth<=%d records<=%d" % (
            name, os.path.getsize(path) / 1024.0, out["visit_counts"][0], out["max_tree_depth"].max(),
            out["n_records"].max()))


if __name__ == "__main__":
    main()
