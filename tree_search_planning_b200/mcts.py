"""Host-side mirror of the reference planner interface.

`MCTS(config).run(model, observation, legal_actions, to_play, add_exploration_noise, ...)
-> (root, extra_info)` keeps the surface of muzero-general/self_play.py:303-434 (and of the
stochastic / risk-sensitive siblings, picked from the config flags exactly like muzero.py:232-250),
so `SelfPlay.play_game`, `select_action`, `GameHistory.store_search_statistics` and
`DiagnoseModel` call it unchanged. Underneath, the search runs on the GPU through the C ABI
(engine.py -> libtsp_b200.so); `run_batch` exposes the batched form the engine is built for.
"""
import numpy

from .configs import variant_of, observation_dim
from .engine import Engine

# numpy.random.choice over n tied actions / nf chance outcomes is replaced by
# stream[k] % n; drawing the stream from [0, 240) keeps that unbiased for n in 1..6, 8
_STREAM_MOD = 240

_engine_cache = {}


def _config_key(config, device):
    keys = ("action_space", "players", "observation_shape", "stacked_observations", "encoding_size", "support_size",
            "discount", "pb_c_base", "pb_c_init", "root_exploration_fraction", "mask_absorbing_states",
            "stochastic_dynamics", "risk_sensitive", "n_futures", "fc_representation_layers", "fc_dynamics_layers",
            "fc_reward_layers", "fc_value_layers", "fc_policy_layers", "fc_prior_layers")
    return (device,) + tuple(repr(getattr(config, k, None)) for k in keys)


def _weights_fingerprint(model):
    """When does the engine have to reload the weights? A state dict is identified by its CONTENT (crc32 over every
    array: ~1 ms for these networks) -- never by id(): a dict fetched fresh each step often lands at the address of the
    one just freed, and a dict updated in place keeps its id. A module is identified by (storage address, version
    counter) of every parameter: in-place updates (optimizer steps, load_state_dict / set_weights) bump the version,
    replaced parameters change the address. `MCTS.refresh_weights()` forces a reload regardless."""
    import zlib
    if isinstance(model, dict):
        crc = 0
        for k in sorted(model.keys()):
            v = model[k]
            if hasattr(v, "detach"):
                v = v.detach().cpu().numpy()
            v = numpy.ascontiguousarray(v)
            crc = zlib.crc32(k.encode(), crc)
            crc = zlib.crc32(v.view(numpy.uint8).reshape(-1) if v.size else b"", crc)
        return ("dict", len(model), crc)
    params = list(model.parameters()) if hasattr(model, "parameters") else []
    return ("module", id(model)) + tuple((int(p.data_ptr()), int(getattr(p, "_version", 0))) for p in params)


class ChildView:
    """What consumers read from a node of the search tree: visit_count, prior, reward, value, children.
    Depth 1 comes from the arrays tsp_run returns; `children` of deeper nodes (diagnose_model.py:156-185 recurses
    through them) are built on first access from the device-resident tree (tsp_export_tree), which stays valid until
    the next search on the same engine -- after that they read as {} like an unexpanded node."""

    def __init__(self, visit_count, prior, reward, value, value_is_attribute, tree=None, block=None, root_slot=None, depth=1):
        self.visit_count = int(visit_count)
        self.prior = float(prior)
        self.reward = float(reward)
        self.hidden_state = None
        self.to_play = 0
        self._value = float(value)
        # where this node's own children live: `tree` is a DeviceTree or a callable returning one (or None); `block` the
        # node's block id (-1: not expanded), or None for a child of the root, whose block id the tree knows by `root_slot`
        self._tree, self._block, self._root_slot, self._depth = tree, block, root_slot, depth
        self._children = None
        if value_is_attribute:  # risk-sensitive nodes expose `.value` as an attribute
            self.value = self._value

    def _resolve(self):
        tree = self._tree() if callable(self._tree) else self._tree
        if tree is None:
            return None, -1
        block = self._block
        if block is None:
            block = int(tree.blocks[0]["child"][self._root_slot])
        return tree, block

    @property
    def children(self):
        if self._children is None:
            tree, block = self._resolve()
            self._children = tree.children_of(block, self._depth) if tree is not None and block >= 0 else {}
        return self._children

    @children.setter
    def children(self, value):
        self._children = value

    def expanded(self):
        """self_play.py:516-517: a node is expanded once it has children -- a visited child that the search did not
        expand (an absorbing state under mask_absorbing_states) stays unexpanded like in the reference."""
        tree, block = self._resolve()
        return block >= 0 if tree is not None else self.visit_count > 0

    def value(self):  # shadowed by the instance attribute in the risk-sensitive variant
        return self._value


class DeviceTree:
    """Host copy of one device-resident tree (Engine.export_tree) and the lazily built Node views below the root.
    Block k holds the statistics of the children of expanded node k; in the chance-node variants node kinds alternate
    by depth (even: StateNode, children keyed by action; odd: TransitionNode, children keyed by future z)."""

    def __init__(self, exported, variant, num_actions, n_futures):
        self.info, self.blocks, self.hidden = exported["info"], exported["blocks"], exported["hidden"]
        self.variant, self.A, self.nf = variant, num_actions, max(int(n_futures), 1)

    def children_of(self, block, depth):
        """{action or future: ChildView} of the expanded node that owns `block`; `depth` is that node's depth."""
        b = self.blocks[block]
        transition = self.variant != 0 and (depth & 1) == 1  # a TransitionNode: its children are the futures z
        if block == 0:
            keys = [int(self.info.root_action[j]) for j in range(self.info.n_root)]
        else:
            keys = list(range(self.nf if transition else self.A))
        out = {}
        for slot, key in enumerate(keys):
            visits, child = int(b["visit"][slot]), int(b["child"][slot])
            vsum = float(b["value_sum"][slot])
            prior = float(self.info.root_prior[slot]) if block == 0 else float(b["prior"][slot])
            value = vsum if self.variant == 2 else (vsum / visits if visits > 0 else 0.0)
            # a StateNode below a TransitionNode receives its reward and hidden state when it is expanded itself
            # (StateNode.expand, self_play_stochastic.py:453-469); until then the transition keeps them
            # (child_rewards / child_hidden_states, :550-551) and the node reads reward 0, hidden_state None
            reward = float(b["reward"][slot]) if not (transition and child < 0) else 0.0
            node = ChildView(visits, prior, reward, value, self.variant == 2, tree=self, block=child, depth=depth + 1)
            # hidden state: deterministic -- the expanded child's own block id is its slot; chance-node variants -- slot
            # 1 + ordinal of the transition * nf + z; TransitionNodes carry none
            hs = -1
            if self.variant == 0:
                hs = child
            elif transition and child >= 0:
                hs = 1 + (int(b["aux"]) >> 8) * self.nf + slot
            if self.hidden is not None and 0 <= hs < len(self.hidden):
                node.hidden_state = self.hidden[hs][None]
            if child < 0:
                node.to_play = -1  # self_play.py:508: set by expand
            out[key] = node
        return out


class RootView(ChildView):
    def __init__(self, res, i, legal_actions, variant, hidden_state=None, tree=None):
        super().__init__(res["visit_counts"][i].sum() if "visit_counts" in res else 0, 0.0, 0.0,
                         res["root_value"][i], variant == 2, tree=tree, block=0, depth=0)
        # depth 1 straight from the arrays of tsp_run (no export); deeper levels through `tree` on demand
        self._children = {
            int(a): ChildView(res["visit_counts"][i][a], res["child_prior"][i][a], res["child_reward"][i][a],
                              res["child_value"][i][a], variant == 2, tree=tree, block=None, root_slot=slot, depth=1)
            for slot, a in enumerate(legal_actions)
        }
        self.hidden_state = hidden_state

    def expanded(self):
        return len(self.children) > 0


class BatchResult(dict):
    """Arrays of one batched search + per-root `root` views."""

    def __init__(self, res, legal_lists, variant, engine=None, resumed=False, hidden=None):
        super().__init__(res)
        self._legal = legal_lists
        self._variant = variant
        self._engine = engine
        self._generation = engine.generation if engine is not None else None
        self._resumed = resumed
        self._hidden = hidden  # root hidden states carried over a continued search
        self._trees = {}       # root index -> DeviceTree (exported on demand)

    def device_tree(self, i):
        """Host copy of root i's whole tree (None once another search has run on the engine, or without an engine)."""
        if self._engine is None or self._generation != self._engine.generation:
            return self._trees.get(i)
        if i not in self._trees:
            cfg = self._engine.config
            self._trees[i] = DeviceTree(self._engine.export_tree(i), self._variant, len(cfg.action_space),
                                        getattr(cfg, "n_futures", 1) or 1)
        return self._trees[i]

    def root(self, i):
        hid = self["root_hidden"][i][None] if "root_hidden" in self else None
        if hid is None and self._hidden is not None:
            hid = self._hidden[i][None]
        r = RootView(self, i, self._legal[i], self._variant, hid, tree=(lambda: self.device_tree(i)) if self._engine is not None else None)
        r._batch, r._index = self, i  # lets MCTS.run(..., override_root_with=root) find the device-resident tree
        return r

    def extra_info(self, i):
        info = {"max_tree_depth": int(self["max_tree_depth"][i]),
                # None when the search continued an existing root (self_play.py:331-333)
                "root_predicted_value": None if self._resumed else float(self["root_predicted_value"][i])}
        if self._variant == 0:
            info["n_env_interactions"] = 0  # self_play.py:432
        return info


class MCTS:
    """Drop-in for self_play.MCTS / self_play_stochastic.MCTS / self_play_risk_sensitive.MCTS."""

    def __init__(self, config, device=0, max_roots=1, max_search_multiple=1):
        """`max_search_multiple`: how many searches of config.num_simulations a tree may accumulate through
        override_root_with / run_batch(resume=...) (sizes the device node pools)."""
        self.config = config
        self.device = device
        self.max_roots = max_roots
        self.max_search_multiple = int(max_search_multiple)
        self.variant = variant_of(config)

    # -- engine management ---------------------------------------------------------------
    def _engine(self, model, batch, num_simulations):
        key = _config_key(self.config, self.device)
        ent = _engine_cache.get(key)
        if ent is None or ent["engine"].max_roots < batch or ent["engine"].max_simulations < num_simulations:
            if ent is not None:
                ent["engine"].close()
            eng = Engine(self.config, max(batch, self.max_roots), max(num_simulations, self.config.num_simulations),
                         device=self.device)
            ent = {"engine": eng, "fingerprint": None}
            _engine_cache[key] = ent
        fp = _weights_fingerprint(model)
        if ent["fingerprint"] != fp:
            sd = model if isinstance(model, dict) else model.state_dict()
            ent["engine"].load_state_dict(sd)
            ent["fingerprint"] = fp
        return ent["engine"]

    def refresh_weights(self):
        """Forget which weights the cached engine of this config holds: the next run reloads them from `model`
        (for callers that mutate weights in ways the fingerprint cannot see, e.g. through raw storage)."""
        ent = _engine_cache.get(_config_key(self.config, self.device))
        if ent is not None:
            ent["fingerprint"] = None

    @staticmethod
    def close_engines():
        """Destroy every cached engine (the cache is process-wide and keyed by game config + device, because the
        reference constructs a new MCTS(config) for every move, self_play.py:148)."""
        for ent in _engine_cache.values():
            ent["engine"].close()
        _engine_cache.clear()

    # -- the reference surface ---------------------------------------------------------------
    def run(self, model, observation, legal_actions, to_play, add_exploration_noise, override_root_with=None,
            policy_only=False, uniform_policy=False):
        if self.variant != 0 and (policy_only or uniform_policy):
            raise TypeError("policy_only / uniform_policy exist only in self_play.MCTS.run")
        if override_root_with is not None:
            # self_play.py:331-333: continue the search of an existing root. The tree lives in the engine's pools,
            # so the root must come from the latest search of this engine (no other search in between).
            batch = getattr(override_root_with, "_batch", None)
            if batch is None or batch._engine is None or batch._generation != batch._engine.generation:
                raise ValueError("override_root_with must be the root returned by the latest MCTS.run / run_batch of "
                                 "this engine: its tree is device-resident and the next search reuses the pools")
            if len(batch._legal) != 1:
                raise ValueError("override_root_with continues a whole batch: use run_batch(resume=batch)")
            res = self.run_batch(model, None, add_exploration_noise=add_exploration_noise, policy_only=policy_only,
                                 uniform_policy=uniform_policy, resume=batch)
            return res.root(0), res.extra_info(0)
        assert legal_actions, f"Legal actions should not be an empty array. Got {legal_actions}."
        assert set(legal_actions).issubset(set(self.config.action_space)), \
            "Legal actions should be a subset of the action space."
        obs = numpy.asarray(observation, dtype=numpy.float32).reshape(1, -1)
        res = self.run_batch(model, obs, [list(legal_actions)], add_exploration_noise, policy_only=policy_only,
                             uniform_policy=uniform_policy, want_root_hidden=True)
        return res.root(0), res.extra_info(0)

    def run_batch(self, model, observations, legal_actions=None, add_exploration_noise=True, noise=None, tie=None,
                  chance=None, policy_only=False, uniform_policy=False, num_simulations=None,
                  want_root_hidden=False, resume=None):
        """B independent searches at once. `legal_actions`: list of B ordered lists (None = all).
        Random sources default to numpy.random like the reference (Dirichlet noise is drawn with the
        very call of self_play.py:546); pass `noise` / `tie` / `chance` for reproducible searches.
        `resume`: the BatchResult of the latest search of this engine -- every one of its roots is searched
        further (what MCTS.run does with override_root_with, self_play.py:331-333)."""
        cfg = self.config
        A = len(cfg.action_space)
        S = 0 if policy_only else int(cfg.num_simulations if num_simulations is None else num_simulations)
        if resume is not None:
            if self.variant == 1:  # self_play_stochastic.py:333 fails on a root that already has visits
                raise AssertionError("the stochastic search cannot be continued: the reference asserts "
                                     "num_sims == sum(child.visit_count) (self_play_stochastic.py:333)")
            legal_lists = resume._legal
            B = len(legal_lists)
            legal_actions = None  # the root's children are already in place
            observations = None
        else:
            observations = numpy.ascontiguousarray(observations, numpy.float32).reshape(-1, observation_dim(cfg))
            B = observations.shape[0]
            legal_lists = [list(cfg.action_space)] * B if legal_actions is None else [list(l) for l in legal_actions]
        la = nl = None
        if legal_actions is not None:
            la = numpy.zeros((B, A), numpy.int32)
            nl = numpy.zeros(B, numpy.int32)
            for i, l in enumerate(legal_lists):
                assert l, f"Legal actions should not be an empty array. Got {l}."
                assert set(l).issubset(set(cfg.action_space)), "Legal actions should be a subset of the action space."
                la[i, :len(l)] = l
                nl[i] = len(l)
        if add_exploration_noise and noise is None:
            noise = numpy.zeros((B, A), numpy.float64)
            for i, l in enumerate(legal_lists):
                noise[i, :len(l)] = numpy.random.dirichlet([cfg.root_dirichlet_alpha] * len(l))
        if tie is None:
            tie = numpy.random.randint(0, _STREAM_MOD, size=(B, 64)).astype(numpy.uint8)
        if chance is None and self.variant != 0:
            chance = numpy.random.randint(0, _STREAM_MOD, size=(B, max(64, 8 * S))).astype(numpy.uint8)
        if resume is not None:
            eng = resume._engine
            if resume._generation != eng.generation:
                raise ValueError("resume: another search has run on this engine since; its trees are gone")
            res = eng.run(S, noise=noise, tie=tie, chance=chance, add_exploration_noise=add_exploration_noise,
                          uniform_policy=uniform_policy, resume=True, num_roots=B)
            hidden = resume["root_hidden"] if "root_hidden" in resume else resume._hidden
            return BatchResult(res, legal_lists, self.variant, engine=eng, resumed=True, hidden=hidden)
        # room for continued searches: pools sized for a multiple of the configured simulations
        eng = self._engine(model, B, S * self.max_search_multiple)
        res = eng.run(S, observations=observations, legal_actions=la, num_legal=nl, noise=noise, tie=tie,
                      chance=chance, add_exploration_noise=add_exploration_noise, uniform_policy=uniform_policy,
                      want_root_hidden=want_root_hidden)
        return BatchResult(res, legal_lists, self.variant, engine=eng)
