"""TEST INFRASTRUCTURE ONLY -- live-reference harness (runs only where /root/reference exists).

Imports the UNMODIFIED reference planner (muzero-general/self_play.py,
self_play_stochastic.py, self_play_risk_sensitive.py, models.py) read-only from
/root/reference, with the two absent third-party modules (`ray`, `gym`) stubbed in
sys.modules, and runs `MCTS(config).run(...)` with the three random sources of the hot
path supplied from outside:

  * root Dirichlet noise          numpy.random.dirichlet   (self_play.py:546)
  * tie-breaking in select_child  numpy.random.choice      (self_play.py:444-450)
  * chance-outcome sampling       numpy.random.choice      (self_play_stochastic.py:530-532)

The externalised semantics (shared by the C oracle and the CUDA engine):
  - dirichlet(...)            -> the caller's noise vector, verbatim;
  - choice(lst), len(lst)==1  -> lst[0], consumes nothing (numpy does not touch its RNG
                                 state for a singleton either);
  - choice(tied actions)      -> lst[tie[k_t % len(tie)] % len(lst)], k_t += 1
                                 (no stream: lowest index, i.e. tie value 0);
  - choice(chance outcomes)   -> lst[chance[k_c % len(chance)] % len(lst)], k_c += 1.

It also records every network evaluation in call order as a fixed-width "evaluation
record" (see REC_* below) so that tree arithmetic can be compared with network outputs
fed identically.

Nothing in the product package imports this file; tests/ and oracle/make_golden.py do.
"""
import os
import sys
import types
import contextlib

import numpy

REFERENCE_DIR = "/root/reference/muzero-general"

# evaluation record layout (float32[REC_WIDTH]); same in oracle/tsp_oracle.c and include/tsp.h
REC_WIDTH = 16
REC_KIND_STATE = 1.0       # deterministic leaf / stochastic StateNode leaf
REC_KIND_TRANSITION = 2.0  # stochastic TransitionNode leaf
# state record:      [0]=kind [1]=value [2]=reward [3]=terminal(raw logit) [4..4+A) priors
# transition record: [0]=kind [1..1+nf) P(z) (softmax of prior logits) [1+nf..1+2nf) rewards


def reference_available():
    return os.path.isdir(REFERENCE_DIR)


_mods = None


def load_reference():
    """Import the reference modules once; returns a namespace with
    models, self_play, self_play_stochastic, self_play_risk_sensitive."""
    global _mods
    if _mods is not None:
        return _mods
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_DIR)
    os.environ.setdefault("CUDA_VISIBLE_DEVICES", "")
    if "ray" not in sys.modules:
        ray = types.ModuleType("ray")

        def remote(*a, **k):
            if a and callable(a[0]) and not k:
                return a[0]
            return lambda c: c

        ray.remote = remote
        sys.modules["ray"] = ray
    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")
        gym.Env = type("Env", (), {})
        sys.modules["gym"] = gym
    sys.path.insert(0, REFERENCE_DIR)
    try:
        import torch

        torch.set_num_threads(1)
        import models
        import self_play
        import self_play_stochastic
        import self_play_risk_sensitive
    finally:
        sys.path.remove(REFERENCE_DIR)
    _mods = types.SimpleNamespace(
        torch=torch,
        models=models,
        self_play=self_play,
        self_play_stochastic=self_play_stochastic,
        self_play_risk_sensitive=self_play_risk_sensitive,
    )
    return _mods


class ExternalStreams:
    """Replacement for numpy.random.{dirichlet,choice} reading caller-supplied streams."""

    def __init__(self, noise=None, tie=None, chance=None):
        self.noise = None if noise is None else numpy.asarray(noise, dtype=numpy.float64)
        self.tie = None if tie is None else numpy.asarray(tie).astype(numpy.int64)
        self.chance = None if chance is None else numpy.asarray(chance).astype(numpy.int64)
        self.k_tie = 0
        self.k_chance = 0
        self.n_tie_events = 0

    def dirichlet(self, alpha, size=None):
        assert self.noise is not None, "dirichlet noise requested but none supplied"
        assert len(alpha) == len(self.noise), (len(alpha), len(self.noise))
        return self.noise.copy()

    def choice(self, a, *args, **kw):
        lst = list(a)
        if len(lst) == 1:
            return lst[0]
        caller = sys._getframe(1)
        owner = caller.f_locals.get("self", None)
        is_chance = type(owner).__name__ == "TransitionNode"
        if is_chance:
            if self.chance is None or len(self.chance) == 0:
                v = 0
            else:
                v = int(self.chance[self.k_chance % len(self.chance)])
            self.k_chance += 1
        else:
            self.n_tie_events += 1
            if self.tie is None or len(self.tie) == 0:
                v = 0
            else:
                v = int(self.tie[self.k_tie % len(self.tie)])
            self.k_tie += 1
        return lst[v % len(lst)]


@contextlib.contextmanager
def patched_rng(streams):
    old_d, old_c = numpy.random.dirichlet, numpy.random.choice
    numpy.random.dirichlet = streams.dirichlet
    numpy.random.choice = streams.choice
    try:
        yield streams
    finally:
        numpy.random.dirichlet, numpy.random.choice = old_d, old_c


def build_model(config, seed):
    """Seeded random-init reference network (the repo checkpoints are Git-LFS stubs)."""
    ref = load_reference()
    ref.torch.manual_seed(seed)
    model = ref.models.MuZeroNetwork(config)
    model.eval()
    return model


def state_dict_numpy(model):
    return {k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()}


class _Recorder:
    """Wraps the model's inference entry points and appends one evaluation record per
    call, computing the scalars/priors with the very expressions MCTS.run uses
    (support_to_scalar(...).item(), torch.softmax(torch.tensor([...]), dim=0).tolist())."""

    def __init__(self, ref, model, config, uniform_policy=False):
        self.ref, self.model, self.config = ref, model, config
        self.uniform = uniform_policy
        self.records = []
        self.root_record = None
        self.root_hidden = None
        self._in_initial = False
        self.A = len(config.action_space)
        self.legal = None

    def _priors(self, logits, actions):
        torch = self.ref.torch
        lg = logits.clone()
        if self.uniform:
            lg.zero_()
        return torch.softmax(torch.tensor([lg[0][a] for a in actions]), dim=0).tolist()

    def _scalar(self, logits):
        return self.ref.models.support_to_scalar(logits, self.config.support_size).item()

    def install(self):
        m = self.model
        self._orig = {}
        for name in ("initial_inference", "recurrent_inference", "prediction", "dynamics"):
            if hasattr(m, name):
                self._orig[name] = getattr(m, name)
        rec = self

        def initial_inference(obs):
            rec._in_initial = True
            try:
                out = rec._orig["initial_inference"](obs)
            finally:
                rec._in_initial = False
            value, reward, _t, logits, _r, hidden = out
            r = numpy.zeros(REC_WIDTH, numpy.float32)
            r[0] = REC_KIND_STATE
            r[1] = rec._scalar(value)
            r[2] = rec._scalar(reward)
            pri = rec._priors(logits, rec.legal)
            r[4:4 + len(pri)] = pri
            rec.root_record = r
            rec.root_hidden = hidden.detach().numpy().copy().reshape(-1)
            rec.root_logits = logits.detach().numpy().copy().reshape(-1)
            return out

        def recurrent_inference(h, a):
            out = rec._orig["recurrent_inference"](h, a)
            value, reward, terminal, logits, _recon, _hs = out
            r = numpy.zeros(REC_WIDTH, numpy.float32)
            r[0] = REC_KIND_STATE
            r[1] = rec._scalar(value)
            r[2] = rec._scalar(reward)
            r[3] = terminal.item()
            r[4:4 + rec.A] = rec._priors(logits, rec.config.action_space)
            rec.records.append(r)
            return out

        def prediction(h):
            out = rec._orig["prediction"](h)
            if rec._in_initial:
                return out
            logits, value = out
            r = numpy.zeros(REC_WIDTH, numpy.float32)
            r[0] = REC_KIND_STATE
            r[1] = rec._scalar(value)
            r[4:4 + rec.A] = rec._priors(logits, rec.config.action_space)
            rec.records.append(r)
            return out

        def dynamics(h, a):
            out = rec._orig["dynamics"](h, a)
            tlogits, _hs, rewards = out
            nf = tlogits.shape[0]
            r = numpy.zeros(REC_WIDTH, numpy.float32)
            r[0] = REC_KIND_TRANSITION
            r[1:1 + nf] = tlogits[:, 0].softmax(dim=0).tolist()
            for z in range(nf):
                r[1 + nf + z] = rec._scalar(rewards[z])
            rec.records.append(r)
            return out

        m.initial_inference = initial_inference
        if "recurrent_inference" in self._orig and not getattr(self.config, "stochastic_dynamics", False):
            m.recurrent_inference = recurrent_inference
        if getattr(self.config, "stochastic_dynamics", False):
            m.prediction = prediction
            m.dynamics = dynamics

    def uninstall(self):
        for name in ("initial_inference", "recurrent_inference", "prediction", "dynamics"):
            if name in self.model.__dict__:
                del self.model.__dict__[name]


def _node_value(node):
    v = node.value
    return float(v() if callable(v) else v)


def _tree_size(node):
    n, stack = 0, [node]
    while stack:
        x = stack.pop()
        n += 1
        stack.extend(x.children.values())
    return n


def run_reference(config, model, observation, legal_actions=None, noise=None, tie=None,
                  chance=None, add_exploration_noise=True, to_play=0, uniform_policy=False,
                  policy_only=False, record=True, continuations=(), keep_root=False):
    """Run the reference MCTS.run once; returns a dict of everything a consumer reads
    from `root` / `extra_info` plus the evaluation records.

    `continuations`: sequence of dict(noise=, tie=, chance=); after the first search the reference is called
    again with override_root_with=root (self_play.py:331-333) once per entry, with those random streams. The
    returned statistics are those of the final tree, `records` is then a list with one array per search."""
    from tree_search_planning_b200.configs import variant_of

    ref = load_reference()
    torch = ref.torch
    variant = variant_of(config)
    module = (ref.self_play, ref.self_play_stochastic, ref.self_play_risk_sensitive)[variant]
    A = len(config.action_space)
    if legal_actions is None:
        legal_actions = list(config.action_space)
    legal_actions = [int(a) for a in legal_actions]
    streams = ExternalStreams(noise=noise, tie=tie, chance=chance)
    rec = _Recorder(ref, model, config, uniform_policy=uniform_policy)
    rec.legal = legal_actions
    if record:
        rec.install()
    try:
        with torch.no_grad(), patched_rng(streams):
            kwargs = {}
            if variant == 0:
                kwargs = dict(policy_only=policy_only, uniform_policy=uniform_policy)
            root, extra = module.MCTS(config).run(
                model, numpy.asarray(observation), legal_actions, to_play,
                add_exploration_noise, **kwargs)
        phase_records = [list(rec.records)]
        for cont in continuations:
            rec.records = []
            streams = ExternalStreams(noise=cont.get("noise"), tie=cont.get("tie"), chance=cont.get("chance"))
            with torch.no_grad(), patched_rng(streams):
                root, extra = module.MCTS(config).run(
                    model, numpy.asarray(observation), legal_actions, to_play,
                    add_exploration_noise, override_root_with=root, **kwargs)
            assert extra["root_predicted_value"] is None
            extra = dict(extra, root_predicted_value=float("nan"))
            phase_records.append(list(rec.records))
    finally:
        if record:
            rec.uninstall()
    visits = numpy.zeros(A, numpy.int32)
    prior = numpy.zeros(A, numpy.float64)
    reward = numpy.zeros(A, numpy.float64)
    q = numpy.zeros(A, numpy.float64)
    for a, ch in root.children.items():
        visits[a] = ch.visit_count
        prior[a] = float(ch.prior)
        reward[a] = float(ch.reward)
        q[a] = _node_value(ch)
    out = dict(
        visit_counts=visits,
        child_prior=prior,
        child_reward=reward,
        child_value=q,
        root_value=_node_value(root),
        root_visit_count=int(root.visit_count),
        max_tree_depth=int(extra["max_tree_depth"]),
        root_predicted_value=float(extra["root_predicted_value"]),
        n_nodes=_tree_size(root),
        n_tie_events=streams.n_tie_events,
        n_chance_draws=streams.k_chance,
        n_tie_draws=streams.k_tie,
    )
    if record:
        out["root_record"] = rec.root_record
        out["records"] = (numpy.stack(rec.records) if rec.records
                          else numpy.zeros((0, REC_WIDTH), numpy.float32))
        if continuations:
            out["records"] = [numpy.stack(r) if r else numpy.zeros((0, REC_WIDTH), numpy.float32) for r in phase_records]
        out["root_hidden"] = rec.root_hidden
        out["root_logits"] = rec.root_logits
    if keep_root:
        out["root"] = root  # the reference's own Node graph (oracle/make_golden_tree.py flattens it)
    return out
