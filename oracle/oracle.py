"""TEST INFRASTRUCTURE ONLY -- ctypes binding of the C oracle (oracle/tsp_oracle.c).

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs only. The product package never imports this module.
"""
import ctypes
import os
import subprocess

import numpy

from tree_search_planning_b200.configs import variant_of, observation_dim
from tree_search_planning_b200.weights import extract_mlps, MLP_IDS

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libtsp_oracle.so")
REC_WIDTH = 16


def build(force=False):
    src = os.path.join(_HERE, "tsp_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"] if force else ["make", "-C", _HERE, "-s"])
    return _LIB_PATH


class _Config(ctypes.Structure):
    _fields_ = [
        ("variant", ctypes.c_int), ("num_actions", ctypes.c_int), ("obs_dim", ctypes.c_int),
        ("encoding_size", ctypes.c_int), ("support_size", ctypes.c_int), ("n_futures", ctypes.c_int),
        ("mask_absorbing_states", ctypes.c_int), ("num_players", ctypes.c_int),
        ("discount", ctypes.c_double), ("pb_c_base", ctypes.c_double), ("pb_c_init", ctypes.c_double),
        ("root_exploration_fraction", ctypes.c_double),
    ]


_P = ctypes.c_void_p


class _RunArgs(ctypes.Structure):
    _fields_ = [
        ("B", ctypes.c_int), ("S", ctypes.c_int), ("add_noise", ctypes.c_int), ("uniform_policy", ctypes.c_int),
        ("obs", _P), ("legal", _P), ("n_legal", _P), ("noise", _P),
        ("tie", _P), ("T", ctypes.c_int), ("chance", _P), ("K", ctypes.c_int),
        ("fed_root", _P), ("fed_records", _P), ("max_records", ctypes.c_int),
        ("visit_counts", _P), ("root_value", _P), ("child_value", _P), ("child_prior", _P),
        ("child_reward", _P), ("max_depth", _P), ("root_pred_value", _P), ("root_hidden", _P),
        ("n_records", _P), ("n_nodes", _P), ("trace_root", _P), ("trace_records", _P),
        ("n_threads", ctypes.c_int), ("to_play", _P), ("n_phases", ctypes.c_int),
    ]


_lib = None


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_LIB_PATH)
        lib.orc_net_create.restype = _P
        lib.orc_net_create.argtypes = [ctypes.POINTER(_Config)]
        lib.orc_net_destroy.argtypes = [_P]
        lib.orc_net_set_layer.argtypes = [_P, ctypes.c_int, ctypes.c_int, _P, _P, ctypes.c_int, ctypes.c_int]
        lib.orc_run.argtypes = [_P, ctypes.POINTER(_RunArgs)]
        lib.orc_run.restype = ctypes.c_int
        lib.orc_max_threads.restype = ctypes.c_int
        lib.orc_inverse_value_transform.restype = ctypes.c_float
        lib.orc_inverse_value_transform.argtypes = [ctypes.c_float]
        lib.orc_support_to_scalar.restype = ctypes.c_float
        lib.orc_support_to_scalar.argtypes = [_P, ctypes.c_int]
        _lib = lib
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_P)


def max_threads():
    return _load().orc_max_threads()


def inverse_value_transform(x):
    return float(_load().orc_inverse_value_transform(ctypes.c_float(float(x))))


def support_to_scalar(logits, support_size):
    lg = numpy.ascontiguousarray(logits, numpy.float32)
    return float(_load().orc_support_to_scalar(_ptr(lg), int(support_size)))


class Oracle:
    """CPU oracle for one game config + one set of weights."""

    def __init__(self, config, state_dict=None):
        self.lib = _load()
        self.config = config
        self.variant = variant_of(config)
        self.A = len(config.action_space)
        self.H = int(config.encoding_size)
        self.obs_dim = observation_dim(config)
        self.nf = int(getattr(config, "n_futures", 1)) if self.variant else 1
        self.V = 2 * int(config.support_size) + 1
        c = _Config(self.variant, self.A, self.obs_dim, self.H, int(config.support_size), self.nf,
                    int(bool(getattr(config, "mask_absorbing_states", False))), len(config.players),
                    float(config.discount), float(config.pb_c_base), float(config.pb_c_init),
                    float(config.root_exploration_fraction))
        self._cfg = c
        self.net = self.lib.orc_net_create(ctypes.byref(c))
        if state_dict is not None:
            self.load_state_dict(state_dict)

    def __del__(self):
        try:
            self.lib.orc_net_destroy(self.net)
        except Exception:
            pass

    def load_state_dict(self, state_dict):
        for name, layers in extract_mlps(state_dict).items():
            for i, (W, b) in enumerate(layers):
                W = numpy.ascontiguousarray(W, numpy.float32)
                b = numpy.ascontiguousarray(b, numpy.float32)
                rc = self.lib.orc_net_set_layer(self.net, MLP_IDS[name], i, _ptr(W), _ptr(b), W.shape[0], W.shape[1])
                assert rc == 0

    # -- network entry points ------------------------------------------------------
    def initial_inference(self, obs):
        obs = numpy.ascontiguousarray(obs, numpy.float32).reshape(-1, self.obs_dim)
        B = obs.shape[0]
        value = numpy.zeros(B, numpy.float32)
        logits = numpy.zeros((B, self.A), numpy.float32)
        hidden = numpy.zeros((B, self.H), numpy.float32)
        vlog = numpy.zeros((B, self.V), numpy.float32)
        self.lib.orc_initial_inference(_P(self.net), _ptr(obs), B, _ptr(value), _ptr(logits), _ptr(hidden), _ptr(vlog))
        return dict(value=value, policy_logits=logits, hidden=hidden, value_logits=vlog)

    def recurrent_inference(self, hidden, action):
        hidden = numpy.ascontiguousarray(hidden, numpy.float32).reshape(-1, self.H)
        B = hidden.shape[0]
        action = numpy.ascontiguousarray(action, numpy.int32).reshape(B)
        value = numpy.zeros(B, numpy.float32)
        reward = numpy.zeros(B, numpy.float32)
        term = numpy.zeros(B, numpy.float32)
        logits = numpy.zeros((B, self.A), numpy.float32)
        nh = numpy.zeros((B, self.H), numpy.float32)
        vlog = numpy.zeros((B, self.V), numpy.float32)
        rlog = numpy.zeros((B, self.V), numpy.float32)
        self.lib.orc_recurrent_inference(_P(self.net), _ptr(hidden), _ptr(action), B, _ptr(value), _ptr(reward),
                                         _ptr(term), _ptr(logits), _ptr(nh), _ptr(vlog), _ptr(rlog))
        return dict(value=value, reward=reward, terminal=term, policy_logits=logits, hidden=nh,
                    value_logits=vlog, reward_logits=rlog)

    def stochastic_dynamics(self, hidden, action):
        hidden = numpy.ascontiguousarray(hidden, numpy.float32).reshape(-1, self.H)
        B = hidden.shape[0]
        action = numpy.ascontiguousarray(action, numpy.int32).reshape(B)
        tl = numpy.zeros((B, self.nf), numpy.float32)
        nh = numpy.zeros((B, self.nf, self.H), numpy.float32)
        rw = numpy.zeros((B, self.nf), numpy.float32)
        self.lib.orc_stochastic_dynamics(_P(self.net), _ptr(hidden), _ptr(action), B, _ptr(tl), _ptr(nh), _ptr(rw))
        return dict(transition_logits=tl, hidden=nh, reward=rw)

    def prediction(self, hidden):
        hidden = numpy.ascontiguousarray(hidden, numpy.float32).reshape(-1, self.H)
        B = hidden.shape[0]
        value = numpy.zeros(B, numpy.float32)
        logits = numpy.zeros((B, self.A), numpy.float32)
        self.lib.orc_prediction(_P(self.net), _ptr(hidden), B, _ptr(value), _ptr(logits))
        return dict(value=value, policy_logits=logits)

    # -- batched search ------------------------------------------------------------
    def max_records(self, S):
        return S if self.variant == 0 else S * (self.A + 2) + self.A + 2

    def run(self, num_simulations, observations=None, legal_actions=None, num_legal=None, noise=None,
            tie=None, chance=None, add_exploration_noise=True, uniform_policy=False,
            fed_root=None, fed_records=None, trace=False, max_records=None, n_threads=0, B=None,
            want_root_hidden=False, to_play=None, n_phases=1):
        """n_phases > 1: the search is continued with override_root_with=root (self_play.py:331-333) n_phases - 1
        times, S simulations each; noise / tie / chance / fed_records then have a leading [n_phases] dimension and
        the trace records come back as [n_phases, B, max_records, REC]."""
        S = int(num_simulations)
        P = int(n_phases)
        A = self.A
        keep = []

        def prep(x, dt, shape=None):
            if x is None:
                return None
            x = numpy.ascontiguousarray(x, dt)
            if shape is not None:
                x = x.reshape(shape)
            keep.append(x)
            return x

        if observations is not None:
            observations = prep(observations, numpy.float32)
            observations = observations.reshape(-1, self.obs_dim)
            B = observations.shape[0]
        elif fed_root is not None:
            B = numpy.asarray(fed_root).reshape(-1, REC_WIDTH).shape[0]
        assert B is not None
        fed_root = prep(fed_root, numpy.float32, (B, REC_WIDTH))
        if fed_records is not None:
            fed_records = prep(fed_records, numpy.float32)
            mr = fed_records.reshape(P * B, -1, REC_WIDTH).shape[1]
            if max_records is None:
                max_records = mr
            assert mr == max_records
        if max_records is None:
            max_records = self.max_records(S) if trace else (self.max_records(S))
        legal_actions = prep(legal_actions, numpy.int32, (B, A))
        num_legal = prep(num_legal, numpy.int32, (B,))
        noise = prep(noise, numpy.float64, (P * B, A))
        to_play = prep(to_play, numpy.int32, (B,))
        tie = prep(tie, numpy.uint8)
        T = 0 if tie is None else tie.reshape(P * B, -1).shape[1]
        chance = prep(chance, numpy.uint8)
        K = 0 if chance is None else chance.reshape(P * B, -1).shape[1]
        if add_exploration_noise:
            assert noise is not None, "exploration noise requested but no noise supplied"
        out = dict(
            visit_counts=numpy.zeros((B, A), numpy.int32),
            root_value=numpy.zeros(B, numpy.float64),
            child_value=numpy.zeros((B, A), numpy.float64),
            child_prior=numpy.zeros((B, A), numpy.float64),
            child_reward=numpy.zeros((B, A), numpy.float64),
            max_tree_depth=numpy.zeros(B, numpy.int32),
            root_predicted_value=numpy.zeros(B, numpy.float32),
            n_records=numpy.zeros(B, numpy.int32),
            n_nodes=numpy.zeros(B, numpy.int32),
        )
        root_hidden = numpy.zeros((B, self.H), numpy.float32) if want_root_hidden else None
        trace_root = numpy.zeros((B, REC_WIDTH), numpy.float32) if trace else None
        trace_records = numpy.zeros((P * B, max_records, REC_WIDTH), numpy.float32) if trace else None
        a = _RunArgs(B, S, int(bool(add_exploration_noise)), int(bool(uniform_policy)),
                     _ptr(observations), _ptr(legal_actions), _ptr(num_legal), _ptr(noise),
                     _ptr(tie), T, _ptr(chance), K,
                     _ptr(fed_root), _ptr(fed_records), int(max_records),
                     _ptr(out["visit_counts"]), _ptr(out["root_value"]), _ptr(out["child_value"]),
                     _ptr(out["child_prior"]), _ptr(out["child_reward"]), _ptr(out["max_tree_depth"]),
                     _ptr(out["root_predicted_value"]), _ptr(root_hidden), _ptr(out["n_records"]),
                     _ptr(out["n_nodes"]), _ptr(trace_root), _ptr(trace_records), int(n_threads), _ptr(to_play), P)
        rc = self.lib.orc_run(_P(self.net), ctypes.byref(a))
        if rc != 0:
            raise RuntimeError("oracle run failed with code %d" % rc)
        if trace:
            out["trace_root"] = trace_root
            out["trace_records"] = trace_records if P == 1 else trace_records.reshape(P, B, max_records, REC_WIDTH)
        if want_root_hidden:
            out["root_hidden"] = root_hidden
        return out
