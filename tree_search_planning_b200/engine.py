"""ctypes binding of libtsp_b200.so (include/tsp.h): the batched search engine.

Host code stays Python like the reference; everything below the C ABI is hand-written CUDA.
There is no CPU fallback: without the shared library or without a B200 the constructor raises.
"""
import ctypes
import os

import numpy

from .configs import observation_dim, variant_of
from .weights import MLP_IDS, extract_mlps
from . import build as _build

REC_WIDTH = 16
MEM_HOST, MEM_DEVICE = 0, 1
_P = ctypes.c_void_p


class TspError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("tsp error %d: %s" % (code, message))
        self.code = code


class _Config(ctypes.Structure):
    _fields_ = [
        ("variant", ctypes.c_int32), ("num_actions", ctypes.c_int32), ("obs_dim", ctypes.c_int32),
        ("encoding_size", ctypes.c_int32), ("support_size", ctypes.c_int32), ("n_futures", ctypes.c_int32),
        ("mask_absorbing_states", ctypes.c_int32), ("num_players", ctypes.c_int32),
        ("discount", ctypes.c_double), ("pb_c_base", ctypes.c_double), ("pb_c_init", ctypes.c_double),
        ("root_exploration_fraction", ctypes.c_double),
        ("max_roots", ctypes.c_int32), ("max_simulations", ctypes.c_int32), ("device", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
    ]


class _RunArgs(ctypes.Structure):
    _fields_ = [
        ("num_roots", ctypes.c_int32), ("num_simulations", ctypes.c_int32), ("memory", ctypes.c_int32),
        ("add_exploration_noise", ctypes.c_int32), ("uniform_policy", ctypes.c_int32),
        ("max_records", ctypes.c_int32), ("tie_len", ctypes.c_int32), ("chance_len", ctypes.c_int32),
        ("resume", ctypes.c_int32), ("outputs_on_device", ctypes.c_int32),
        ("observations", _P), ("legal_actions", _P), ("num_legal", _P), ("noise", _P),
        ("tie_stream", _P), ("chance_stream", _P), ("fed_root", _P), ("fed_records", _P),
        ("visit_counts", _P), ("root_value", _P), ("child_value", _P), ("child_prior", _P),
        ("child_reward", _P), ("max_tree_depth", _P), ("root_predicted_value", _P), ("root_hidden", _P),
        ("num_records", _P), ("total_select_depth", _P), ("trace_root", _P), ("trace_records", _P),
    ]


class Stats(ctypes.Structure):
    _fields_ = [
        ("kernel_launches", ctypes.c_int64), ("searches", ctypes.c_int64), ("bytes_h2d", ctypes.c_int64),
        ("bytes_d2h", ctypes.c_int64), ("pool_bytes", ctypes.c_int64), ("last_search_ms", ctypes.c_float),
        ("last_root_ms", ctypes.c_float), ("sum_search_ms", ctypes.c_double), ("sum_root_ms", ctypes.c_double),
        ("timed_runs", ctypes.c_int64), ("sm_count", ctypes.c_int32), ("search_ctas", ctypes.c_int32),
        ("search_smem_bytes", ctypes.c_int32), ("trees_per_cta", ctypes.c_int32),
        ("threads_per_cta", ctypes.c_int32), ("search_kernel", ctypes.c_int32), ("root_kernel", ctypes.c_int32),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# every symbol include/tsp.h declares (tests check that the library exports each of them)
class TreeInfo(ctypes.Structure):
    """tsp_tree_info of include/tsp.h."""
    _fields_ = [("n_blocks", ctypes.c_int32), ("block_bytes", ctypes.c_int32), ("num_children", ctypes.c_int32),
                ("n_root", ctypes.c_int32), ("root_visit", ctypes.c_int32), ("n_hidden", ctypes.c_int32),
                ("root_value_sum", ctypes.c_double), ("root_prior", ctypes.c_double * 8), ("root_action", ctypes.c_int32 * 8)]


def tree_block_dtype(num_children):
    """numpy view of one exported node block (tsp_export_tree; NodeBlock of csrc/tsp_device.cuh)."""
    nc = int(num_children)
    return numpy.dtype({"names": ["visit", "child", "prior", "reward", "parent", "aux", "value_sum"],
                        "formats": [(numpy.int32, nc), (numpy.int32, nc), (numpy.float32, nc), (numpy.float32, nc),
                                    numpy.int32, numpy.int32, (numpy.float64, nc)],
                        "offsets": [0, 4 * nc, 8 * nc, 12 * nc, 16 * nc, 16 * nc + 4, 16 * nc + 8],
                        "itemsize": 128 if nc == 5 else 256})


ABI_SYMBOLS = (
    "tsp_abi_version", "tsp_create", "tsp_destroy", "tsp_last_error", "tsp_set_layer", "tsp_finalize_weights",
    "tsp_weight_arena", "tsp_run", "tsp_initial_inference", "tsp_recurrent_inference", "tsp_stochastic_dynamics",
    "tsp_prediction", "tsp_synchronize", "tsp_get_stats", "tsp_reset_timing", "tsp_debug_phase_cycles", "tsp_stream",
    "tsp_stack_observations", "tsp_select_action", "tsp_search_statistics", "tsp_export_tree", "tsp_peer_put",
)

_lib = None


def load_library(path=None):
    """Load libtsp_b200.so. Fails loudly when it is missing. It is (re)built in-tree only when it does not exist, or
    when TSP_AUTOBUILD=1 and a source is newer (after a checkout / rsync the mtimes say nothing, and on a GPU box with
    one process per GPU nobody wants N concurrent nvcc runs); `__graft_entry__.build()` is the explicit build."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or os.environ.get("TSP_LIBRARY")  # (developer A/B runs: another build of the same ABI)
    p = path or _build.LIB_PATH
    if path is None and (not os.path.exists(p) or (os.environ.get("TSP_AUTOBUILD") == "1" and _build.needs_build())):
        try:
            _build.build()
        except RuntimeError as e:
            raise TspError(-2, "CUDA extension %s is missing or stale and could not be built: %s" % (p, e))
    if not os.path.exists(p):
        raise TspError(-2, "CUDA extension %s is missing; run `python -c 'import __graft_entry__ as g; g.build()'`" % p)
    lib = ctypes.CDLL(p)
    lib.tsp_abi_version.restype = ctypes.c_int
    lib.tsp_last_error.restype = ctypes.c_char_p
    lib.tsp_last_error.argtypes = [_P]
    lib.tsp_create.argtypes = [ctypes.POINTER(_Config), ctypes.POINTER(_P)]
    lib.tsp_destroy.argtypes = [_P]
    lib.tsp_set_layer.argtypes = [_P, ctypes.c_int32, ctypes.c_int32, _P, _P, ctypes.c_int32, ctypes.c_int32]
    lib.tsp_finalize_weights.argtypes = [_P]
    lib.tsp_weight_arena.argtypes = [_P, ctypes.POINTER(_P), ctypes.POINTER(ctypes.c_int64)]
    lib.tsp_run.argtypes = [_P, ctypes.POINTER(_RunArgs)]
    lib.tsp_initial_inference.argtypes = [_P, ctypes.c_int32, ctypes.c_int32] + [_P] * 5
    lib.tsp_recurrent_inference.argtypes = [_P, ctypes.c_int32, ctypes.c_int32] + [_P] * 9
    lib.tsp_stochastic_dynamics.argtypes = [_P, ctypes.c_int32, ctypes.c_int32] + [_P] * 5
    lib.tsp_prediction.argtypes = [_P, ctypes.c_int32, ctypes.c_int32] + [_P] * 3
    lib.tsp_synchronize.argtypes = [_P]
    lib.tsp_get_stats.argtypes = [_P, ctypes.POINTER(Stats)]
    lib.tsp_reset_timing.argtypes = [_P]
    lib.tsp_debug_phase_cycles.argtypes = [_P, ctypes.POINTER(ctypes.c_uint64 * 64), ctypes.c_int32]
    lib.tsp_stream.argtypes = [_P, ctypes.POINTER(_P)]
    lib.tsp_stack_observations.argtypes = [_P] + [ctypes.c_int32] * 5 + [_P] * 4
    lib.tsp_select_action.argtypes = [_P, ctypes.c_int32, ctypes.c_int32, _P, _P, _P, ctypes.c_double, _P, _P]
    lib.tsp_search_statistics.argtypes = [_P, ctypes.c_int32, ctypes.c_int32, _P, _P]
    lib.tsp_peer_put.argtypes = [_P, _P, ctypes.c_int64, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int32]
    lib.tsp_export_tree.argtypes = [_P, ctypes.c_int32, ctypes.POINTER(TreeInfo), _P, ctypes.c_int32, _P, ctypes.c_int32]
    if path is None:
        _lib = lib
    return lib


def _is_torch_tensor(x):
    return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")


def _ptr(x):
    if x is None:
        return None
    if _is_torch_tensor(x):
        return _P(x.data_ptr())
    return x.ctypes.data_as(_P)


class Engine:
    """One engine per (game config, CUDA device). Node pools are sized at construction."""

    def __init__(self, config, max_roots, max_simulations=None, device=0, state_dict=None):
        self.lib = load_library()
        self.config = config
        self.variant = variant_of(config)
        self.A = len(config.action_space)
        if list(config.action_space) != list(range(self.A)):
            raise ValueError("action_space must be list(range(A)) (the networks one-hot encode the action index)")
        self.H = int(config.encoding_size)
        self.obs_dim = observation_dim(config)
        self.nf = int(getattr(config, "n_futures", 1)) if self.variant else 1
        self.V = 2 * int(config.support_size) + 1
        self.max_roots = int(max_roots)
        self.max_simulations = int(config.num_simulations if max_simulations is None else max_simulations)
        self.device = int(device)
        self.generation = 0
        if len(config.players) > 2:
            raise NotImplementedError("More than two player mode not implemented.")  # self_play.py:503
        c = _Config(self.variant, self.A, self.obs_dim, self.H, int(config.support_size), self.nf,
                    int(bool(getattr(config, "mask_absorbing_states", False))), len(config.players),
                    float(config.discount), float(config.pb_c_base), float(config.pb_c_init),
                    float(config.root_exploration_fraction), self.max_roots, self.max_simulations, self.device, 0)
        h = _P()
        rc = self.lib.tsp_create(ctypes.byref(c), ctypes.byref(h))
        if rc != 0:
            msg = self.lib.tsp_last_error(None).decode()
            if rc == -5:
                raise NotImplementedError(msg)
            raise TspError(rc, msg)
        self.handle = h
        self._keep = []
        if state_dict is not None:
            self.load_state_dict(state_dict)

    # -- lifetime ---------------------------------------------------------------------
    def close(self):
        if getattr(self, "handle", None):
            self.lib.tsp_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise TspError(rc, self.lib.tsp_last_error(self.handle).decode())

    # -- weights ------------------------------------------------------------------------
    def load_state_dict(self, state_dict):
        """model.set_weights equivalent (models.py:128-129): accepts the reference state dict
        (torch tensors or numpy arrays, with or without the DataParallel `.module.` infix)."""
        mlps = extract_mlps(state_dict)
        for name, layers in mlps.items():
            if name == "prior" and self.variant == 0:
                continue
            for i, (W, b) in enumerate(layers):
                W = numpy.ascontiguousarray(W, numpy.float32)
                b = numpy.ascontiguousarray(b, numpy.float32)
                self._check(self.lib.tsp_set_layer(self.handle, MLP_IDS[name], i, _ptr(W), _ptr(b),
                                                   W.shape[0], W.shape[1]))
        self._check(self.lib.tsp_finalize_weights(self.handle))

    def weight_arena(self):
        p, n = _P(), ctypes.c_int64()
        self._check(self.lib.tsp_weight_arena(self.handle, ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    # -- search -------------------------------------------------------------------------
    def max_records(self, S):
        return S if self.variant == 0 else S * (self.A + 2) + self.A + 2

    def run_raw(self, args):
        self.generation += 1  # the pools now hold the trees of this search
        self._check(self.lib.tsp_run(self.handle, ctypes.byref(args)))

    def run(self, num_simulations=None, observations=None, legal_actions=None, num_legal=None, noise=None,
            tie=None, chance=None, add_exploration_noise=True, uniform_policy=False, fed_root=None,
            fed_records=None, trace=False, max_records=None, want=("visit_counts", "root_value", "child_value",
                                                                   "child_prior", "child_reward", "max_tree_depth",
                                                                   "root_predicted_value", "num_records"),
            want_root_hidden=False, synchronize=True, to_play=None, resume=False, num_roots=None):
        """Batched MCTS.run over HOST numpy arrays; returns a dict of numpy arrays.

        `to_play` ([B], MCTS.run's argument) is accepted for interface parity only: with players == [0, 1] the
        reference's arithmetic depends on whether a node and the leaf are an even number of plies apart
        (self_play.py:492-500), never on which player the root belongs to.

        `resume=True` continues the previous run's searches (MCTS.run's override_root_with, self_play.py:331-333):
        no observations, `num_roots` roots, fresh MinMaxStats, noise mixed into the root priors again."""
        S = self.max_simulations if num_simulations is None else int(num_simulations)
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
            observations = prep(observations, numpy.float32).reshape(-1, self.obs_dim)
            B = observations.shape[0]
        elif fed_root is not None:
            B = numpy.asarray(fed_root).reshape(-1, REC_WIDTH).shape[0]
        elif resume and num_roots is not None:
            B = int(num_roots)
        else:
            raise ValueError("observations (or fed_root, or resume=True with num_roots) required")
        fed_root = prep(fed_root, numpy.float32, (B, REC_WIDTH))
        if fed_records is not None:
            fed_records = prep(fed_records, numpy.float32).reshape(B, -1, REC_WIDTH)
            max_records = fed_records.shape[1]
        if max_records is None:
            max_records = self.max_records(S) if trace else 0
        legal_actions = prep(legal_actions, numpy.int32, (B, A))
        num_legal = prep(num_legal, numpy.int32, (B,))
        noise = prep(noise, numpy.float64, (B, A))
        tie = prep(tie, numpy.uint8)
        T = 0 if tie is None else tie.reshape(B, -1).shape[1]
        chance = prep(chance, numpy.uint8)
        K = 0 if chance is None else chance.reshape(B, -1).shape[1]
        shapes = dict(visit_counts=((B, A), numpy.int32), root_value=((B,), numpy.float64),
                      child_value=((B, A), numpy.float64), child_prior=((B, A), numpy.float64),
                      child_reward=((B, A), numpy.float64), max_tree_depth=((B,), numpy.int32),
                      root_predicted_value=((B,), numpy.float32), num_records=((B,), numpy.int32),
                      total_select_depth=((B,), numpy.int32))
        out = {k: numpy.zeros(*shapes[k]) for k in want}
        if want_root_hidden:
            out["root_hidden"] = numpy.zeros((B, self.H), numpy.float32)
        if trace:
            out["trace_root"] = numpy.zeros((B, REC_WIDTH), numpy.float32)
            out["trace_records"] = numpy.zeros((B, max(max_records, 1), REC_WIDTH), numpy.float32)
        a = _RunArgs(B, S, MEM_HOST, int(bool(add_exploration_noise)), int(bool(uniform_policy)),
                     int(max_records), T, K, int(bool(resume)), 0,
                     _ptr(observations), _ptr(legal_actions), _ptr(num_legal), _ptr(noise), _ptr(tie), _ptr(chance),
                     _ptr(fed_root), _ptr(fed_records),
                     _ptr(out.get("visit_counts")), _ptr(out.get("root_value")), _ptr(out.get("child_value")),
                     _ptr(out.get("child_prior")), _ptr(out.get("child_reward")), _ptr(out.get("max_tree_depth")),
                     _ptr(out.get("root_predicted_value")), _ptr(out.get("root_hidden")), _ptr(out.get("num_records")),
                     _ptr(out.get("total_select_depth")), _ptr(out.get("trace_root")), _ptr(out.get("trace_records")))
        self.run_raw(a)
        if synchronize:
            self.synchronize()
        else:
            self._keep = keep  # keep host inputs alive until the caller synchronizes
        return out

    # -- root I/O formats on the device ---------------------------------------------
    def stack_observations(self, frame_pool, frame_index, action_value, out, stacked_observations, frame_size,
                           plane_size):
        """GameHistory.get_stacked_observations for B rows; DEVICE buffers (torch tensors), asynchronous."""
        B = int(frame_index.shape[0])
        self._check(self.lib.tsp_stack_observations(self.handle, MEM_DEVICE, B, int(stacked_observations),
                                                    int(frame_size), int(plane_size), _ptr(frame_pool),
                                                    _ptr(frame_index), _ptr(action_value), _ptr(out)))

    def select_action(self, visit_counts, temperature, uniform=None, legal_actions=None, num_legal=None, out=None,
                      memory=None):
        """SelfPlay.select_action for B roots. numpy arrays (HOST, synchronous) or torch CUDA tensors (DEVICE)."""
        dev = _is_torch_tensor(visit_counts)
        mem = (MEM_DEVICE if dev else MEM_HOST) if memory is None else memory
        B = int(visit_counts.shape[0])
        if not dev:
            visit_counts = numpy.ascontiguousarray(visit_counts, numpy.int32)
            uniform = None if uniform is None else numpy.ascontiguousarray(uniform, numpy.float64)
            legal_actions = None if legal_actions is None else numpy.ascontiguousarray(legal_actions, numpy.int32)
            num_legal = None if num_legal is None else numpy.ascontiguousarray(num_legal, numpy.int32)
            out = numpy.zeros(B, numpy.int32) if out is None else out
        self._check(self.lib.tsp_select_action(self.handle, mem, B, _ptr(visit_counts), _ptr(legal_actions),
                                               _ptr(num_legal), float(temperature), _ptr(uniform), _ptr(out)))
        if not dev:
            self.synchronize()
        return out

    def search_statistics(self, visit_counts, out=None, memory=None):
        """The policy target of GameHistory.store_search_statistics for B roots: visits / max(sum, 1), float64."""
        dev = _is_torch_tensor(visit_counts)
        mem = (MEM_DEVICE if dev else MEM_HOST) if memory is None else memory
        B = int(visit_counts.shape[0])
        if not dev:
            visit_counts = numpy.ascontiguousarray(visit_counts, numpy.int32)
            out = numpy.zeros((B, self.A), numpy.float64) if out is None else out
        self._check(self.lib.tsp_search_statistics(self.handle, mem, B, _ptr(visit_counts), _ptr(out)))
        if not dev:
            self.synchronize()
        return out

    # -- network entry points --------------------------------------------------------------
    def initial_inference(self, observations):
        obs = numpy.ascontiguousarray(observations, numpy.float32).reshape(-1, self.obs_dim)
        B = obs.shape[0]
        out = dict(value=numpy.zeros(B, numpy.float32), policy_logits=numpy.zeros((B, self.A), numpy.float32),
                   hidden=numpy.zeros((B, self.H), numpy.float32), value_logits=numpy.zeros((B, self.V), numpy.float32))
        self._check(self.lib.tsp_initial_inference(self.handle, MEM_HOST, B, _ptr(obs), _ptr(out["value"]),
                                                   _ptr(out["policy_logits"]), _ptr(out["hidden"]),
                                                   _ptr(out["value_logits"])))
        self.synchronize()
        return out

    def initial_inference_device(self, observations, value=None, policy_logits=None, hidden=None, value_logits=None):
        """tsp_initial_inference over DEVICE buffers (torch CUDA tensors), asynchronous on the engine's stream."""
        B = int(observations.shape[0])
        self._check(self.lib.tsp_initial_inference(self.handle, MEM_DEVICE, B, _ptr(observations), _ptr(value),
                                                   _ptr(policy_logits), _ptr(hidden), _ptr(value_logits)))

    def recurrent_inference(self, hidden, action):
        h = numpy.ascontiguousarray(hidden, numpy.float32).reshape(-1, self.H)
        B = h.shape[0]
        a = numpy.ascontiguousarray(action, numpy.int32).reshape(B)
        out = dict(value=numpy.zeros(B, numpy.float32), reward=numpy.zeros(B, numpy.float32),
                   terminal=numpy.zeros(B, numpy.float32), policy_logits=numpy.zeros((B, self.A), numpy.float32),
                   hidden=numpy.zeros((B, self.H), numpy.float32), value_logits=numpy.zeros((B, self.V), numpy.float32),
                   reward_logits=numpy.zeros((B, self.V), numpy.float32))
        self._check(self.lib.tsp_recurrent_inference(
            self.handle, MEM_HOST, B, _ptr(h), _ptr(a), _ptr(out["value"]), _ptr(out["reward"]), _ptr(out["terminal"]),
            _ptr(out["policy_logits"]), _ptr(out["hidden"]), _ptr(out["value_logits"]), _ptr(out["reward_logits"])))
        self.synchronize()
        return out

    def stochastic_dynamics(self, hidden, action):
        h = numpy.ascontiguousarray(hidden, numpy.float32).reshape(-1, self.H)
        B = h.shape[0]
        a = numpy.ascontiguousarray(action, numpy.int32).reshape(B)
        out = dict(transition_logits=numpy.zeros((B, self.nf), numpy.float32),
                   hidden=numpy.zeros((B, self.nf, self.H), numpy.float32), reward=numpy.zeros((B, self.nf), numpy.float32))
        self._check(self.lib.tsp_stochastic_dynamics(self.handle, MEM_HOST, B, _ptr(h), _ptr(a),
                                                     _ptr(out["transition_logits"]), _ptr(out["hidden"]),
                                                     _ptr(out["reward"])))
        self.synchronize()
        return out

    def prediction(self, hidden):
        h = numpy.ascontiguousarray(hidden, numpy.float32).reshape(-1, self.H)
        B = h.shape[0]
        out = dict(value=numpy.zeros(B, numpy.float32), policy_logits=numpy.zeros((B, self.A), numpy.float32))
        self._check(self.lib.tsp_prediction(self.handle, MEM_HOST, B, _ptr(h), _ptr(out["value"]),
                                            _ptr(out["policy_logits"])))
        self.synchronize()
        return out

    # -- misc ----------------------------------------------------------------------------------
    def export_tree(self, root_index, want_hidden=True):
        """The device-resident tree of one root of the latest search (tsp_export_tree): dict(info=TreeInfo,
        blocks=structured array [n_blocks], hidden=[n_hidden, H] float32 or None)."""
        info = TreeInfo()
        self._check(self.lib.tsp_export_tree(self.handle, int(root_index), ctypes.byref(info), None, 0, None, 0))
        blocks = numpy.zeros(info.n_blocks, tree_block_dtype(info.num_children))
        hidden = numpy.zeros((info.n_hidden, self.H), numpy.float32) if want_hidden and info.n_hidden > 0 else None
        self._check(self.lib.tsp_export_tree(self.handle, int(root_index), ctypes.byref(info), _ptr(blocks), info.n_blocks,
                                             _ptr(hidden) if hidden is not None else None, info.n_hidden if hidden is not None else 0))
        return dict(info=info, blocks=blocks, hidden=hidden)

    def peer_put(self, src, num_bytes, peer_ptrs):
        """tsp_peer_put: `src` (device tensor / pointer) -> every device pointer of `peer_ptrs`, on the engine's stream."""
        arr = (ctypes.c_void_p * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
        self._check(self.lib.tsp_peer_put(self.handle, _ptr(src), int(num_bytes), arr, len(peer_ptrs)))

    def synchronize(self):
        self._check(self.lib.tsp_synchronize(self.handle))
        self._keep = []

    def stats(self):
        s = Stats()
        self._check(self.lib.tsp_get_stats(self.handle, ctypes.byref(s)))
        return s.as_dict()

    def phase_cycles(self, reset=True):
        """Developer instrumentation (needs TSP_PHASE_TIMING=1 in the environment at construction)."""
        buf = (ctypes.c_uint64 * 64)()
        self._check(self.lib.tsp_debug_phase_cycles(self.handle, ctypes.byref(buf), int(reset)))
        v = list(buf)
        n = max(1, v[4])
        stages = [(round(v[8 + 2 * s] / n), round(v[9 + 2 * s] / n)) for s in range(14) if v[8 + 2 * s] or v[9 + 2 * s]]
        nt = max(1, v[43])
        return dict(select=v[0] / n, wait=v[1] / n, mlp=v[2] / n, backup=v[3] / n, samples=v[4],
                    stages_work_wait=stages, tile_init_kloop_store=(round(v[40] / nt), round(v[41] / nt), round(v[42] / nt)),
                    tiles_per_iter=v[43] / n, exact_selects=v[5], select_levels=v[6],
                    # streamed tcgen05 kernel (tsp_stream.cuh), cycles per simulation: issuer waiting for the B operand of
                    # stage 0..3, for full slots, issuer total; producer waiting for a free slot, total; next-state stage:
                    # accumulators + min / max, REDUX + atomics, barrier, second pass
                    stream=dict(issuer_wait_b=[round(v[16 + i] / n) for i in range(4)], issuer_wait_full=round(v[20] / n),
                                issuer_total=round(v[21] / n), producer_wait_empty=round(v[24] / n), producer_total=round(v[25] / n),
                                xs_pass1=round(v[28] / n), xs_redux=round(v[29] / n), xs_barrier=round(v[30] / n),
                                xs_pass2=round(v[31] / n)) if v[21] else None,
                    tc_issue_wait_epilogue=[tuple(round(v[44 + 4 * s + i] / n) for i in range(3)) for s in range(5)
                                            if any(v[44 + 4 * s: 47 + 4 * s])])

    def reset_timing(self):
        self._check(self.lib.tsp_reset_timing(self.handle))

    def make_args(self, num_roots, num_simulations, memory, inputs, outputs, add_exploration_noise=True,
                  uniform_policy=False, max_records=0, tie_len=0, chance_len=0, resume=False, outputs_on_device=False):
        """Build a tsp_run_args over caller-owned buffers (numpy arrays, or torch tensors for device /
        pinned memory). `inputs` / `outputs` map the field names of tsp_run_args to buffers."""
        a = _RunArgs()
        a.num_roots, a.num_simulations, a.memory = int(num_roots), int(num_simulations), int(memory)
        a.add_exploration_noise, a.uniform_policy = int(bool(add_exploration_noise)), int(bool(uniform_policy))
        a.max_records, a.tie_len, a.chance_len = int(max_records), int(tie_len), int(chance_len)
        a.resume = int(bool(resume))
        a.outputs_on_device = int(bool(outputs_on_device))
        names = {n for n, _ in _RunArgs._fields_}
        for k, v in list(inputs.items()) + list(outputs.items()):
            if k not in names:
                raise KeyError(k)
            setattr(a, k, _ptr(v))
        a._keep = (inputs, outputs)
        return a

    def stream(self):
        p = _P()
        self._check(self.lib.tsp_stream(self.handle, ctypes.byref(p)))
        return p.value
