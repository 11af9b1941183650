"""Weight ingestion: reference ``state_dict`` -> per-MLP lists of (W, b) float32 arrays.

The reference wraps every sub-MLP in ``torch.nn.DataParallel`` (models.py:156-195, 502-518),
so keys look like ``dynamics_encoded_state_network.module.0.weight``; ``mlp()`` interleaves
``Linear`` and activation modules (models.py:1150-1162), so Linear layers sit at the even
indices of the ``Sequential``. PyTorch is only used by callers to obtain the dict; this
module works on anything with ``.items()`` whose values convert with ``numpy.asarray``.
"""
import re
import numpy

# engine MLP ids (same numbering as include/tsp.h TSP_MLP_*)
MLP_IDS = {
    "representation": 0,
    "dynamics": 1,
    "reward": 2,
    "terminal": 3,
    "policy": 4,
    "value": 5,
    "prior": 6,
    # MuZeroStochastic (models.py:334-345): dynamics_encoded_state_network is a ModuleList with one MLP per future;
    # future 0 is "dynamics", futures 1..3 follow
    "dynamics_f1": 7,
    "dynamics_f2": 8,
    "dynamics_f3": 9,
}

_PREFIX = {
    "representation": "representation_network",
    "dynamics": "dynamics_encoded_state_network",
    "reward": "dynamics_reward_network",
    "terminal": "dynamics_terminal_network",
    "policy": "prediction_policy_network",
    "value": "prediction_value_network",
    "prior": "transition_prior_network",
}


def _to_numpy(v):
    if hasattr(v, "detach"):
        v = v.detach().cpu().numpy()
    return numpy.ascontiguousarray(numpy.asarray(v, dtype=numpy.float32))


def extract_mlps(state_dict):
    """Return {mlp name: [(W[out,in], b[out]), ...]} for the MLPs the planner evaluates.
    The reconstruction and posterior heads are never evaluated by MCTS and are skipped."""
    out = {}
    items = dict(state_dict.items())
    prefixes = dict(_PREFIX)
    # separate dynamics heads: keys "dynamics_encoded_state_network.<z>.module.<i>.weight"
    sep = re.compile(r"^dynamics_encoded_state_network\.(\d+)\.module\.\d+\.weight$")
    futures = sorted({int(m.group(1)) for m in map(sep.match, items) if m})
    if futures:
        if futures != list(range(len(futures))) or len(futures) > 4:
            raise ValueError("separate dynamics heads: expected futures 0..n-1 with n <= 4, got %s" % futures)
        for z in futures:
            prefixes["dynamics" if z == 0 else "dynamics_f%d" % z] = "dynamics_encoded_state_network.%d" % z
    for name, prefix in prefixes.items():
        layers = {}
        pat = re.compile(r"^" + re.escape(prefix) + r"\.(?:module\.)?(\d+)\.(weight|bias)$")
        for k, v in items.items():
            m = pat.match(k)
            if m:
                layers.setdefault(int(m.group(1)), {})[m.group(2)] = _to_numpy(v)
        if not layers:
            continue
        seq = []
        for idx in sorted(layers):
            lw = layers[idx]
            if "weight" not in lw or "bias" not in lw:
                raise ValueError("incomplete Linear layer %s.%d in state dict" % (prefix, idx))
            seq.append((lw["weight"], lw["bias"]))
        for (w0, _), (w1, _) in zip(seq[:-1], seq[1:]):
            if w1.shape[1] != w0.shape[0]:
                raise ValueError("layer shape mismatch in %s" % prefix)
        out[name] = seq
    return out
