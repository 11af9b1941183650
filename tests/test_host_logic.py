"""CPU tests of the host-side logic that needs neither a GPU nor the reference tree."""
import os
import sys

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tree_search_planning_b200 import get_config  # noqa: E402
from tree_search_planning_b200.mcts import _weights_fingerprint  # noqa: E402
from tree_search_planning_b200.synthetic import random_state_dict  # noqa: E402


def test_state_dict_fingerprint_follows_the_content_not_the_address():
    cfg = get_config("highway_env_ttc_flat")
    sd = random_state_dict(cfg, 0)
    fp = _weights_fingerprint(sd)
    assert _weights_fingerprint({k: v.copy() for k, v in sd.items()}) == fp  # a fresh dict with the same weights
    k = sorted(sd)[3]
    sd[k][...] = sd[k] + 1e-3  # updated in place: same id(), new content
    assert _weights_fingerprint(sd) != fp
    assert _weights_fingerprint(random_state_dict(cfg, 1)) != fp


def test_module_fingerprint_sees_in_place_updates_and_replaced_parameters():
    import torch

    m = torch.nn.Linear(4, 3)
    fp = _weights_fingerprint(m)
    assert _weights_fingerprint(m) == fp
    with torch.no_grad():
        m.weight.add_(1.0)  # optimizer step / load_state_dict: the version counter moves
    fp2 = _weights_fingerprint(m)
    assert fp2 != fp
    m.weight = torch.nn.Parameter(torch.zeros(3, 4))  # replaced tensor: new storage
    assert _weights_fingerprint(m) != fp2


def test_quantised_columns_mask():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from golden_util import quantised_columns, assert_records_close

    rec = numpy.zeros((3, 16), numpy.float32)
    rec[0, 0] = 1.0  # state record: value, reward
    rec[1, 0] = 2.0  # transition record with 2 futures: rewards at columns 3, 4
    q = quantised_columns(rec, 2)
    assert q[0].nonzero()[0].tolist() == [1, 2] and q[1].nonzero()[0].tolist() == [3, 4] and not q[2].any()
    a = rec.copy()
    a[0, 1] += 1e-4  # one quantisation step on a value: allowed
    assert_records_close(a, rec, 2)
    a[0, 5] += 1e-4  # the same error on a prior: not allowed
    try:
        assert_records_close(a, rec, 2)
    except AssertionError:
        pass
    else:
        raise AssertionError("a prior off by 1e-4 must fail")
