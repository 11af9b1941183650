"""CPU, world_size 2 over gloo: the multi-rank host logic (contiguous root shards, weight broadcast
from rank 0, all-gather of the root statistics). The search function is a parameter of
`sharded_search`; here it is the CPU oracle (tests may use it as the checker), on the GPU box the
same code path runs with the CUDA engine over NCCL (bench.py)."""
import os
import socket
import sys

import numpy
import pytest

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tree_search_planning_b200.sharding import (shard_bounds, flatten_state_dict, unflatten_state_dict)  # noqa: E402


def test_shard_bounds_partition_the_roots():
    for B in (1, 7, 64, 4096, 65536, 65537):
        for W in (1, 2, 3, 8):
            spans = [shard_bounds(B, W, r) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            for (a, b), (c, d) in zip(spans[:-1], spans[1:]):
                assert b == c and b >= a
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_flatten_roundtrip():
    from tree_search_planning_b200 import get_config
    from tree_search_planning_b200.synthetic import random_state_dict

    sd = random_state_dict(get_config("roundabout_env_ttc_flat_stochastic_concat"), 3)
    flat, meta = flatten_state_dict(sd)
    back = unflatten_state_dict(flat, meta)
    assert set(back) == set(sd)
    for k in sd:
        assert numpy.array_equal(back[k], sd[k])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, B, S, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from tree_search_planning_b200 import get_config
    from tree_search_planning_b200.synthetic import random_state_dict, synthetic_inputs
    from tree_search_planning_b200.sharding import broadcast_state_dict, sharded_search
    from oracle.oracle import Oracle

    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = get_config("roundabout_env_ttc_flat", num_simulations=S)
    sd = random_state_dict(cfg, 5) if rank == 0 else None  # only rank 0 has the weights
    sd = broadcast_state_dict(sd, torch.device("cpu"))
    orc = Oracle(cfg, sd)
    obs, kw = synthetic_inputs(cfg, B, S, seed=9)  # every rank builds the same global inputs, uses its shard

    def search(obs_shard, **shard):
        r = orc.run(S, observations=obs_shard, n_threads=1, **shard)
        return {k: r[k] for k in ("visit_counts", "root_value", "max_tree_depth")}

    res, (lo, hi) = sharded_search(search, obs, kw, world, rank)
    numpy.savez(os.path.join(out_dir, "rank%d.npz" % rank), lo=lo, hi=hi, **res)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_match_one(tmp_path):
    B, S, world = 37, 12, 2  # odd batch: shards of 19 and 18 roots
    port = _free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, S, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    from tree_search_planning_b200 import get_config
    from tree_search_planning_b200.synthetic import random_state_dict, synthetic_inputs
    from oracle.oracle import Oracle

    cfg = get_config("roundabout_env_ttc_flat", num_simulations=S)
    obs, kw = synthetic_inputs(cfg, B, S, seed=9)
    want = Oracle(cfg, random_state_dict(cfg, 5)).run(S, observations=obs, **kw)
    spans = []
    for r in range(world):
        z = numpy.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        spans.append((int(z["lo"]), int(z["hi"])))
        # every rank holds the gathered statistics of ALL roots, identical to the single-process search
        assert numpy.array_equal(z["visit_counts"], want["visit_counts"])
        assert numpy.array_equal(z["root_value"], want["root_value"])
        assert numpy.array_equal(z["max_tree_depth"], want["max_tree_depth"])
    assert spans == [(0, 19), (19, 37)]


def _packed_worker(rank, world, port, B, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from tree_search_planning_b200.sharding import PackedGather, shard_bounds

    dist.init_process_group("gloo", rank=rank, world_size=world)
    A = 5
    pg = PackedGather([("visit_counts", (A,), torch.int32), ("root_value", (), torch.float64)], B, world, rank, "cpu")
    lo, hi = shard_bounds(B, world, rank)
    # what the engine would write into the views of this rank's block of roots
    pg.local["visit_counts"][:hi - lo] = torch.arange(lo * A, hi * A, dtype=torch.int32).reshape(-1, A)
    pg.local["root_value"][:hi - lo] = torch.arange(lo, hi, dtype=torch.float64) * 0.5
    pg.gather()  # ONE collective for both fields
    out = pg.unpack()
    numpy.savez(os.path.join(out_dir, "packed%d.npz" % rank), **{k: v.numpy() for k, v in out.items()})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [37, 64])
def test_packed_gather_one_collective_two_ranks(tmp_path, B):
    """sharding.PackedGather (the multi-GPU product path's single all-gather of visit counts | root values): every
    rank ends up with every root's statistics in root order, for equal and for unequal shards."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_packed_worker, args=(r, world, port, B, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for r in range(world):
        z = numpy.load(os.path.join(str(tmp_path), "packed%d.npz" % r))
        assert numpy.array_equal(z["visit_counts"], numpy.arange(B * 5, dtype=numpy.int32).reshape(B, 5))
        assert numpy.array_equal(z["root_value"], numpy.arange(B, dtype=numpy.float64) * 0.5)
