"""-m gpu, needs TWO GPUs (skipped otherwise): the multi-GPU product path (sharding.ShardedEngine: one process per GPU,
NCCL weight broadcast, roots sharded, ONE all-gather of the packed root statistics) gives every rank exactly the
single-GPU result, bit for bit, for an odd batch (unequal shards)."""
import os
import socket
import sys

import numpy
import pytest

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, game, B, S, out_dir, p2p):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      TSP_NO_P2P_GATHER="0" if p2p else "1")
    import torch.distributed as dist
    import tree_search_planning_b200 as tsp
    from tree_search_planning_b200.sharding import ShardedEngine
    from tree_search_planning_b200.synthetic import random_state_dict, synthetic_inputs

    device = torch.device("cuda", rank)
    torch.cuda.set_device(device)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    cfg = tsp.get_config(game, num_simulations=S)
    se = ShardedEngine(cfg, num_roots=B, max_simulations=S, device=device,
                       state_dict=random_state_dict(cfg, 3) if rank == 0 else None)  # only rank 0 has the weights
    obs, kw = synthetic_inputs(cfg, B, S, seed=5)  # the same global inputs on every rank; each uses its block
    lo, hi = se.lo, se.hi
    dev = lambda x: torch.from_numpy(numpy.ascontiguousarray(x[lo:hi])).to(device)  # noqa: E731
    # the gather as peer stores into symmetric buffers (tsp_peer_put + barrier) or as ncclAllGather
    assert (se.pack.p2p is not None) == bool(p2p), "symmetric memory unavailable between these GPUs" if p2p else "unexpected"
    for _ in range(3):  # (several searches: the gathered buffers are double-buffered)
        res = se.run_batch(S, dev(obs), noise=dev(kw["noise"]), tie_stream=dev(kw["tie"]),
                           chance_stream=dev(kw["chance"]) if tsp.variant_of(cfg) else None)
    numpy.savez(os.path.join(out_dir, "rank%d.npz" % rank), lo=lo, hi=hi,
                **{k: v.cpu().numpy() for k, v in res.items()})
    se.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("game,B,S,p2p", [("roundabout_env_ttc_flat", 1001, 20, True), ("roundabout_env_ttc_flat", 1001, 20, False),
                                          ("roundabout_env_ttc_flat_stochastic_concat", 513, 10, True)])
def test_two_nccl_ranks_match_one_gpu_bit_exactly(tmp_path, game, B, S, p2p):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, world, port, game, B, S, str(tmp_path), p2p)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    import tree_search_planning_b200 as tsp
    from tree_search_planning_b200.synthetic import random_state_dict, synthetic_inputs

    cfg = tsp.get_config(game, num_simulations=S)
    eng = tsp.Engine(cfg, max_roots=B, max_simulations=S, state_dict=random_state_dict(cfg, 3))
    obs, kw = synthetic_inputs(cfg, B, S, seed=5)
    want = eng.run(S, observations=obs, **kw)
    eng.close()
    for r in range(world):
        z = numpy.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        assert z["visit_counts"].shape == (B, len(cfg.action_space))
        assert numpy.array_equal(z["visit_counts"], want["visit_counts"])
        assert numpy.array_equal(z["root_value"].view(numpy.int64), want["root_value"].view(numpy.int64))
