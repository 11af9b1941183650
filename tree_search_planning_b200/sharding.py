"""Multi-GPU plumbing: independent roots shard across ranks (one process per GPU).

Each root's search is independent (own MinMaxStats, own nodes; self_play.py:378), so there is
no collective inside the simulation loop. `torch.distributed` (NCCL over NVLink on GPUs, gloo in
the CPU tests) is used for exactly two things:
  * broadcast of the flattened network weights from rank 0 (once per weight refresh -- the
    reference ships pickled CPU state dicts through the Ray object store, self_play.py:43);
  * all-gather of the root visit distributions / root values at the end of a search.
"""
import numpy


def shard_bounds(num_roots, world_size, rank):
    """Contiguous block of roots owned by `rank`: [lo, hi). Blocks differ by at most one root."""
    base, rem = divmod(int(num_roots), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def flatten_state_dict(state_dict):
    """-> (float32 vector, [(key, shape)]) in a deterministic key order."""
    keys = sorted(state_dict.keys())
    meta, parts = [], []
    for k in keys:
        v = state_dict[k]
        if hasattr(v, "detach"):
            v = v.detach().cpu().numpy()
        v = numpy.asarray(v, numpy.float32)
        meta.append((k, tuple(v.shape)))
        parts.append(v.reshape(-1))
    flat = numpy.concatenate(parts) if parts else numpy.zeros(0, numpy.float32)
    return flat, meta


def unflatten_state_dict(flat, meta):
    out, pos = {}, 0
    for k, shape in meta:
        n = int(numpy.prod(shape)) if shape else 1
        out[k] = numpy.asarray(flat[pos:pos + n], numpy.float32).reshape(shape).copy()
        pos += n
    return out


def broadcast_state_dict(state_dict, device, src=0):
    """Rank `src` passes its state dict, every other rank passes None; all ranks get the dict.
    The tensor data crosses as ONE flat float32 buffer (NCCL broadcast when `device` is cuda)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return state_dict
    rank = dist.get_rank()
    if rank == src:
        flat, meta = flatten_state_dict(state_dict)
        box = [meta]
    else:
        flat, box = None, [None]
    dist.broadcast_object_list(box, src=src)
    meta = box[0]
    total = sum(int(numpy.prod(s)) if s else 1 for _, s in meta)
    buf = torch.empty(total, dtype=torch.float32, device=device)
    if rank == src:
        buf.copy_(torch.from_numpy(flat))
    dist.broadcast(buf, src=src)
    return unflatten_state_dict(buf.cpu().numpy(), meta)


def all_gather_roots(local, world_size=None):
    """All-gather per-root result tensors of equal shard size along dim 0.
    `local`: dict name -> torch tensor [B_local, ...]; returns dict name -> [world * B_local, ...]."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(local)
    world = dist.get_world_size() if world_size is None else world_size
    out = {}
    for k, t in local.items():
        g = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(g, t.contiguous())
        out[k] = g
    return out


def sharded_search(search_fn, observations, per_root_inputs, world_size, rank, gather=True):
    """Run `search_fn(obs_shard, **input_shards) -> dict of numpy arrays [B_local, ...]` on this rank's
    contiguous block of roots and (optionally) all-gather the root statistics.

    Shards of unequal size are padded to the largest one for the collective and trimmed after."""
    import torch
    import torch.distributed as dist

    B = observations.shape[0]
    lo, hi = shard_bounds(B, world_size, rank)
    shard_inputs = {k: (None if v is None else v[lo:hi]) for k, v in per_root_inputs.items()}
    res = search_fn(observations[lo:hi], **shard_inputs)
    if not gather or world_size == 1 or not (dist.is_available() and dist.is_initialized()):
        return res, (lo, hi)
    max_local = (B + world_size - 1) // world_size
    out = {}
    for k, v in res.items():
        v = numpy.asarray(v)
        pad = numpy.zeros((max_local,) + v.shape[1:], v.dtype)
        pad[:v.shape[0]] = v
        t = torch.from_numpy(pad)
        g = [torch.empty_like(t) for _ in range(world_size)]
        dist.all_gather(g, t)
        pieces = []
        for r in range(world_size):
            rlo, rhi = shard_bounds(B, world_size, r)
            pieces.append(g[r].numpy()[:rhi - rlo])
        out[k] = numpy.concatenate(pieces, axis=0)
    return out, (lo, hi)
