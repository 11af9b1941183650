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


class PackedGather:
    """ONE collective per search for the root statistics of all ranks.

    Every rank owns a contiguous block of roots (`shard_bounds`). The per-root result arrays of a rank
    (e.g. visit counts int32 [n, A] and root values float64 [n]) live as VIEWS into one byte buffer
    `[field 0 of max_local roots | field 1 of max_local roots | ...]` (8-byte aligned fields); the engine
    writes its outputs straight into those views, a single `all_gather_into_tensor` of the byte buffers
    moves everything, and `unpack` returns per-field arrays of all `num_roots` roots in root order.
    Works on CUDA tensors over NCCL and on CPU tensors over gloo (the world-size-2 CPU test).

    `engine` (CUDA, world > 1): the gather becomes ONE kernel of peer stores over NVLink / NVSwitch (`tsp_peer_put`: this
    rank's slice into the same offset of every rank's SYMMETRIC buffer, torch.distributed._symmetric_memory) followed by
    the symmetric-memory barrier -- ~12 us where ncclAllGather of these ~100 KB costs 17 (2 GPUs) to ~30 us (8). The
    gathered buffers are double-buffered: a rank may still be reading step i's result while a faster rank already puts
    step i + 1; by the time step i + 2 reuses the buffer, every rank has passed the barrier of step i + 1 behind its reads.
    Falls back to NCCL when symmetric memory is unavailable (no P2P, TSP_NO_P2P_GATHER=1)."""

    def __init__(self, fields, num_roots, world_size, rank, device, engine=None, group=None):
        import torch

        self.torch = torch
        self.B, self.world, self.rank = int(num_roots), int(world_size), int(rank)
        self.lo, self.hi = shard_bounds(self.B, self.world, self.rank)
        self.max_local = (self.B + self.world - 1) // self.world
        self.fields = []
        off = 0
        for name, inner, dtype in fields:
            inner = tuple(int(x) for x in inner)
            n = self.max_local * int(numpy.prod(inner)) if inner else self.max_local
            nbytes = n * torch.empty(0, dtype=dtype).element_size()
            self.fields.append((name, inner, dtype, off, nbytes))
            off += (nbytes + 7) & ~7
        self.chunk = off
        self.chunk = (self.chunk + 15) & ~15  # tsp_peer_put moves 16-byte words
        self.local_bytes = torch.zeros(self.chunk, dtype=torch.uint8, device=device)
        self.all_bytes = torch.zeros(self.world * self.chunk, dtype=torch.uint8, device=device)
        self.local = {name: self._view(self.local_bytes, o, nb, dt, inner) for name, inner, dt, o, nb in self.fields}
        self.p2p = None
        if engine is not None and self.world > 1 and torch.device(device).type == "cuda":
            self._init_p2p(engine, group)

    def _init_p2p(self, engine, group):
        import os
        import torch.distributed as dist
        torch = self.torch
        ok = os.environ.get("TSP_NO_P2P_GATHER", "0") in ("", "0")
        bufs, hdls, ptrs = [], [], []
        if ok:
            try:
                import torch.distributed._symmetric_memory as symm_mem
                for _ in range(2):
                    b = symm_mem.empty(self.world * self.chunk, dtype=torch.uint8, device=self.local_bytes.device)
                    h = symm_mem.rendezvous(b, group if group is not None else dist.group.WORLD)
                    b.zero_()
                    bufs.append(b)
                    hdls.append(h)
                    ptrs.append([int(p) + self.rank * self.chunk for p in h.buffer_ptrs])  # my slice in every rank's buffer
            except Exception:
                ok = False
        # every rank must take the same path: one that cannot use symmetric memory sends all back to NCCL
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.local_bytes.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 1:
            self.p2p = dict(engine=engine, bufs=bufs, hdls=hdls, ptrs=ptrs, flip=0)
            self.all_bytes = bufs[0]

    def _view(self, buf, off, nbytes, dtype, inner):
        return buf[off:off + nbytes].view(dtype).reshape((self.max_local,) + inner)

    def gather(self, group=None):
        """The collective (asynchronous on the current CUDA stream for NCCL)."""
        import torch.distributed as dist

        if self.world == 1 or not (dist.is_available() and dist.is_initialized()):
            self.all_bytes.copy_(self.local_bytes)
        elif self.p2p is not None:
            # (on the CURRENT stream, which callers make the engine's: ShardedEngine.run_args, bench.py)
            q = self.p2p
            k = q["flip"]
            q["flip"] ^= 1
            q["engine"].peer_put(self.local_bytes, self.chunk, q["ptrs"][k])
            q["hdls"][k].barrier(channel=0)
            self.all_bytes = q["bufs"][k]
        else:
            dist.all_gather_into_tensor(self.all_bytes, self.local_bytes, group=group)
        return self.all_bytes

    def unpack(self, all_bytes=None):
        """-> dict name -> tensor [num_roots, ...] in root order (views when the shards are equal)."""
        torch = self.torch
        buf = self.all_bytes if all_bytes is None else all_bytes
        out = {}
        for name, inner, dt, o, nb in self.fields:
            parts = []
            for r in range(self.world):
                rlo, rhi = shard_bounds(self.B, self.world, r)
                parts.append(self._view(buf[r * self.chunk:(r + 1) * self.chunk], o, nb, dt, inner)[:rhi - rlo])
            out[name] = parts[0] if self.world == 1 else torch.cat(parts, 0)
        return out


class ShardedEngine:
    """The multi-GPU product path: one process per GPU (torch.distributed, NCCL), each rank searches its
    contiguous block of the `num_roots` roots on its own engine (own node pools, a full weight copy) and the
    root statistics of all ranks are exchanged with ONE all-gather per search (PackedGather). Weights are
    broadcast from rank `src`. There is no collective inside the simulation loop: roots are independent
    (self_play.py:378).

        se = ShardedEngine(cfg, num_roots=65536, max_simulations=200, device=torch.device("cuda", local_rank),
                           state_dict=sd_on_rank0)
        res = se.run_batch(200, observations=obs_local, noise=noise_local, tie_stream=tie_local)   # device tensors
        res["visit_counts"]  # [65536, A] on every rank
    """

    def __init__(self, config, num_roots, max_simulations, device, state_dict=None, src=0, group=None, engine_cls=None):
        import torch
        import torch.distributed as dist
        from .engine import Engine

        self.torch, self.group = torch, group
        on = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if on else 1
        self.rank = dist.get_rank(group) if on else 0
        self.device = torch.device(device)
        self.config = config
        self.num_roots = int(num_roots)
        self.lo, self.hi = shard_bounds(self.num_roots, self.world, self.rank)
        self.n_local = self.hi - self.lo
        A = len(config.action_space)
        self.A = A
        sd = broadcast_state_dict(state_dict, self.device, src=src)
        cls = Engine if engine_cls is None else engine_cls
        self.engine = cls(config, max_roots=max(1, (self.num_roots + self.world - 1) // self.world),
                          max_simulations=max_simulations, device=self.device.index or 0, state_dict=sd)
        fields = [("visit_counts", (A,), torch.int32), ("root_value", (), torch.float64)]
        self.pack = PackedGather(fields, self.num_roots, self.world, self.rank, self.device, engine=self.engine, group=group)
        self._extra = {}

    def refresh_weights(self, state_dict, src=0):
        """Hot weight refresh on every rank: rank `src` passes the new state dict, the others None."""
        self.engine.load_state_dict(broadcast_state_dict(state_dict, self.device, src=src))

    def make_args(self, num_simulations, memory, inputs, extra_outputs=None, outputs_on_device=False, **kw):
        """tsp_run_args whose visit_counts / root_value outputs are the packed gather buffer's views (DEVICE memory;
        with HOST inputs pass outputs_on_device=True) -- or caller-given HOST outputs otherwise."""
        outs = dict(extra_outputs or {})
        from .engine import MEM_DEVICE
        if memory == MEM_DEVICE or outputs_on_device:
            outs["visit_counts"] = self.pack.local["visit_counts"]
            outs["root_value"] = self.pack.local["root_value"]
        return self.engine.make_args(self.n_local, num_simulations, memory, inputs, outs,
                                     outputs_on_device=outputs_on_device, **kw)

    def run_args(self, args, gather=True):
        """Launch a prepared search on the engine's stream and (optionally) the one all-gather behind it."""
        torch = self.torch
        stream = torch.cuda.ExternalStream(self.engine.stream(), device=self.device)
        with torch.cuda.stream(stream):
            self.engine.run_raw(args)
            if gather:
                self.pack.gather(self.group)

    def run_batch(self, num_simulations, observations, noise=None, tie_stream=None, chance_stream=None,
                  legal_actions=None, num_legal=None, add_exploration_noise=True, gather=True):
        """One sharded search over this rank's block of roots (DEVICE tensors [n_local, ...]); returns the
        root statistics of ALL roots (device tensors) after synchronising the engine's stream."""
        from .engine import MEM_DEVICE
        ins = dict(observations=observations)
        for k, v in (("noise", noise), ("tie_stream", tie_stream), ("chance_stream", chance_stream),
                     ("legal_actions", legal_actions), ("num_legal", num_legal)):
            if v is not None:
                ins[k] = v
        a = self.make_args(num_simulations, MEM_DEVICE, ins, add_exploration_noise=add_exploration_noise and noise is not None,
                           tie_len=0 if tie_stream is None else int(tie_stream.shape[1]),
                           chance_len=0 if chance_stream is None else int(chance_stream.shape[1]))
        self.run_args(a, gather=gather)
        self.engine.synchronize()
        if not gather:
            return {k: v[:self.n_local] for k, v in self.pack.local.items()}
        return self.pack.unpack()

    def close(self):
        self.engine.close()
