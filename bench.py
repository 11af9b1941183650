#!/usr/bin/env python
"""bench.py -- MCTS simulations/sec of the batched MuZero tree search on B200.

Contract (driver):  python bench.py --gpus N --steps K --warmup W        (torchrun for N > 1)
                    python bench.py --impl reference --gpus N --steps K --warmup W

One "step" = one full batched search (root inference + S simulations) of the workload
BASELINE.json's metric is quoted on: roundabout_env_ttc_flat, 4,096 roots x 50 simulations per
GPU (configs[1]); weak scaling: every rank searches its own 4,096 roots.

  value   roots*sims / device time with inputs and outputs resident in HBM (CUDA events on the
          engine's stream; for N > 1 the NCCL all-gather of the root visit distributions is inside
          the timed region); L2 is flushed between timed steps.
  e2e     the same metric through the C ABI call with HOST (pinned) buffers: the host->device copy
          of observations / noise / tie stream and the device->host read of the root statistics are
          inside the timed region.
  roofline  dominant kernel = tsp::search_tc_kernel (MLP on tcgen05, weights in tensor memory; cross-merge rows:
          tsp::search_tcs_kernel, weights streamed by TMA; tsp::search_fast_kernel / tsp::search_kernel for networks
          outside both envelopes); `kernels` adds the root kernel (tsp::root_tcs_kernel); achieved = algorithmic bytes per launch
          ((108*D + 8*H + 112) B per simulation, SURVEY 8d, D = measured mean selection depth)
          / its mean launch duration (CUDA events recorded by the engine around every launch).
  cpu_baseline  the CPU oracle (C restatement of the reference algorithm, oracle/) timed on the
          host cores of this box on a bounded sample of the same workload (rank 0, N = 1 only);
          cpu_baseline.python_reference = the unmodified Python reference measured in the build container
          (profiles/python_reference_rate.json; /root/reference does not exist on the GPU box).
  configs   the other BASELINE.json configurations, same measurement, one row each (N = 1), and the
          STRONG-scaling rows (every N): the north-star headline (highway net, 65,536 roots) and config 5
          (cross-merge 65,536 roots x 200 sims) with the roots split B/N per rank (sharding.shard_bounds)
          and the one all-gather of the root statistics inside the timed region.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "mcts_simulations_per_sec"
UNIT = "sims/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--game", default="roundabout_env_ttc_flat")
    ap.add_argument("--roots", type=int, default=4096, help="roots per GPU")
    ap.add_argument("--sims", type=int, default=50)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU-baseline sample duration")
    ap.add_argument("--no-configs", action="store_true", help="only the headline workload (skip the other BASELINE configs)")
    ap.add_argument("--only", default="", help="developer: comma-separated tags of the extra rows to run")
    return ap.parse_args()


def workload_name(game, roots, sims):
    return "%s MuZero MCTS batched %s roots x %d sims per GPU" % (game, format(roots, ","), sims)


def config_dict(game, roots, sims):
    """The `config` object: identical in both arms (the driver compares them)."""
    return {"workload": workload_name(game, roots, sims), "game": game, "roots_per_gpu": roots, "num_simulations": sims,
            "network": "fullyconnected, random-init same shapes (repo checkpoints are Git-LFS stubs)",
            "arithmetic": "float32 network, float64 tree statistics"}


# The other BASELINE.json configurations (one GPU each; tag, game, roots, sims, timed steps).
EXTRA_CONFIGS = [
    ("config3_stochastic", "roundabout_env_ttc_flat_stochastic_concat", 16384, 25, 10),
    ("config4_risk_sensitive", "roundabout_env_ttc_flat_risk_sensitive", 16384, 25, 10),
    ("highway_65536x50", "highway_env_ttc_flat", 65536, 50, 5),
    ("highway_8192x200", "highway_env_ttc_flat", 8192, 200, 5),
    ("cross_merge_8192x200", "cross_merge_env", 8192, 200, 3),
]
# Strong scaling: the TOTAL number of roots is fixed and split over the ranks.
STRONG_CONFIGS = [
    ("headline_highway_65536x200", "highway_env_ttc_flat", 65536, 200, 3),
    ("config5_cross_merge_65536x200", "cross_merge_env", 65536, 200, 2),
]


# ------------------------------------------------------------------ clocks ----
_REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
            0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reason_bits, self.power = [], 0, []
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:  # pragma: no cover
            log("clock sampling unavailable:", e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._stop_evt.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.reason_bits |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.ok:
            self.join(timeout=2)
        reasons = [name for bit, name in _REASONS.items() if self.reason_bits & bit]
        return {
            "sm_mhz": float(numpy.median(self.samples)) if self.samples else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": reasons,
            "samples": len(self.samples),
            "power_w_max": max(self.power) if self.power else None,
        }


# ------------------------------------------------------------ CPU baseline ----
def cpu_oracle_rate(cfg, sd, roots, sims, seed, threads=0):
    """Time the CPU oracle (all host threads by default) on `roots` roots; returns (sims/s, seconds, threads)."""
    from oracle.oracle import Oracle, max_threads
    from tree_search_planning_b200.synthetic import synthetic_inputs

    orc = Oracle(cfg, sd)
    obs, kw = synthetic_inputs(cfg, roots, sims, seed)
    t0 = time.perf_counter()
    out = orc.run(sims, observations=obs, n_threads=threads, **kw)
    dt = time.perf_counter() - t0
    assert int(out["visit_counts"].sum()) == roots * sims
    return roots * sims / dt, dt, (threads or max_threads())


def cpu_baseline(cfg, sd, sims, target_seconds):
    rate, _dt, threads = cpu_oracle_rate(cfg, sd, 512, sims, seed=7)  # calibration
    roots = int(max(512, min(4_000_000, rate * target_seconds / sims)))
    rate, dt, threads = cpu_oracle_rate(cfg, sd, roots, sims, seed=8)
    return {
        "value": rate, "unit": UNIT, "cores": threads, "kind": "port",
        "sample": "%d roots x %d sims of the same workload, C oracle (oracle/tsp_oracle.c), %d host threads, %.1f s"
                  % (roots, sims, threads, dt),
    }


def run_reference_arm(args, rank):
    """--impl reference: the reference algorithm's CPU implementation on the host cores. The reference
    itself is Python + torch-CPU and /root/reference does not exist on the GPU box, so this arm times
    the C restatement (oracle/) with every host thread -- a far FASTER stand-in than the real thing
    (the unmodified Python reference measured ~1.0e3 sims/s per core)."""
    if rank != 0:
        return
    from tree_search_planning_b200.configs import get_config
    from tree_search_planning_b200.synthetic import random_state_dict

    cfg = get_config(args.game, num_simulations=args.sims)
    sd = random_state_dict(cfg, 0)
    rate, _dt, threads = cpu_oracle_rate(cfg, sd, 512, args.sims, seed=7)
    budget = 150.0 / max(1, args.steps + args.warmup)  # whole run within a few minutes
    roots = int(max(64, min(args.roots, rate * min(2.0, budget) / args.sims)))
    for i in range(args.warmup):
        cpu_oracle_rate(cfg, sd, roots, args.sims, seed=100 + i)
    t_total = 0.0
    for i in range(args.steps):
        _r, dt, _t = cpu_oracle_rate(cfg, sd, roots, args.sims, seed=200 + i)
        t_total += dt
    value = roots * args.sims * args.steps / t_total
    cb = {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
          "sample": "each step = %d roots x %d sims (bounded sample of the %d-root workload), C oracle, %d host threads"
                    % (roots, args.sims, args.roots, threads)}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64", "data": "synthetic",
        "config": config_dict(args.game, args.roots, args.sims),
        "run": {"roots_per_step": roots},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------- GPU arm ----
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def profiled_traffic(tag):
    """dram bytes per launch of the search kernel from the committed ncu summary, if there is one."""
    p = os.path.join(ROOT, "profiles", "search_kernel_traffic.json")
    try:
        d = json.load(open(p))
        return d.get(tag)
    except Exception:
        return None


def time_config(torch, dist, se, cfg, S, steps, warmup, device, world, rank, host_mode, flush, seed):
    """Time `steps` searches of this rank's block of roots on the sharded engine `se`; returns
    (seconds summed over the steps [max over ranks], extras). The one all-gather of the root statistics
    (sharding.PackedGather) is inside the timed region when world > 1."""
    from tree_search_planning_b200.engine import MEM_DEVICE, MEM_HOST
    from tree_search_planning_b200.synthetic import synthetic_inputs
    from tree_search_planning_b200.configs import variant_of

    eng = se.engine
    B = se.n_local
    A = len(cfg.action_space)
    obs, kw = synthetic_inputs(cfg, B, S, seed=seed + rank)
    stochastic = variant_of(cfg) != 0
    host = dict(observations=torch.from_numpy(obs), noise=torch.from_numpy(kw["noise"]),
                tie_stream=torch.from_numpy(kw["tie"]))
    if stochastic:
        host["chance_stream"] = torch.from_numpy(kw["chance"])
    extra_shapes = dict(max_tree_depth=((B,), torch.int32), root_predicted_value=((B,), torch.float32),
                        total_select_depth=((B,), torch.int32))
    packed = dict(visit_counts=((B, A), torch.int32), root_value=((B,), torch.float64))
    gather = world > 1
    host_all = None
    if host_mode:
        ins = {k: v.pin_memory() for k, v in host.items()}
        if gather:
            # N > 1: HOST inputs, the root statistics stay on the device for the all-gather, and the gathered
            # buffer of all ranks is read back to pinned host memory (outputs_on_device, include/tsp.h)
            outs = {k: torch.zeros(s, dtype=dt, device=device) for k, (s, dt) in extra_shapes.items()}
            host_all = torch.zeros(se.pack.all_bytes.shape, dtype=torch.uint8).pin_memory()
        else:
            outs = {k: torch.zeros(s, dtype=dt).pin_memory() for k, (s, dt) in {**extra_shapes, **packed}.items()}
        mem = MEM_HOST
    else:
        ins = {k: v.to(device) for k, v in host.items()}
        outs = {k: torch.zeros(s, dtype=dt, device=device) for k, (s, dt) in extra_shapes.items()}
        mem = MEM_DEVICE
    kwargs = dict(add_exploration_noise=True, tie_len=kw["tie"].shape[1], chance_len=kw["chance"].shape[1] if stochastic else 0)
    if host_mode and not gather:
        a = eng.make_args(B, S, mem, ins, outs, **kwargs)
    else:
        a = se.make_args(S, mem, ins, extra_outputs=outs, outputs_on_device=host_mode, **kwargs)
    stream = torch.cuda.ExternalStream(eng.stream(), device=device)
    torch.cuda.synchronize(device)

    def step():
        eng.run_raw(a)
        if gather:  # ONE collective: visit counts | root values of every rank over NCCL / NVLink
            se.pack.gather()
            if host_all is not None:
                host_all.copy_(se.pack.all_bytes, non_blocking=True)

    with torch.cuda.stream(stream):
        for _ in range(warmup):
            step()
    torch.cuda.synchronize(device)
    eng.reset_timing()
    launches0 = eng.stats()["kernel_launches"]
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(device)
    wall0 = time.perf_counter()
    with torch.cuda.stream(stream):
        for i in range(steps):
            flush.fill_(i & 1)  # evict the previous step's working set from the 126 MB L2 (untimed)
            ev0[i].record()
            step()
            ev1[i].record()
    torch.cuda.synchronize(device)
    wall = time.perf_counter() - wall0
    if world > 1:
        dist.barrier()
    ms = sum(e0.elapsed_time(e1) for e0, e1 in zip(ev0, ev1))
    t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    st = eng.stats()
    if host_mode and not gather:
        visits = outs["visit_counts"].numpy()
    else:
        visits = se.pack.local["visit_counts"][:B].cpu().numpy()
    assert (visits.sum(1) == S).all(), "search did not complete"
    if gather:  # every rank holds every root's statistics
        allv = se.pack.unpack()["visit_counts"]
        assert int(allv.shape[0]) == se.num_roots and bool((allv.sum(1) == S).all()), "gathered statistics incomplete"
        if host_all is not None:
            assert bool((host_all == se.pack.all_bytes.cpu()).all())
    h2d = sum(v.numel() * v.element_size() for v in ins.values()) if host_mode else 0
    if not host_mode:
        d2h = 0
    elif gather:
        d2h = host_all.numel()
    else:
        d2h = sum(v.numel() * v.element_size() for v in outs.values())
    extras = dict(
        launches=int(st["kernel_launches"] - launches0),
        search_ms=st["sum_search_ms"] / max(1, st["timed_runs"]),
        root_ms=st["sum_root_ms"] / max(1, st["timed_runs"]),
        mean_depth=float(outs["total_select_depth"].sum().item()) / (B * S),
        wall_ms=float(t[1].item()), h2d=h2d, d2h=d2h, stats=st,
    )
    return float(t[0].item()) / 1e3, extras


KERNEL_NAMES = {3: "tsp::search_tcs_kernel", 2: "tsp::search_tc_kernel", 1: "tsp::search_fast_kernel", 0: "tsp::search_kernel"}


def kernel_name(st, variant):
    k = st.get("search_kernel")
    if k == 2 and variant != 0:
        return "tsp::search_tc_variant_kernel"
    if k == 1 and variant != 0:
        return "tsp::search_fast_variant_kernel"
    return KERNEL_NAMES.get(k, "tsp::search_kernel")


def roofline_of(cfg, B, S, ex, peak, peak_src, tag=None):
    """HBM roofline of the search kernel: algorithmic bytes per simulation (108 D + 8 H + 112, D = the
    measured mean selection depth) x the B x S simulations of one launch / its mean duration (CUDA events recorded by the
    engine around every launch)."""
    from tree_search_planning_b200.configs import variant_of
    D = ex["mean_depth"]
    bps = 108.0 * D + 8.0 * cfg.encoding_size + 112.0
    achieved = bps * B * S / (ex["search_ms"] * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": profiled_traffic(tag) if tag else None, "kernel": kernel_name(ex["stats"], variant_of(cfg)),
            "kernel_ms": ex["search_ms"], "algorithmic_bytes_per_sim": bps, "mean_select_depth": D, "peak_source": peak_src,
            "kernel_share_of_step": ex["search_ms"] / (ex["search_ms"] + ex["root_ms"])}


def root_roofline(cfg, B, ex, peak):
    """The root kernel (representation + prediction + tree init): 4 obs_dim + 4 H + 128 algorithmic bytes per root."""
    from tree_search_planning_b200.configs import observation_dim
    bpr = 4.0 * observation_dim(cfg) + 4.0 * cfg.encoding_size + 128.0
    achieved = bpr * B / max(ex["root_ms"] * 1e-3, 1e-9) / 1e9
    name = "tsp::root_tcs_kernel" if ex["stats"].get("root_kernel") == 1 else "tsp::net_kernel (root)"
    return {"bound": "hbm", "kernel": name, "kernel_ms": ex["root_ms"], "achieved": achieved, "peak": peak,
            "unit": "GB/s", "frac": achieved / peak, "algorithmic_bytes_per_root": bpr}


def measure_row(torch, dist, tsp, tag, game, B_total, S, steps, warmup, device, local_rank, world, rank, flush, peak, peak_src,
                scaling, e2e=True):
    """One extra row: the same measurement as the headline for another BASELINE configuration."""
    from tree_search_planning_b200.synthetic import random_state_dict
    from tree_search_planning_b200.sharding import ShardedEngine
    cfg = tsp.get_config(game, num_simulations=S)
    row = {"tag": tag, "game": game, "roots_total": B_total, "num_simulations": S, "scaling": scaling, "steps": steps}
    se = None
    try:
        se = ShardedEngine(cfg, num_roots=B_total, max_simulations=S, device=device,
                           state_dict=random_state_dict(cfg, 0) if rank == 0 else None)
        t, ex = time_config(torch, dist, se, cfg, S, steps, warmup, device, world, rank, False, flush, 2000)
        row.update(value=B_total * S * steps / t, unit=UNIT, ms_per_step=1e3 * t / steps, roots_per_gpu=se.n_local,
                   kernel=kernel_name(ex["stats"], tsp.variant_of(cfg)), kernel_ms=ex["search_ms"], root_ms=ex["root_ms"],
                   trees_per_cta=ex["stats"]["trees_per_cta"], search_ctas=ex["stats"]["search_ctas"],
                   roofline=roofline_of(cfg, se.n_local, S, ex, peak, peak_src, tag="%s_B%d_S%d" % (game, se.n_local, S)),
                   root_roofline=root_roofline(cfg, se.n_local, ex, peak))
        if e2e:
            t2, ex2 = time_config(torch, dist, se, cfg, S, steps, warmup, device, world, rank, True, flush, 2000)
            row["e2e"] = {"value": B_total * S * steps / t2, "unit": UNIT, "ms_per_step": 1e3 * t2 / steps,
                          "wall_ms_per_step": ex2["wall_ms"] / steps, "h2d_bytes_per_step": ex2["h2d"],
                          "d2h_bytes_per_step": ex2["d2h"]}
    except Exception as e:  # pragma: no cover
        row["error"] = "%s: %s" % (type(e).__name__, e)
    finally:
        if se is not None:
            se.close()
    log("row", json.dumps(row))
    return row


_JSON_FD = None


def guard_stdout():
    """Libraries (NCCL prints its version banner) write to stdout; the driver wants exactly ONE JSON line there.
    Everything but that line goes to stderr: fd 1 is pointed at fd 2 and the line is written to the saved fd."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_JSON_FD, data)


def main():
    args = parse_args()
    guard_stdout()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    import tree_search_planning_b200 as tsp
    from tree_search_planning_b200.synthetic import random_state_dict

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    from tree_search_planning_b200.sharding import ShardedEngine
    B, S = args.roots, args.sims
    cfg = tsp.get_config(args.game, num_simulations=S)
    sd = random_state_dict(cfg, 0) if rank == 0 else None
    # weak scaling: every rank searches its own B roots; weights broadcast from rank 0 over NCCL
    se = ShardedEngine(cfg, num_roots=world * B, max_simulations=S, device=device, state_dict=sd)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)

    sampler = ClockSampler(local_rank)
    sampler.start()
    t_dev, ex_dev = time_config(torch, dist, se, cfg, S, args.steps, args.warmup, device, world, rank, False, flush, 1000)
    t_e2e, ex_e2e = time_config(torch, dist, se, cfg, S, args.steps, args.warmup, device, world, rank, True, flush, 1000)
    clocks = sampler.stop()
    se.close()

    units = float(world) * B * S * args.steps
    value = units / t_dev
    e2e = units / t_e2e
    peak, peak_src = measured_peaks()
    roofline = roofline_of(cfg, B, S, ex_dev, peak, peak_src, tag="%s_B%d_S%d" % (args.game, B, S))
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32/f64", "data": "synthetic",
        "config": config_dict(args.game, B, S),
        "run": {"l2": "flushed between timed steps (256 MiB write, untimed)",
                "trees_per_cta": ex_dev["stats"]["trees_per_cta"], "search_ctas": ex_dev["stats"]["search_ctas"],
                "parallelism": "roots sharded, %d rank(s), no collective inside the search; one all-gather of the packed "
                               "root statistics per search" % world,
                "wall_ms_per_step": ex_dev["wall_ms"] / args.steps},
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": ex_e2e["h2d"], "d2h_bytes_per_step": ex_e2e["d2h"],
                "ms_per_step": 1e3 * t_e2e / args.steps, "wall_ms_per_step": ex_e2e["wall_ms"] / args.steps,
                # where an end-to-end step goes (CUDA events of the engine around its launches): `upload_wait_and_root_ms` is
                # the interval from the start of tsp_run's device work to the end of the root kernel, i.e. what is left of
                # the host->device copy after its overlap with the previous search + the root kernel
                "search_kernel_ms": ex_e2e["search_ms"], "upload_wait_and_root_ms": ex_e2e["root_ms"],
                "note": "sum of per-step CUDA-event intervals on the engine's stream; wall_ms_per_step is the host clock over the "
                        "same steps (includes the untimed L2 flushes); the upload of step i + 1 overlaps the search of step i"},
        "gpu_launches": ex_dev["launches"],
        "roofline": roofline,
        "kernels": [roofline, root_roofline(cfg, B, ex_dev, peak)],
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(cfg, sd, S, args.cpu_seconds)
        pr = python_reference_rate()
        if pr:
            line["cpu_baseline"]["python_reference"] = pr
    if not args.no_configs:
        only = set(t for t in args.only.split(",") if t)
        rows = []
        common = (device, local_rank, world, rank, flush, peak, peak_src)
        if world == 1:
            for tag, game, Bx, Sx, steps in EXTRA_CONFIGS:
                if only and tag not in only:
                    continue
                rows.append(measure_row(torch, dist, tsp, tag, game, Bx, Sx, steps, 3, *common, scaling="single-gpu"))
        for tag, game, Bx, Sx, steps in STRONG_CONFIGS:
            if only and tag not in only:
                continue
            # (a rank's share shrinks with the world size: the same timed work per row at every N)
            rows.append(measure_row(torch, dist, tsp, tag, game, Bx, Sx, steps * world, 3, *common, scaling="strong"))
        line["configs"] = rows
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def python_reference_rate():
    """The unmodified Python reference (MCTS.run of muzero-general/self_play.py on torch-CPU), measured in the build
    container where /root/reference exists (oracle/time_python_reference.py); it cannot run on the GPU box."""
    p = os.path.join(ROOT, "profiles", "python_reference_rate.json")
    try:
        return json.load(open(p))
    except Exception:
        return None


if __name__ == "__main__":
    main()
