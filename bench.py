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
  roofline  dominant kernel = tsp::search_tc_kernel (MLP on tcgen05, weights in tensor memory; tsp::search_fast_kernel /
          tsp::search_kernel for networks outside its envelope); achieved = algorithmic bytes per launch
          ((108*D + 8*H + 112) B per simulation, SURVEY 8d, D = measured mean selection depth)
          / its mean launch duration (CUDA events recorded by the engine around every launch).
  cpu_baseline  the CPU oracle (C restatement of the reference algorithm, oracle/) timed on the
          host cores of this box on a bounded sample of the same workload (rank 0, N = 1 only).
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
    ap.add_argument("--sweep", action="store_true", help="also time the other BASELINE configs (stderr + gpurun_out/)")
    return ap.parse_args()


def workload_name(game, roots, sims):
    return "%s MuZero MCTS batched %s roots x %d sims per GPU" % (game, format(roots, ","), sims)


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
        "config": {"workload": workload_name(args.game, args.roots, args.sims), "roots_per_step": roots,
                   "num_simulations": args.sims, "weights": "random-init, same shapes (repo checkpoints are LFS stubs)"},
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


def time_config(torch, dist, eng, cfg, B, S, steps, warmup, device, world, rank, host_mode, flush, seed):
    """Time `steps` searches; returns (seconds summed over steps [max over ranks], extras)."""
    from tree_search_planning_b200.engine import MEM_DEVICE, MEM_HOST
    from tree_search_planning_b200.synthetic import synthetic_inputs
    from tree_search_planning_b200.configs import variant_of

    A = len(cfg.action_space)
    obs, kw = synthetic_inputs(cfg, B, S, seed=seed + rank)
    stochastic = variant_of(cfg) != 0
    host = dict(observations=torch.from_numpy(obs), noise=torch.from_numpy(kw["noise"]),
                tie_stream=torch.from_numpy(kw["tie"]))
    if stochastic:
        host["chance_stream"] = torch.from_numpy(kw["chance"])
    out_shapes = dict(visit_counts=((B, A), torch.int32), root_value=((B,), torch.float64),
                      max_tree_depth=((B,), torch.int32), root_predicted_value=((B,), torch.float32),
                      total_select_depth=((B,), torch.int32))
    if host_mode:
        ins = {k: v.pin_memory() for k, v in host.items()}
        outs = {k: torch.zeros(s, dtype=dt).pin_memory() for k, (s, dt) in out_shapes.items()}
        mem = MEM_HOST
    else:
        ins = {k: v.to(device) for k, v in host.items()}
        outs = {k: torch.zeros(s, dtype=dt, device=device) for k, (s, dt) in out_shapes.items()}
        mem = MEM_DEVICE
    a = eng.make_args(B, S, mem, ins, outs, add_exploration_noise=True, tie_len=kw["tie"].shape[1],
                      chance_len=kw["chance"].shape[1] if stochastic else 0)
    gathered = None
    if world > 1 and not host_mode:
        gathered = {k: torch.empty((world * B,) + tuple(outs[k].shape[1:]), dtype=outs[k].dtype, device=device)
                    for k in ("visit_counts", "root_value")}
    stream = torch.cuda.ExternalStream(eng.stream(), device=device)
    torch.cuda.synchronize(device)

    def step():
        eng.run_raw(a)
        if gathered is not None:  # gather of the root visit distributions over NCCL / NVLink
            for k, g in gathered.items():
                dist.all_gather_into_tensor(g, outs[k])

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
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    st = eng.stats()
    outs_np = {k: v.cpu().numpy() for k, v in outs.items()}
    assert (outs_np["visit_counts"].sum(1) == S).all(), "search did not complete"
    extras = dict(
        launches=int(st["kernel_launches"] - launches0),
        search_ms=st["sum_search_ms"] / max(1, st["timed_runs"]),
        root_ms=st["sum_root_ms"] / max(1, st["timed_runs"]),
        mean_depth=float(outs_np["total_select_depth"].sum()) / (B * S),
        wall_s=wall,
        h2d=sum(v.numel() * v.element_size() for v in ins.values()) if host_mode else 0,
        d2h=sum(v.numel() * v.element_size() for v in outs.values()) if host_mode else 0,
        stats=st,
    )
    return float(t.item()) / 1e3, extras


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
    from tree_search_planning_b200.sharding import broadcast_state_dict

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    B, S = args.roots, args.sims
    cfg = tsp.get_config(args.game, num_simulations=S)
    sd = random_state_dict(cfg, 0) if rank == 0 else None
    sd = broadcast_state_dict(sd, device)  # NCCL broadcast of the flattened weights from rank 0
    eng = tsp.Engine(cfg, max_roots=B, max_simulations=S, device=local_rank, state_dict=sd)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)

    sampler = ClockSampler(local_rank)
    sampler.start()
    t_dev, ex_dev = time_config(torch, dist, eng, cfg, B, S, args.steps, args.warmup, device, world, rank, False, flush, 1000)
    t_e2e, ex_e2e = time_config(torch, dist, eng, cfg, B, S, args.steps, args.warmup, device, world, rank, True, flush, 1000)
    clocks = sampler.stop()

    units = float(world) * B * S * args.steps
    value = units / t_dev
    e2e = units / t_e2e
    H = cfg.encoding_size
    D = ex_dev["mean_depth"]
    bytes_per_sim = 108.0 * D + 8.0 * H + 112.0
    peak, peak_src = measured_peaks()
    achieved = bytes_per_sim * B * S / (ex_dev["search_ms"] * 1e-3) / 1e9
    tag = "%s_B%d_S%d" % (args.game, B, S)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": profiled_traffic(tag),
                "kernel": {2: "tsp::search_tc_kernel", 1: "tsp::search_fast_kernel"}.get(ex_dev["stats"].get("search_kernel"),
                                                                                         "tsp::search_kernel"),
                "kernel_ms": ex_dev["search_ms"], "algorithmic_bytes_per_sim": bytes_per_sim, "mean_select_depth": D,
                "peak_source": peak_src,
                "kernel_share_of_step": ex_dev["search_ms"] / (ex_dev["search_ms"] + ex_dev["root_ms"])}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32/f64", "data": "synthetic",
        "config": {"workload": workload_name(args.game, B, S), "roots_per_gpu": B, "num_simulations": S,
                   "network": "fullyconnected, random-init same shapes (repo checkpoints are Git-LFS stubs)",
                   "arithmetic": "float32 network, float64 tree statistics",
                   "l2": "flushed between timed steps (256 MiB write, untimed)",
                   "trees_per_cta": ex_dev["stats"]["trees_per_cta"], "search_ctas": ex_dev["stats"]["search_ctas"],
                   "parallelism": "roots sharded, %d rank(s), no collective inside the search" % world},
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": ex_e2e["h2d"], "d2h_bytes_per_step": ex_e2e["d2h"],
                "ms_per_step": 1e3 * t_e2e / args.steps},
        "gpu_launches": ex_dev["launches"],
        "roofline": roofline,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(cfg, sd, S, args.cpu_seconds)
    if args.sweep and world == 1:
        sweep(torch, dist, tsp, device, local_rank, flush, line)
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def sweep(torch, dist, tsp, device, local_rank, flush, line):
    """Other BASELINE.json configs (not bench lines): absolute numbers for DESIGN.md / profiles."""
    from tree_search_planning_b200.synthetic import random_state_dict
    rows = []
    for game, B, S, steps in [("highway_env_ttc_flat", 8192, 50, 50), ("highway_env_ttc_flat", 65536, 50, 10),
                              ("highway_env_ttc_flat", 8192, 200, 10),
                              ("roundabout_env_ttc_flat_stochastic_concat", 16384, 25, 20),
                              ("roundabout_env_ttc_flat_risk_sensitive", 16384, 25, 20),
                              ("cross_merge_env", 8192, 200, 5), ("cross_merge_env", 65536, 200, 2)]:
        cfg = tsp.get_config(game, num_simulations=S)
        try:
            eng = tsp.Engine(cfg, max_roots=B, max_simulations=S, device=local_rank, state_dict=random_state_dict(cfg, 0))
            t, ex = time_config(torch, dist, eng, cfg, B, S, steps, 2, device, 1, 0, False, flush, 2000)
            H = cfg.encoding_size
            bps = 108.0 * ex["mean_depth"] + 8.0 * H + 112.0
            row = dict(game=game, roots=B, sims=S, sims_per_s=B * S * steps / t, search_ms=ex["search_ms"],
                       root_ms=ex["root_ms"], mean_depth=ex["mean_depth"],
                       achieved_gbs=bps * B * S / (ex["search_ms"] * 1e-3) / 1e9, trees_per_cta=ex["stats"]["trees_per_cta"])
            eng.close()
        except Exception as e:  # pragma: no cover
            row = dict(game=game, roots=B, sims=S, error=str(e))
        rows.append(row)
        log("sweep", json.dumps(row))
    line["sweep"] = rows


if __name__ == "__main__":
    main()
