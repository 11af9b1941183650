"""Developer probe: does a concurrent host->device copy slow the (latency-bound) search kernel down?
Device-resident searches on the engine's stream; on a second stream a pinned-host -> device copy of `mb` MB per step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tree_search_planning_b200 as tsp
from tree_search_planning_b200.engine import MEM_DEVICE
from tree_search_planning_b200.synthetic import random_state_dict, synthetic_inputs

B, S = 4096, 50
cfg = tsp.get_config("roundabout_env_ttc_flat", num_simulations=S)
eng = tsp.Engine(cfg, max_roots=B, max_simulations=S, state_dict=random_state_dict(cfg, 0))
obs, kw = synthetic_inputs(cfg, B, S, seed=1)
dev = torch.device("cuda:0")
ins = dict(observations=torch.from_numpy(obs).to(dev), noise=torch.from_numpy(kw["noise"]).to(dev), tie_stream=torch.from_numpy(kw["tie"]).to(dev))
outs = dict(visit_counts=torch.zeros((B, 5), dtype=torch.int32, device=dev), root_value=torch.zeros((B,), dtype=torch.float64, device=dev))
a = eng.make_args(B, S, MEM_DEVICE, ins, outs, add_exploration_noise=True, tie_len=kw["tie"].shape[1], chance_len=0)
stream = torch.cuda.ExternalStream(eng.stream(), device=dev)
side = torch.cuda.Stream(device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for mb, direction in [(0, "-"), (8, "h2d"), (64, "h2d"), (8, "d2h"), (0, "-"), (8, "h2d")]:
    hostbuf = torch.empty(max(mb, 1) << 20, dtype=torch.uint8).pin_memory()
    devbuf = torch.empty(max(mb, 1) << 20, dtype=torch.uint8, device=dev)
    for rep in range(2):
        torch.cuda.synchronize()
        eng.reset_timing()
        for i in range(100):
            with torch.cuda.stream(stream):
                flush.fill_(i & 1)
                ev = torch.cuda.Event()
                ev.record()
            if mb:
                with torch.cuda.stream(side):
                    side.wait_event(ev)  # start the copy when the flush is done, i.e. together with the root / search kernels
                    if direction == "h2d":
                        devbuf.copy_(hostbuf, non_blocking=True)
                    else:
                        hostbuf.copy_(devbuf, non_blocking=True)
            with torch.cuda.stream(stream):
                eng.run_raw(a)
        torch.cuda.synchronize()
        st = eng.stats()
        print("concurrent copy %3d MB %s: search kernel %.4f ms, root %.4f ms" % (mb, direction, st["sum_search_ms"] / st["timed_runs"], st["sum_root_ms"] / st["timed_runs"]))
