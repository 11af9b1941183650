"""Batched self-play driver: the per-step loop of SelfPlay.play_game
(muzero-general/self_play.py:115-220) for B games at once, with the root I/O formats on the device.

Per environment step and for all B games together:
  1. the newest observation frame of every game is uploaded to a device-resident frame pool
     (B x frame_size floats -- the only host -> device traffic of a step);
  2. `tsp_stack_observations` builds GameHistory.get_stacked_observations(-1, stacked_observations)
     (self_play.py:146, 587-624) on the device;
  3. `tsp_run` searches every root (device buffers, no copy);
  4. `tsp_select_action` samples one action per game from the root visit counts with the game-file
     temperature (self_play.py:178, 277-299), `tsp_search_statistics` turns the visit counts into the
     policy target that GameHistory.store_search_statistics appends (self_play.py:205, 570-585);
  5. one action per game (+ the statistics for the histories) comes back to the host; the games step.

The environments are whatever the caller supplies (the reference's Game protocol: reset() -> observation,
step(action) -> (observation, reward, done), legal_actions(), to_play()); gym / highway-env are not part of
this package. Random sources: Dirichlet noise, tie / chance streams and the action sampling (one uniform sample per
game replaces numpy.random.choice(actions, p=...)) are all drawn from ONE numpy Generator seeded by `seed`
(seed=None: fresh OS entropy), so a seeded driver is reproducible. The temperature follows the game file's
visit_softmax_temperature_fn / temperature_threshold like play_game (self_play.py:171-185) when they are given.
Finished games are not searched again: their rows keep a one-action legal list and their results are ignored.
"""
import numpy

from .configs import observation_dim, variant_of
from .engine import Engine, MEM_DEVICE

_STREAM_MOD = 240


class GameHistory:
    """The fields of self_play.GameHistory (self_play.py:552-568) that the planner side fills."""

    def __init__(self):
        self.observation_history = []
        self.action_history = []
        self.reward_history = []
        self.to_play_history = []
        self.child_visits = []
        self.root_values = []


class BatchedSelfPlay:
    def __init__(self, config, state_dict, games, device=0, temperature=1.0, add_exploration_noise=True, seed=None,
                 temperature_fn=None, temperature_threshold=None, trained_steps=0):
        """`temperature`: fixed sampling temperature, unless `temperature_fn(trained_steps)` is given (the game file's
        visit_softmax_temperature_fn); `temperature_threshold`: moves after which play is greedy (self_play.py:181-184)."""
        import torch

        self.torch = torch
        self.config = config
        self.games = list(games)
        self.B = len(self.games)
        self.A = len(config.action_space)
        self.so = int(config.stacked_observations)
        c, h, w = config.observation_shape
        self.F, self.HW = c * h * w, h * w
        self.obs_dim = observation_dim(config)
        self.temperature = float(temperature if temperature_fn is None else temperature_fn(trained_steps))
        self.temperature_threshold = temperature_threshold
        self.moves = 0
        self.add_noise = bool(add_exploration_noise)
        self.rng = numpy.random.default_rng(seed)
        self.variant = variant_of(config)
        self.engine = Engine(config, max_roots=self.B, device=device, state_dict=state_dict)
        self.dev = torch.device("cuda", device)
        B, A, so = self.B, self.A, self.so
        S = int(config.num_simulations)
        # device state: ring of the last so + 1 frames of every game, index / action-plane tables, search I/O
        self.hist = so + 1
        self.pool = torch.zeros((B * self.hist, self.F), dtype=torch.float32, device=self.dev)
        self.frame_index = torch.zeros((B, so + 1), dtype=torch.int32, device=self.dev)
        self.action_value = torch.zeros((B, max(so, 1)), dtype=torch.float32, device=self.dev)
        self.obs = torch.zeros((B, self.obs_dim), dtype=torch.float32, device=self.dev)
        self.d_noise = torch.zeros((B, A), dtype=torch.float64, device=self.dev)
        self.d_tie = torch.zeros((B, 32), dtype=torch.uint8, device=self.dev)
        self.d_chance = torch.zeros((B, max(64, 8 * S)), dtype=torch.uint8, device=self.dev)
        self.d_legal = torch.zeros((B, A), dtype=torch.int32, device=self.dev)
        self.d_nlegal = torch.zeros((B,), dtype=torch.int32, device=self.dev)
        self.d_uniform = torch.zeros((B,), dtype=torch.float64, device=self.dev)
        self.d_visits = torch.zeros((B, A), dtype=torch.int32, device=self.dev)
        self.d_rootv = torch.zeros((B,), dtype=torch.float64, device=self.dev)
        self.d_action = torch.zeros((B,), dtype=torch.int32, device=self.dev)
        self.d_child_visits = torch.zeros((B, A), dtype=torch.float64, device=self.dev)
        self.h_frame = torch.zeros((B, self.F), dtype=torch.float32).pin_memory()
        self.h_action = torch.zeros((B,), dtype=torch.int32).pin_memory()
        self.h_child_visits = torch.zeros((B, A), dtype=torch.float64).pin_memory()
        self.h_rootv = torch.zeros((B,), dtype=torch.float64).pin_memory()
        ins = dict(observations=self.obs, noise=self.d_noise, tie_stream=self.d_tie, legal_actions=self.d_legal,
                   num_legal=self.d_nlegal)
        if self.variant != 0:
            ins["chance_stream"] = self.d_chance
        self.args = self.engine.make_args(B, S, MEM_DEVICE, ins, dict(visit_counts=self.d_visits, root_value=self.d_rootv),
                                          add_exploration_noise=self.add_noise, tie_len=32,
                                          chance_len=self.d_chance.shape[1] if self.variant != 0 else 0)
        self.stream = torch.cuda.ExternalStream(self.engine.stream(), device=self.dev)
        self.histories = [GameHistory() for _ in range(B)]
        self.length = numpy.zeros(B, numpy.int64)  # frames seen per game
        self.done = numpy.zeros(B, bool)
        for b, g in enumerate(self.games):
            self._append(b, g.reset(), 0, 0.0, g.to_play())

    # -- histories -----------------------------------------------------------------------------
    def _append(self, b, observation, action, reward, to_play):
        h = self.histories[b]
        h.observation_history.append(numpy.asarray(observation, numpy.float32))
        h.action_history.append(int(action))
        h.reward_history.append(reward)
        h.to_play_history.append(to_play)
        self.h_frame[b] = self.torch.from_numpy(h.observation_history[-1].reshape(-1))
        self.length[b] += 1

    def _tables(self):
        """frame_index / action_value of get_stacked_observations(-1, so) for every game (self_play.py:598-617)."""
        B, so, hist = self.B, self.so, self.hist
        idx = numpy.full((B, so + 1), -1, numpy.int32)
        act = numpy.zeros((B, max(so, 1)), numpy.float32)
        for b in range(B):
            n = int(self.length[b])
            last = n - 1
            idx[b, 0] = b * hist + last % hist
            ah = self.histories[b].action_history
            for i in range(1, so + 1):
                past = last - i
                if past >= 0:
                    idx[b, i] = b * hist + past % hist
                    act[b, i - 1] = ah[past + 1]
        return idx, act

    # -- one environment step of every live game ---------------------------------------------------
    def step(self):
        torch = self.torch
        B, A = self.B, self.A
        live = ~self.done
        idx, act = self._tables()
        legal = numpy.zeros((B, A), numpy.int32)
        nlegal = numpy.zeros(B, numpy.int32)
        noise = numpy.zeros((B, A), numpy.float64)
        rng = self.rng
        for b, g in enumerate(self.games):
            if not live[b]:  # a finished game is not asked for legal actions again; its row is a dummy
                legal[b, 0] = self.config.action_space[0]
                nlegal[b] = 1
                noise[b, 0] = 1.0
                continue
            la = list(g.legal_actions())
            assert la, f"Legal actions should not be an empty array. Got {la}."
            assert set(la).issubset(set(self.config.action_space)), "Legal actions should be a subset of the action space."
            legal[b, :len(la)] = la
            nlegal[b] = len(la)
            if self.add_noise:
                noise[b, :len(la)] = rng.dirichlet([self.config.root_dirichlet_alpha] * len(la))
        newest = torch.from_numpy(idx[:, 0].astype(numpy.int64)).to(self.dev)
        with torch.cuda.stream(self.stream):
            self.pool.index_copy_(0, newest, self.h_frame.to(self.dev, non_blocking=True))
            self.frame_index.copy_(torch.from_numpy(idx), non_blocking=True)
            self.action_value.copy_(torch.from_numpy(act), non_blocking=True)
            self.d_legal.copy_(torch.from_numpy(legal), non_blocking=True)
            self.d_nlegal.copy_(torch.from_numpy(nlegal), non_blocking=True)
            self.d_noise.copy_(torch.from_numpy(noise), non_blocking=True)
            self.d_tie.copy_(torch.from_numpy(rng.integers(0, _STREAM_MOD, size=(B, 32)).astype(numpy.uint8)),
                             non_blocking=True)
            if self.variant != 0:
                self.d_chance.copy_(torch.from_numpy(rng.integers(
                    0, _STREAM_MOD, size=tuple(self.d_chance.shape)).astype(numpy.uint8)), non_blocking=True)
            self.d_uniform.copy_(torch.from_numpy(rng.random(B)), non_blocking=True)
            eng = self.engine
            eng.stack_observations(self.pool, self.frame_index, self.action_value, self.obs, self.so, self.F, self.HW)
            eng.run_raw(self.args)
            temperature = self.temperature  # self_play.py:178-185
            if self.temperature_threshold is not None and self.moves >= self.temperature_threshold:
                temperature = 0.0
            eng.select_action(self.d_visits, temperature, self.d_uniform, self.d_legal, self.d_nlegal,
                              out=self.d_action)
            eng.search_statistics(self.d_visits, out=self.d_child_visits)
            self.h_action.copy_(self.d_action, non_blocking=True)
            self.h_child_visits.copy_(self.d_child_visits, non_blocking=True)
            self.h_rootv.copy_(self.d_rootv, non_blocking=True)
        # waits for the engine's stream AND reads the tcgen05 status flag: a search whose mbarrier wait timed out raises
        # here instead of feeding invalid actions into the histories
        self.engine.synchronize()
        self.moves += 1
        actions = self.h_action.numpy().copy()
        for b, g in enumerate(self.games):
            if not live[b]:
                continue
            h = self.histories[b]
            h.child_visits.append(self.h_child_visits[b].tolist())  # store_search_statistics, self_play.py:205
            h.root_values.append(float(self.h_rootv[b]))
            observation, reward, done = g.step(int(actions[b]))
            self._append(b, observation, actions[b], reward, g.to_play())
            self.done[b] = bool(done)
        return actions

    def play(self, max_moves=None):
        max_moves = int(self.config.max_moves if max_moves is None else max_moves)
        moves = 0
        while not self.done.all() and moves < max_moves:
            self.step()
            moves += 1
        return self.histories

    def close(self):
        # release every tensor that was used on the engine's stream BEFORE the stream is destroyed (torch's pinned
        # allocator records an event on the streams a block was used on when the block is freed)
        self.stream.synchronize()
        self.args = None
        for name, value in list(vars(self).items()):
            if self.torch.is_tensor(value):
                delattr(self, name)
        import gc
        gc.collect()
        self.torch.cuda.synchronize(self.dev)
        self.stream = None
        self.engine.close()


def reanalyse_values(engine, config, histories, device=0):
    """Reanalyse.reanalyse's value refresh (replay_buffer.py:386-406) for whole games at once: every position of
    every history is stacked on the device (get_stacked_observations(i, stacked_observations)) and goes through
    ONE batched initial_inference + support_to_scalar. Returns a list with one float32 array of
    `reanalysed_predicted_root_values` per history."""
    import torch

    so = int(config.stacked_observations)
    c, h, w = config.observation_shape
    F, HW = c * h * w, h * w
    dev = torch.device("cuda", device)
    frames, index, action, owner = [], [], [], []
    base = 0
    for g, hist in enumerate(histories):
        n = len(hist.root_values)
        frames.extend(numpy.asarray(o, numpy.float32).reshape(-1) for o in hist.observation_history)
        for i in range(n):
            row = [-1] * (so + 1)
            act = [0.0] * max(so, 1)
            row[0] = base + i
            for k in range(1, so + 1):
                if i - k >= 0:
                    row[k] = base + i - k
                    act[k - 1] = float(hist.action_history[i - k + 1])
            index.append(row)
            action.append(act)
            owner.append(g)
        base += len(hist.observation_history)
    if not index:
        return [numpy.zeros(0, numpy.float32) for _ in histories]
    N = len(index)
    pool = torch.from_numpy(numpy.stack(frames)).to(dev)
    d_index = torch.tensor(index, dtype=torch.int32, device=dev)
    d_action = torch.tensor(action, dtype=torch.float32, device=dev)
    obs = torch.empty((N, observation_dim(config)), dtype=torch.float32, device=dev)
    values = torch.empty((N,), dtype=torch.float32, device=dev)
    torch.cuda.synchronize(dev)
    engine.stack_observations(pool, d_index, d_action, obs, so, F, HW)
    engine.initial_inference_device(obs, value=values)
    engine.synchronize()
    v = values.cpu().numpy()
    owner = numpy.asarray(owner)
    return [v[owner == g] for g in range(len(histories))]
