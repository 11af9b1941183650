"""Synthetic weights and inputs of the reference shapes (the repo checkpoints are Git-LFS pointer
stubs, so benchmarks and large parity runs use same-shape random-init networks and say so)."""
import numpy

from .configs import observation_dim, variant_of


def random_state_dict(config, seed=0):
    """Random-init weights with the reference's key names (models.py:156-195, 502-513) and
    torch.nn.Linear's default init range U(-1/sqrt(fan_in), 1/sqrt(fan_in))."""
    rng = numpy.random.default_rng(seed)
    A, H, V = len(config.action_space), config.encoding_size, 2 * config.support_size + 1
    stochastic = variant_of(config) != 0
    separate = stochastic and getattr(config, "network", "") == "stochastic"  # MuZeroStochastic: one dynamics MLP per future
    nets = {
        "representation_network": [observation_dim(config)] + list(config.fc_representation_layers) + [H],
        "dynamics_encoded_state_network":
            [H + A + (config.n_futures if stochastic and not separate else 0)] + list(config.fc_dynamics_layers) + [H],
        "dynamics_reward_network": [H] + list(config.fc_reward_layers) + [V],
        "dynamics_terminal_network": [H] + list(config.fc_reward_layers) + [1],
        "prediction_policy_network": [H] + list(config.fc_policy_layers) + [A],
        "prediction_value_network": [H] + list(config.fc_value_layers) + [V],
    }
    if stochastic:
        nets["transition_prior_network"] = [H] + list(config.fc_prior_layers) + [config.n_futures]
    if separate:  # models.py:334-345: keys dynamics_encoded_state_network.<z>.module.<i>
        sizes = nets.pop("dynamics_encoded_state_network")
        for z in range(config.n_futures):
            nets["dynamics_encoded_state_network.%d" % z] = sizes
    sd = {}
    for name, sizes in nets.items():
        for i in range(len(sizes) - 1):
            bound = 1.0 / numpy.sqrt(sizes[i])
            sd["%s.module.%d.weight" % (name, 2 * i)] = rng.uniform(
                -bound, bound, (sizes[i + 1], sizes[i])).astype(numpy.float32)
            sd["%s.module.%d.bias" % (name, 2 * i)] = rng.uniform(-bound, bound, sizes[i + 1]).astype(numpy.float32)
    return sd


def synthetic_inputs(config, num_roots, num_simulations, seed=0, ragged_legal=False, tie_len=32):
    """Seeded observations + external random streams for `num_roots` roots."""
    rng = numpy.random.default_rng(seed)
    A = len(config.action_space)
    B = int(num_roots)
    obs = rng.random((B, observation_dim(config)), dtype=numpy.float32)
    kw = dict(noise=rng.dirichlet([config.root_dirichlet_alpha] * A, size=B),
              tie=rng.integers(0, 240, size=(B, tie_len)).astype(numpy.uint8),
              chance=rng.integers(0, 240, size=(B, 8 * max(1, int(num_simulations)))).astype(numpy.uint8))
    if ragged_legal:
        la = numpy.zeros((B, A), numpy.int32)
        nl = rng.integers(1, A + 1, size=B).astype(numpy.int32)
        for b in range(B):
            la[b, :nl[b]] = rng.permutation(A)[:nl[b]]
            nz = rng.dirichlet([config.root_dirichlet_alpha] * nl[b]) if nl[b] > 1 else numpy.ones(1)
            kw["noise"][b] = 0
            kw["noise"][b, :nl[b]] = nz
        kw.update(legal_actions=la, num_legal=nl)
    return obs, kw
