"""Search / network hyper-parameters of the reference game files, as plain attribute bags.

The reference keeps these constants in ``MuZeroConfig`` classes inside game files that
import gym / highway_env / rl_agents (absent here), so only the attributes the planning
hot path reads are restated, with the attribute names the reference uses so that the same
object can be handed to the reference ``MCTS(config)`` and to ours.

Sources (reference file:line):
  highway_env_ttc_flat / roundabout_env_ttc_flat:
      muzero-general/games/roundabout_env_ttc_flat.py:30-33,45-58,64,80-86,106
      muzero-general/games/highway_env_ttc_flat.py:29-32,44-57,63,79-85,105
  roundabout_env_ttc_flat_stochastic_concat:
      muzero-general/games/roundabout_env_ttc_flat_stochastic_concat.py:30-33,45-55,60-68,82-87
  roundabout_env_ttc_flat_stochastic / highway_env_ttc_flat_stochastic (MuZeroStochastic: one dynamics MLP per future):
      muzero-general/games/roundabout_env_ttc_flat_stochastic.py:30-33,45-55,60-68,82-87
      muzero-general/games/highway_env_ttc_flat_stochastic.py (same constants)
  roundabout_env_ttc_flat_risk_sensitive:
      muzero-general/games/roundabout_env_ttc_flat_risk_sensitive.py:30-33,45-55,60-69,83-88
  cross_merge_env (cfg "one"):
      muzero-general/games/cross_merge_env.py:31-34,66-69,81-91,95-96,112-118,148
"""
import copy


class SearchConfig:
    """Attribute bag with the reference's ``MuZeroConfig`` attribute names."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def replace(self, **kw):
        c = copy.deepcopy(self)
        c.__dict__.update(kw)
        return c

    def __repr__(self):
        return "SearchConfig(%s)" % ", ".join("%s=%r" % kv for kv in sorted(self.__dict__.items()))


_TTC_OBS = 5 * 3 * 10 + 5  # NUM_SPEEDS * NUM_LANES * HORIZON + NUM_SPEEDS = 155

_COMMON = dict(
    action_space=list(range(5)),
    players=list(range(1)),
    stacked_observations=1,
    root_dirichlet_alpha=0.25,
    root_exploration_fraction=0.25,
    pb_c_base=19652,
    pb_c_init=1.25,
    num_simulations=25,
    discount=0.975,
    mask_absorbing_states=False,
    fc_reward_layers=[32],
    fc_value_layers=[32],
    fc_policy_layers=[32],
    fc_reconstruction_layers=[32],
)

_PRESETS = {
    "highway_env_ttc_flat": dict(
        _COMMON,
        name="highway_env_ttc_flat",
        network="fullyconnected",
        observation_shape=(1, 1, _TTC_OBS),
        support_size=10,
        encoding_size=64,
        fc_representation_layers=[64, 64, 32],
        fc_dynamics_layers=[32],
    ),
    "roundabout_env_ttc_flat": dict(
        _COMMON,
        name="roundabout_env_ttc_flat",
        network="fullyconnected",
        observation_shape=(1, 1, _TTC_OBS),
        support_size=10,
        encoding_size=64,
        fc_representation_layers=[64, 64, 32],
        fc_dynamics_layers=[32],
    ),
    "roundabout_env_ttc_flat_stochastic_concat": dict(
        _COMMON,
        name="roundabout_env_ttc_flat_stochastic_concat",
        network="stochastic_concat",
        stochastic_dynamics=True,
        n_futures=2,
        fc_prior_layers=[32],
        fc_posterior_layers=[32],
        observation_shape=(1, 1, _TTC_OBS),
        support_size=10,
        encoding_size=64,
        fc_representation_layers=[64, 64, 32],
        fc_dynamics_layers=[32],
    ),
    "highway_env_ttc_flat_stochastic": dict(
        _COMMON,
        name="highway_env_ttc_flat_stochastic",
        network="stochastic",
        stochastic_dynamics=True,
        n_futures=2,
        fc_prior_layers=[32],
        fc_posterior_layers=[32],
        observation_shape=(1, 1, _TTC_OBS),
        support_size=10,
        encoding_size=64,
        fc_representation_layers=[64, 64, 32],
        fc_dynamics_layers=[32],
    ),
    "roundabout_env_ttc_flat_stochastic": dict(
        _COMMON,
        name="roundabout_env_ttc_flat_stochastic",
        network="stochastic",
        stochastic_dynamics=True,
        n_futures=4,
        fc_prior_layers=[32],
        fc_posterior_layers=[32],
        observation_shape=(1, 1, _TTC_OBS),
        support_size=10,
        encoding_size=64,
        fc_representation_layers=[64, 64, 32],
        fc_dynamics_layers=[32],
    ),
    "roundabout_env_ttc_flat_risk_sensitive": dict(
        _COMMON,
        name="roundabout_env_ttc_flat_risk_sensitive",
        network="stochastic_concat",
        stochastic_dynamics=True,
        risk_sensitive=True,
        n_futures=2,
        fc_prior_layers=[32],
        fc_posterior_layers=[32],
        observation_shape=(1, 1, _TTC_OBS),
        support_size=10,
        encoding_size=64,
        fc_representation_layers=[64, 64, 32],
        fc_dynamics_layers=[32],
    ),
    "cross_merge_env": dict(
        _COMMON,
        name="cross_merge_env",
        network="fullyconnected",
        observation_shape=(1, 1, 183),
        support_size=3,
        encoding_size=256,
        fc_representation_layers=[256, 128, 128],
        fc_dynamics_layers=[128],
        discount=0.995,
        pb_c_init=1.5,
    ),
}

PRESET_NAMES = tuple(_PRESETS)


def get_config(name, **overrides):
    """Return the ``SearchConfig`` of a reference game file, optionally overridden
    (the reference applies overrides the same way, muzero.py:69-74)."""
    if name not in _PRESETS:
        raise KeyError("unknown game config %r (have: %s)" % (name, ", ".join(_PRESETS)))
    d = copy.deepcopy(_PRESETS[name])
    d.update(overrides)
    return SearchConfig(**d)


def variant_of(config):
    """0 = MuZero MCTS (self_play.py), 1 = stochastic (self_play_stochastic.py),
    2 = risk-sensitive (self_play_risk_sensitive.py); same dispatch as muzero.py:232-250."""
    if getattr(config, "risk_sensitive", False):
        return 2
    if getattr(config, "stochastic_dynamics", False):
        return 1
    return 0


def observation_dim(config):
    """Flattened root input size, models.py:157-163."""
    c, h, w = config.observation_shape
    so = config.stacked_observations
    return c * h * w * (so + 1) + so * h * w
