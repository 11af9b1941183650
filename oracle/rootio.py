"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the data formats on either side of the search:
GameHistory.get_stacked_observations (self_play.py:587-624), SelfPlay.select_action (self_play.py:277-299,
with numpy.random.choice(actions, p=...) replaced by its own inverse-CDF algorithm on an external uniform
sample) and GameHistory.store_search_statistics (self_play.py:570-585). Pinned against the live reference
functions by tests/test_rootio.py through tests/golden/rootio.npz (oracle/make_golden_rootio.py).
Nothing under tree_search_planning_b200/ imports this file."""
import numpy


def stacked_observations(observation_history, action_history, index, num_stacked):
    """self_play.py:587-624; observation frames are (C, H, W) arrays."""
    index = index % len(observation_history)
    stacked = numpy.array(observation_history[index], copy=True)
    for past in reversed(range(index - num_stacked, index)):
        if 0 <= past:
            prev = numpy.concatenate((observation_history[past],
                                      [numpy.ones_like(stacked[0]) * action_history[past + 1]]))
        else:
            prev = numpy.concatenate((numpy.zeros_like(observation_history[index]), [numpy.zeros_like(stacked[0])]))
        stacked = numpy.concatenate((stacked, prev))
    return stacked


def select_action(visit_counts, actions, temperature, u):
    """self_play.py:277-299 for one root; `u` in [0, 1) replaces the RNG draw of numpy.random.choice:
    choice(actions, p=p): cdf = p.cumsum(); cdf /= cdf[-1]; idx = cdf.searchsorted(u, side='right');
    choice(actions): actions[floor(u * n)]."""
    visit_counts = numpy.array(visit_counts, dtype="int32")
    if temperature == 0:
        return actions[int(numpy.argmax(visit_counts))]
    if temperature == float("inf"):
        return actions[min(int(numpy.floor(u * len(actions))), len(actions) - 1)]
    dist = visit_counts ** (1 / temperature)
    dist = dist / sum(dist)
    cdf = dist.cumsum()
    cdf /= cdf[-1]
    return actions[min(int(cdf.searchsorted(u, side="right")), len(actions) - 1)]


def child_visits(visit_counts_by_action, action_space):
    """self_play.py:570-581: visit_counts_by_action maps legal action -> visit count."""
    sum_visits = sum(visit_counts_by_action.values())
    return [visit_counts_by_action[a] / max(sum_visits, 1) if a in visit_counts_by_action else 0 for a in action_space]
