"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/rootio.npz from the LIVE reference:
GameHistory.get_stacked_observations / store_search_statistics and SelfPlay.select_action of
muzero-general/self_play.py, called unmodified. For select_action the global numpy RNG is seeded per case and
the uniform sample numpy.random.choice(actions, p=...) consumes (its first random_sample()) is stored beside
the action the reference chose.   Run:  python -m oracle.make_golden_rootio"""
import os
import sys
import types

import numpy

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import reference_live as RL  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "rootio.npz")


def main():
    ref = RL.load_reference()
    sp = ref.self_play
    rng = numpy.random.default_rng(77)
    out = {}
    # ---- select_action: 400 roots, ragged ordered legal lists, temperatures of the game files (1.0, 0.5, 0.25) and 0
    A, N = 5, 400
    visits = numpy.zeros((N, A), numpy.int32)
    legal = numpy.zeros((N, A), numpy.int32)
    nlegal = numpy.zeros(N, numpy.int32)
    temps = numpy.zeros(N)
    us = numpy.zeros(N)
    acts = numpy.zeros(N, numpy.int32)
    for i in range(N):
        n = int(rng.integers(1, A + 1))
        la = [int(a) for a in rng.permutation(A)[:n]]
        S = int(rng.choice([25, 50, 200]))
        v = rng.multinomial(S, rng.dirichlet([0.5] * n))
        root = types.SimpleNamespace(children={a: types.SimpleNamespace(visit_count=int(c)) for a, c in zip(la, v)})
        t = float(rng.choice([0.0, 0.25, 0.5, 1.0]))
        numpy.random.seed(1000 + i)
        us[i] = numpy.random.random_sample()
        numpy.random.seed(1000 + i)
        acts[i] = sp.SelfPlay.select_action(root, t)
        legal[i, :n] = la
        nlegal[i] = n
        for a, c in zip(la, v):
            visits[i, a] = c
        temps[i] = t
    out.update(sa_visits=visits, sa_legal=legal, sa_nlegal=nlegal, sa_temperature=temps, sa_uniform=us, sa_action=acts)
    # ---- store_search_statistics
    cv = numpy.zeros((N, A))
    for i in range(N):
        gh = sp.GameHistory()
        root = types.SimpleNamespace(children={int(legal[i, k]): types.SimpleNamespace(visit_count=int(visits[i, legal[i, k]]))
                                               for k in range(nlegal[i])}, value=lambda: 0.0)
        gh.store_search_statistics(root, list(range(A)))
        cv[i] = gh.child_visits[0]
    out["ss_child_visits"] = cv
    # ---- get_stacked_observations: histories of (C, H, W) frames, several stackings and indices
    cases = []
    for ci, (C, H, W, so, L) in enumerate([(1, 1, 155, 1, 6), (1, 1, 183, 1, 4), (3, 2, 5, 3, 7), (1, 1, 31, 0, 3)]):
        gh = sp.GameHistory()
        frames = rng.random((L, C, H, W)).astype(numpy.float32)
        actions = rng.integers(0, 5, size=L)
        for t in range(L):
            gh.observation_history.append(frames[t])
            gh.action_history.append(int(actions[t]))
        for index in list(range(L)) + [-1]:
            st = gh.get_stacked_observations(index, so)
            out["so_case%d_index%d" % (ci, index)] = numpy.asarray(st, numpy.float32)
        out["so_case%d_frames" % ci] = frames
        out["so_case%d_actions" % ci] = actions.astype(numpy.int32)
        cases.append((C, H, W, so, L))
    out["so_cases"] = numpy.array(cases, numpy.int32)
    numpy.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
