"""Loading of the committed golden fixtures (written by oracle/make_golden.py from the LIVE
reference) and the comparisons shared by the CPU (oracle) and GPU (engine) parity tests."""
import ast
import glob
import os

import numpy

from tree_search_planning_b200.configs import get_config

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    # (rootio.npz holds the root I/O format goldens of tests/test_rootio.py, trees.npz the flattened Node graphs of
    # tests/test_tree_export.py: not search fixtures)
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                  if os.path.basename(p) not in ("rootio.npz", "trees.npz"))


class Golden:
    def __init__(self, name):
        z = numpy.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        self.meta = ast.literal_eval(str(z["meta"]))
        self.out = {k[4:]: z[k] for k in z.files if k.startswith("out:")}
        self.net = {k[4:]: z[k] for k in z.files if k.startswith("net:")}
        self.state_dict = {k[2:]: z[k] for k in z.files if k.startswith("w:")}
        for k in ("obs", "legal", "n_legal", "noise", "tie", "chance"):
            setattr(self, k, z[k])
        self.to_play = z["to_play"] if "to_play" in z.files else None
        self.S = int(self.meta["S"])   # simulations per search
        self.N = int(self.meta["N"])
        # > 1: the search was continued P - 1 times with override_root_with=root; noise / tie / chance / records
        # then carry a leading [P] dimension and the statistics are those of the final tree
        self.P = int(self.meta.get("phases", 1))
        cfg = get_config(self.meta["game"], num_simulations=self.S)
        if self.meta.get("mask_absorbing_states"):
            cfg = cfg.replace(mask_absorbing_states=True)
        if len(self.meta.get("players", [0])) != 1:
            cfg = cfg.replace(players=list(self.meta["players"]))
        self.config = cfg

    def run_kwargs(self, phase=None):
        """Keyword arguments of Oracle.run / Engine.run; `phase` selects one search of a continued fixture."""
        noise, tie, chance = self.noise, self.tie, self.chance
        if self.P > 1 and phase is not None:
            noise, tie, chance = noise[phase], tie[phase], chance[phase]
        return dict(legal_actions=self.legal, num_legal=self.n_legal, noise=noise, tie=tie,
                    chance=chance, add_exploration_noise=bool(self.meta["add_noise"]),
                    uniform_policy=bool(self.meta["uniform_policy"]), to_play=self.to_play)

    def oracle_kwargs(self):
        kw = self.run_kwargs()
        if self.P > 1:
            kw["n_phases"] = self.P
        return kw


TREE_KEYS_EXACT = ("visit_counts", "root_value", "child_value", "child_prior", "child_reward", "max_tree_depth")


def assert_tree_outputs_bit_exact(got, want, keys=TREE_KEYS_EXACT, what=""):
    """Bit-exact comparison of everything a consumer reads from `root`."""
    for k in keys:
        a, b = numpy.asarray(got[k]), numpy.asarray(want[k])
        assert a.shape == b.shape, (what, k, a.shape, b.shape)
        if a.dtype.kind == "f":
            same = (a.view(numpy.int64) == b.astype(a.dtype).view(numpy.int64)) if a.dtype == numpy.float64 else (a == b)
            # -0.0 vs 0.0 are the same number to every consumer
            same = same | ((a == 0) & (b == 0))
        else:
            same = a == b
        assert same.all(), "%s: %s differs at %s: got %s want %s" % (
            what, k, numpy.argwhere(~same)[:4].tolist(), a[~same][:4], b[~same][:4])


def assert_close_quantised(got, want, tol=1e-5, min_fraction=0.98, hard_tol=2e-3, what=""):
    """Tolerance for scalars that went through the reference's float32 inverse value transform
    (models.py:1192-1200). That formula subtracts 1 from sqrt(1 + 0.004*(|x|+1.001)) in float32,
    which quantises its output in steps of ~1.2e-4*(1+|x|); a 1-ulp difference upstream (or torch's
    not-correctly-rounded CPU sqrt) moves a result by one whole step. So: within `tol` (relative to
    max(1,|x|)) on at least `min_fraction` of the entries and within `hard_tol` on all of them."""
    a, b = numpy.asarray(got, numpy.float64), numpy.asarray(want, numpy.float64)
    err = numpy.abs(a - b) / numpy.maximum(1.0, numpy.abs(b))
    assert err.max() <= hard_tol, "%s: max error %g > %g" % (what, err.max(), hard_tol)
    frac = float((err <= tol).mean())
    # (a single quantisation step on a fixture with a handful of roots must not fail the fraction test)
    assert frac >= min_fraction or int((err > tol).sum()) <= 1, "%s: only %.4f of entries within %g" % (what, frac, tol)


def quantised_columns(records, n_futures):
    """Boolean mask [..., REC] of the record entries that went through the reference's float32 inverse value transform
    (support_to_scalar, models.py:1180-1201) and may therefore differ by one quantisation step: value and reward of a
    state record (kind 1: columns 1, 2), the per-future rewards of a transition record (kind 2: columns 1 + nf .. 1 + 2 nf).
    Everything else -- priors, transition probabilities, the terminal logit, the kind -- is plain float32 arithmetic."""
    r = numpy.asarray(records)
    q = numpy.zeros(r.shape, bool)
    state = r[..., 0] == 1.0
    trans = r[..., 0] == 2.0
    q[..., 1] |= state
    q[..., 2] |= state
    for c in range(1 + n_futures, 1 + 2 * n_futures):
        q[..., c] |= trans
    return q


def undisturbed_trees(got, want, n_futures, live=None, tol=1e-5, hard_tol=2e-3):
    """[trees] bool: the two searches evaluated the same leaves in the same order. A flipped near-tie can swap the order of
    two evaluations without changing a single visit count (sibling leaves with nearly equal outputs: two records then
    differ by anything from 1e-5 to 1e-1 and the trees re-converge), so no threshold separates "same evaluation,
    inaccurate network" from "other evaluation" record by record. What holds: entries that are not quantised sit at ~1e-7
    on an undisturbed tree, so a tree counts as undisturbed iff ALL its prior / probability entries are within `tol` and
    its value / reward entries within the quantisation step; callers bound the share of the others (a systematic network
    error fails that bound), and tests/test_tree_export.py checks every expansion of the search kernels path-independently."""
    a, b = numpy.asarray(got, numpy.float64), numpy.asarray(want, numpy.float64)
    err = numpy.abs(a - b) / numpy.maximum(1.0, numpy.abs(b))
    if live is not None:
        err = err * numpy.broadcast_to(numpy.asarray(live, bool)[..., None], a.shape)
    q = quantised_columns(want, n_futures)
    plain = numpy.where(q, 0.0, err)
    axes = tuple(range(1, a.ndim))
    return (plain.max(axis=axes) <= tol) & (err.max(axis=axes) <= hard_tol)


def assert_records_close(got, want, n_futures, live=None, tol=1e-5, min_fraction=0.97, hard_tol=2e-3, what="records"):
    """Column-wise tolerance for evaluation records: priors / probabilities / logits within `tol` on EVERY live entry;
    only the entries of `quantised_columns` get the quantisation allowance of assert_close_quantised."""
    a, b = numpy.asarray(got, numpy.float64), numpy.asarray(want, numpy.float64)
    err = numpy.abs(a - b) / numpy.maximum(1.0, numpy.abs(b))
    q = quantised_columns(want, n_futures)
    if live is None:
        live = numpy.ones(a.shape[:-1], bool)
    live = numpy.broadcast_to(numpy.asarray(live, bool)[..., None], a.shape)
    plain = live & ~q
    assert err[plain].size == 0 or err[plain].max() <= tol, "%s: prior / probability entry off by %g (> %g)" % (
        what, err[plain].max(), tol)
    e = err[live & q]
    if e.size:
        assert e.max() <= hard_tol, "%s: max error of a value / reward entry %g > %g" % (what, e.max(), hard_tol)
        frac = float((e <= tol).mean())
        assert frac >= min_fraction or int((e > tol).sum()) <= 1, "%s: only %.4f of value / reward entries within %g" % (what, frac, tol)
