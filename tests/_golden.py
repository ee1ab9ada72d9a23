"""Helpers to read tests/golden/*.npz (recorded from the live reference by oracle/make_golden.py)."""
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def section(g, prefix):
    """All entries under ``prefix/`` as a fresh {name: array} dict."""
    n = len(prefix) + 1
    return {k[n:]: v.copy() for k, v in g.items() if k.startswith(prefix + "/")}


def batch_at(g, s):
    return {k[6:]: v[s] for k, v in g.items() if k.startswith("batch/")}


def parse_train_name(name):
    # train_<net>_F<k>_<opt>[_refinit]
    parts = name.split("_")
    net, F = parts[1], int(parts[2][1:])
    opt = "_".join(p for p in parts[3:] if p != "refinit")
    return net, F, opt


def well_conditioned_rows(net_type, opt, init, batch0, key, n_rows):
    """ENTRIES (not whole rows: the name is historic) of table ``key`` whose first update is a stable function of the
    gradient, as a boolean mask of the table's shape.

    The first Adagrad / Adam step is lr * g / (|g| + eps): where the summed gradient is tiny but NOT zero (pos and neg
    contributions cancelling to rounding noise, a saturated sigmoid) the noise decides the step, and no two
    implementations -- nor the reference at two thread counts -- agree.  Exactly-zero gradients (an inactive hinge,
    -g + g on a bias) give an exactly-zero step and stay in.  Only 0 < |g| < 1e-5 is left out of the one-step comparison,
    and the share of touched entries this drops is bounded here: < 2 % of an embedding table's touched entries, and at
    most a quarter of the handful of touched entries of an FM first-order table (delta+ + delta- nearly cancels at
    initialisation, where both sigmoids sit at 0.5)."""
    import numpy as np
    from oracle import cf_oracle as O
    shape = init[key].shape
    ok = np.ones(shape, dtype=bool)
    if opt == "sgd":
        return ok
    _, grads = (O.linear_grads if net_type == "linear" else O.fm_grads)(init, batch0)
    rows, gsum = O.coalesce(*grads[key])
    a = np.abs(gsum.reshape(len(rows), -1))
    tiny = (a > 0) & (a < 1e-5)
    share = float(tiny.mean()) if tiny.size else 0.0
    limit = 0.02 if shape[1] > 1 else 0.25
    assert share <= limit, f"{key}: {share:.1%} of the touched entries are ill-conditioned (limit {limit:.0%})"
    ok[rows] = ~tiny
    return ok
