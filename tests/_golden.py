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
    """Rows of table ``key`` whose first update is a stable function of the gradient.

    The first Adagrad / Adam step is lr * g / (|g| + eps): where the true gradient is ~eps (pos and
    neg contributions cancelling, a saturated sigmoid) its rounding noise decides the step, and no
    two implementations -- nor the reference at two thread counts -- agree.  Those rows are left out
    of the one-step comparison."""
    import numpy as np
    from oracle import cf_oracle as O
    ok = np.ones(n_rows, dtype=bool)
    if opt == "sgd":
        return ok
    _, grads = (O.linear_grads if net_type == "linear" else O.fm_grads)(init, batch0)
    rows, gsum = O.coalesce(*grads[key])
    tiny = np.abs(gsum).min(axis=1) < 1e-5
    if not (net_type == "linear" and key == "user_bias.weight"):  # exactly-zero gradient: no step at all
        ok[rows[tiny]] = False
    return ok
