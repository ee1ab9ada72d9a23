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
