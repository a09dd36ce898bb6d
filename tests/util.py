import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "ref_small.npz")


def load_golden():
    g = np.load(GOLDEN)
    return {k: g[k] for k in g.files}


def canon(g):
    from oracle.pyoracle import canonical_groups
    return canonical_groups(g)


def rows_equal_as_sets(a, b):
    return np.array_equal(np.sort(a, axis=1), np.sort(b, axis=1))


def csr_rows_sorted(off, idx):
    return [np.sort(idx[off[i]:off[i + 1]]) for i in range(len(off) - 1)]


GOLDEN_EXTRA = os.path.join(ROOT, "tests", "golden", "ref_extra.npz")


def load_golden_extra():
    g = np.load(GOLDEN_EXTRA)
    return {k: g[k] for k in g.files}


GOLDEN_PHASE = os.path.join(ROOT, "tests", "golden", "ref_phase.npz")


def load_golden_phase():
    g = np.load(GOLDEN_PHASE)
    return {k: g[k] for k in g.files}


# numpy restatements of the reference algorithms live with the oracle (oracle/restate_np.py); re-exported for the tests
from oracle.restate_np import ball_min_d2, crit_rows, gather_density, gather_veldensity, reflect_images, wsm_table  # noqa: E402,F401
