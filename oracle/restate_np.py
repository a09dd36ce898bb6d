"""oracle/restate_np.py -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

numpy restatements of the reference algorithms behind the single-target estimators, the criterion / dense searches and the
filtered kNN.  They are pinned against vectors produced by the unmodified
reference (tests/golden/ref_extra.npz, written by tests/golden/make_golden_extra.py from oracle/_ref) in
tests/test_oracle_cpu.py, and state what the CUDA kernels must compute.  Every function cites the reference lines it follows.
"""
import numpy as np


def wsm_table(kern, r):
    """Wsm of the reference (KDCalcSmoothQuantities.cxx:12-15) for r = rij/hi"""
    size = len(kern)
    delta = 2.0 / (size - 1)
    i = int(r * 0.5 * (size - 1))
    if i < size - 1:
        return kern[i] + (kern[i + 1] - kern[i]) * (r - delta * i) / delta
    return kern[i]


def gather_density(kern, d2_row, w_row):
    """CalcDensityParticle / CalcSmoothLocalValue sum (KDCalcSmoothQuantities.cxx:768-844, 1704-1735): d2_row ascending,
    summed from the farthest neighbour inwards like the reference's heap pops"""
    hi = 0.5 * np.sqrt(d2_row[-1])
    norm = 1.0 / hi ** 3.0
    acc = 0.0
    for j in range(len(d2_row) - 1, -1, -1):
        acc += wsm_table(kern, np.sqrt(d2_row[j]) / hi) * norm * w_row[j]
    return acc


def gather_veldensity(kern, vq, vnb, kv):
    """CalcVelDensityParticle (KDCalcSmoothQuantities.cxx:887-911): kv smallest velocity distances among the spatial
    neighbours vnb, h = half the largest of them, sum in descending order"""
    dv = vq[None, :] - vnb
    u = np.sqrt((dv[:, 0] * dv[:, 0] + dv[:, 1] * dv[:, 1]) + dv[:, 2] * dv[:, 2])
    u = np.sort(u)[:kv][::-1]
    hi = 0.5 * u[0]
    norm = 1.0 / hi ** 3.0
    acc = 0.0
    for x in u:
        acc += wsm_table(kern, x / hi) * norm
    return acc


def reflect_images(x, period):
    """the 8 query positions of the reference's periodic searches (DistFunc.h:326-355): +p if x < p/2 else -p per axis"""
    if period is None:
        return [np.asarray(x, dtype=np.float64)]
    x = np.asarray(x, dtype=np.float64)
    s = np.where(x < period / 2.0, x + period, x - period)
    out = []
    for img in range(8):
        m = np.array([img & 1, (img >> 1) & 1, (img >> 2) & 1], dtype=bool)
        out.append(np.where(m, s, x))
    return out


def crit_rows(pos, vel, xq, vq, crit, params, period, exclude=None):
    """brute-force SearchCriterionTagged rows (sorted particle IDs) with the reference's arithmetic (FOFFunc.h:30-55)"""
    rows = []
    for q in range(len(xq)):
        hit = np.zeros(len(pos), dtype=bool)
        for xi in reflect_images(xq[q], period):
            d = xi[None, :] - pos
            if crit == 0:
                t = (d[:, 0] * d[:, 0] / params[6] + d[:, 1] * d[:, 1] / params[6]) + d[:, 2] * d[:, 2] / params[6]
            else:
                w = vq[q][None, :] - vel
                t = d[:, 0] * d[:, 0] / params[6]
                t = t + w[:, 0] * w[:, 0] / params[7]
                t = t + d[:, 1] * d[:, 1] / params[6]
                t = t + w[:, 1] * w[:, 1] / params[7]
                t = t + d[:, 2] * d[:, 2] / params[6]
                t = t + w[:, 2] * w[:, 2] / params[7]
            hit |= t < 1
        if exclude is not None:
            hit[exclude[q]] = False
        rows.append(np.nonzero(hit)[0])
    return rows


def ball_min_d2(pos, x, period):
    """smallest squared distance over the reference's images, and the plain (unreflected) one"""
    best = np.full(len(pos), np.inf)
    for xi in reflect_images(x, period):
        d = xi[None, :] - pos
        best = np.minimum(best, (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2])
    return best
