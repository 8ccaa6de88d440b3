"""CPU restatement of the reference's Wasserstein metrics -- TEST INFRASTRUCTURE ONLY (see fdiff_oracle.py header).

Follows fdiff/utils/wasserstein.py:12-199 and fdiff/sampling/metrics.py:101-216.  The 1-D transport cost is computed
by the third-party dependency POT (``ot.emd2_1d``; pyproject.toml:50 lists ``"POT"`` unpinned; POT is NOT installed in
this image and not vendored under /root/reference), so **parity is unpinned by reference runs**: ``emd2_1d`` below
restates POT's published algorithm (ot/lp/emd_wrap.pyx ``emd_1d_sorted``: sort both samples, walk the two sorted
arrays, move min(remaining masses) at squared-Euclidean cost) and is pinned instead to the DEFINITION of the cost by
solving the transport linear programme on small cases (tests/test_oracle_golden.py) and to closed forms
(translation: W2 = |shift|).  Everything above ``emd2_1d`` (directions, projection, normalisation, sqrt, dict keys) is
the reference's own numpy code restated line by line.
"""

from __future__ import annotations

import numpy as np


def emd2_1d(xa: np.ndarray, xb: np.ndarray) -> float:
    """Exact OT cost between two 1-D samples, uniform weights, squared-Euclidean ground cost
    (what ``ot.emd2_1d(x_a, x_b)`` returns with its defaults; wasserstein.py:117,141)."""
    a = np.sort(np.asarray(xa, np.float64))
    b = np.sort(np.asarray(xb, np.float64))
    n, m = len(a), len(b)
    # masses in units of 1 / (n m): a_i owns [i m, (i + 1) m), b_j owns [j n, (j + 1) n)
    cuts = np.union1d(np.arange(n + 1, dtype=np.int64) * m, np.arange(m + 1, dtype=np.int64) * n)
    lo, hi = cuts[:-1], cuts[1:]
    i, j = lo // m, lo // n
    return float(np.sum((a[i] - b[j]) ** 2 * (hi - lo)) / (n * m))


def emd2_1d_walk(xa, xb) -> float:
    """The same by POT's two-pointer walk (kept literal; O(n + m) Python loop -- small cases only)."""
    a, b = np.sort(np.asarray(xa, np.float64)), np.sort(np.asarray(xb, np.float64))
    n, m = len(a), len(b)
    wa, wb = 1.0 / n, 1.0 / m
    i = j = 0
    ra, rb = wa, wb
    cost = 0.0
    while i < n and j < m:
        mass = min(ra, rb)
        cost += mass * (a[i] - b[j]) ** 2
        ra -= mass
        rb -= mass
        if ra <= 1e-15:
            i += 1
            ra = wa
        if rb <= 1e-15:
            j += 1
            rb = wb
    return cost


class WassersteinDistances:
    """wasserstein.py:12-199."""

    def __init__(self, original_data, other_data, normalisation="none", seed=None):
        self.original_data = np.asarray(original_data)
        self.other_data = np.asarray(other_data)
        self.normalisation = normalisation
        self.rng = np.random.default_rng(seed)           # :40

    def random_direction(self, dim):                      # :42-60
        vector = self.rng.normal(size=dim)
        return vector / np.linalg.norm(vector)

    def get_random_directions(self, n_directions):        # :62-78
        dimension = self.original_data.shape[1]
        return [self.random_direction(dimension) for _ in range(n_directions)]

    def _normalise(self, orig, other):                    # :151-160
        if self.normalisation == "none":
            return orig, other
        if self.normalisation == "standardise":
            sd = np.std(orig)
            return orig / sd, other / sd
        raise ValueError(f"Unrecognised normalisation type: {self.normalisation}")

    def feature_distance(self, feature):                  # :96-119
        o, x = self._normalise(self.original_data[:, feature], self.other_data[:, feature])
        return float(np.sqrt(emd2_1d(o, x)))

    def directional_distance(self, direction):            # :121-143
        o, x = self._normalise(self.original_data @ direction, self.other_data @ direction)
        return float(np.sqrt(emd2_1d(o, x)))

    def sliced_distances(self, num_directions):           # :162-180
        return np.array([self.directional_distance(d) for d in self.get_random_directions(num_directions)])

    def marginal_distances(self):                         # :182-199
        return np.array([self.feature_distance(f) for f in range(self.original_data.shape[1])])


def sliced_wasserstein(original, other, random_seed, num_directions):
    """SlicedWasserstein.__call__ (metrics.py:116-128) on flattened (n, L*C) arrays."""
    d = WassersteinDistances(original.reshape(len(original), -1), other.reshape(len(other), -1),
                             seed=random_seed).sliced_distances(num_directions)
    return {"sliced_wasserstein_mean": float(np.mean(d)), "sliced_wasserstein_max": float(np.max(d))}, d


def marginal_wasserstein(original, other, random_seed):
    """MarginalWasserstein.__call__ (metrics.py:174-186)."""
    d = WassersteinDistances(original.reshape(len(original), -1), other.reshape(len(other), -1),
                             seed=random_seed).marginal_distances()
    return {"marginal_wasserstein_mean": float(np.mean(d)), "marginal_wasserstein_max": float(np.max(d))}, d
