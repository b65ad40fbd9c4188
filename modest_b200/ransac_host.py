"""Host half of the RANSAC stage in *parity mode*: drawing the minimal sets.

sklearn's RANSACRegressor (called by the reference's estimate_plane,
generate_cluster_mask/utils/pointcloud_utils.py:52) takes its 3-point subsets from numpy's
global RandomState through sklearn.utils.random.sample_without_replacement
(sklearn/linear_model/_ransac.py:472-474).  To reproduce the reference under a given
`np.random.seed`, the same stream has to be consumed in the same way, so the draws happen
here on the host with numpy -- random-number bookkeeping, not numerics.  The hypotheses
themselves are scored on the GPU (csrc/ransac.cu).

Protocol per fit:  `peek_triples(n, H)` returns the next H subsets *without* consuming the
stream; after the device reports how many trials sklearn's loop would have run,
`consume_trials(n, n_trials)` advances the stream by exactly that many draws.
"""
from __future__ import annotations

import numpy as np


def _draw_subset(n_population: int, n_samples: int, rng) -> np.ndarray:
    """sample_without_replacement(n_population, n_samples, method='auto') restated from
    sklearn/utils/_random.pyx:222-265 (tracking selection / permutation / reservoir)."""
    ratio = n_samples / n_population if n_population else 1.0
    if 0.01 < ratio < 0.99:
        return rng.permutation(n_population)[:n_samples]
    if ratio < 0.2:
        chosen, out = set(), np.empty(n_samples, dtype=np.int64)
        for i in range(n_samples):
            j = rng.randint(n_population)
            while j in chosen:
                j = rng.randint(n_population)
            chosen.add(j)
            out[i] = j
        return out
    out = np.arange(n_samples, dtype=np.int64)
    for i in range(n_samples, n_population):
        j = rng.randint(0, i + 1)
        if j < n_samples:
            out[j] = i
    return out


def global_rng():
    return np.random.mtrand._rand


def peek_triples(n_candidates: int, n_trials: int, rng=None) -> np.ndarray:
    """The next `n_trials` minimal sets the stream would yield, stream state restored."""
    rng = global_rng() if rng is None else rng
    out = np.zeros((n_trials, 3), dtype=np.int32)
    if n_candidates < 3:
        return out
    state = rng.get_state()
    try:
        for t in range(n_trials):
            out[t] = _draw_subset(n_candidates, 3, rng)
    finally:
        rng.set_state(state)
    return out


def consume_trials(n_candidates: int, n_trials: int, rng=None) -> None:
    rng = global_rng() if rng is None else rng
    if n_candidates < 3:
        return
    for _ in range(int(n_trials)):
        _draw_subset(n_candidates, 3, rng)
