"""``Pod.inverse_transform`` and the 8-line ``Pod._fit`` (pod/pod.py:198-228, 230-262).

TEST INFRASTRUCTURE ONLY.
"""
import numpy as np


def inverse_transform(mean_snapshot, U, samples):
    """pred = mean_snapshot + (U @ samples.T).T   (pod/pod.py:226-228)."""
    samples = np.asarray(samples)
    pred = mean_snapshot + np.dot(U, samples.T).T
    return np.atleast_2d(pred)


def fit(snapshots, tolerance, dim_max):
    """snapshots (n_features, n_samples) -> mean_snapshot, U, S, V after filtering.

    pod/pod.py:210-217 (centre, reduced SVD) and 246-260 (cumulative-S filter:
    the ratio is on S, not S**2)."""
    snapshots = np.array(snapshots, dtype=np.float64)
    mean_snapshot = np.average(snapshots, 1)
    snapshots = snapshots - mean_snapshot[:, None]
    U, S, Vt = np.linalg.svd(snapshots, full_matrices=False)
    V = Vt.T
    ratio = np.cumsum(S) / np.sum(S)
    dim = min(int(np.searchsorted(ratio, tolerance)) + 1, dim_max)
    return mean_snapshot, U[:, :dim].copy(), S[:dim].copy(), V[:, :dim].copy()
