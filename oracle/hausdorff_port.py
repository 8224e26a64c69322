"""CPU restatement of cytopath's trajectory distance matrix.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/cytopath/trajectory_estimation/sample_clustering.py:38-71
(`coordinate_assigner`, `distance_matrix`).  The inner function `hausdorff_distance` comes from the
third-party PyPI package `hausdorff` (setup.py:19, unpinned; latest 0.2.6), which is NOT installed in
this image and cannot be fetched: PARITY UNPINNED for this row.  Its published algorithm (numba loops
over both point sets with an early break) computes
    H(A, B) = max( max_a min_b dist(a, b), max_b min_a dist(a, b) )
with dist in {euclidean, manhattan, chebyshev, cosine}; the early break does not change the value.
"""
import numpy as np


def _pairwise(A, B, distance):
    if distance == "euclidean":
        return np.sqrt(((A[:, None, :] - B[None, :, :]) ** 2).sum(-1))
    if distance == "manhattan":
        return np.abs(A[:, None, :] - B[None, :, :]).sum(-1)
    if distance == "chebyshev":
        return np.abs(A[:, None, :] - B[None, :, :]).max(-1)
    if distance == "cosine":
        return 1.0 - (A @ B.T) / np.sqrt((A * A).sum(1)[:, None] * (B * B).sum(1)[None, :])
    raise ValueError(distance)


def hausdorff_distance(XA, XB, distance="euclidean"):
    D = _pairwise(np.asarray(XA, dtype=np.float64), np.asarray(XB, dtype=np.float64), distance)
    return max(D.min(axis=1).max(), D.min(axis=0).max())


def coordinate_assigner(map_state, all_seq_cluster):
    """:38-55 — the double loop of vstack/stack is a fancy-index gather."""
    return np.stack([np.vstack([map_state[int(c), :] for c in row]) for row in all_seq_cluster], axis=0)


def distance_matrix(sequence_coordinates, distance="euclidean"):
    """:57-71 — same loop nest: upper triangle, mirrored."""
    S = len(sequence_coordinates)
    out = np.zeros((S, S))
    for i in range(S):
        for j in range(S - i):
            h = hausdorff_distance(sequence_coordinates[j + i], sequence_coordinates[i], distance=distance)
            out[i, j + i] = h
            out[j + i, i] = h
    return out
