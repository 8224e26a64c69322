"""f2: `coordinate_assigner` / `distance_matrix` of cytopath/trajectory_estimation/sample_clustering.py:38-71, the latter as
one CUDA kernel (csrc/hausdorff.cu) instead of S(S+1)/2 Python-level calls of a numba loop."""
from __future__ import annotations

import numpy as np

from . import _lib

__all__ = ["coordinate_assigner", "distance_matrix"]


def coordinate_assigner(adata, all_seq_cluster, basis="umap"):
    """sample_clustering.py:38-55 — coordinates [n_sequences, n_steps, n_dims] of the cells of each sequence."""
    if basis is None:
        map_state = adata.layers["Ms"]
    elif type(basis) == list:
        map_state = adata[:, basis].layers["Ms"]
    else:
        map_state = adata.obsm["X_" + basis]
    seq = np.asarray(all_seq_cluster).astype(np.int64)
    return np.asarray(map_state)[seq]


def distance_matrix(sequence_coordinates, distance="euclidean", num_cores=1, *, device=0):
    """sample_clustering.py:57-71 — symmetric matrix of Hausdorff distances between all chains."""
    print("Calculating hausdorff distances")
    return _lib.hausdorff_matrix(np.asarray(sequence_coordinates), distance, device)
