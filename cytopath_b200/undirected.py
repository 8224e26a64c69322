"""`undirected_simulations` — the second caller of the sampler (reference cytopath/utils.py:42-101).

Every (or a random subset of) cell is a root and every cell is an end point, so no walker is
filtered; the outputs are the sampler's `uns['samples']` (stored as `uns['undirected_samples']`) and
`obs['log_terminal_state_freq']` = log1p(visits as final state / sim_number).  Same signature and
side effects as the reference; the sampling itself is `cytopath_b200.sampling` (CUDA).
"""
from __future__ import annotations

import numpy as np

from .sampler import sampling

__all__ = ["undirected_simulations"]


def undirected_simulations(data, matrix_key="T_forward", max_steps=10, sim_number=None, normalize=False, unique=True,
                           num_cores=1, copy=False, **b200_kwargs):
    """See cytopath.utils.undirected_simulations (utils.py:42-71) for the arguments.  `b200_kwargs`
    (seed, cdf_mode, device, dist, ...) are passed to `cytopath_b200.sampling`."""
    adata = data.copy()                                                     # utils.py:73

    adata.obs["undirected_sim_included"] = "1"                              # utils.py:75-76
    adata.obs["undirected_sim_included"] = adata.obs["undirected_sim_included"].astype("category")

    n = adata.shape[0]
    if sim_number is None:                                                  # utils.py:78-79
        sim_number = n
    if sim_number < n:                                                      # utils.py:81-84
        root_cells = np.random.choice(np.arange(n), size=sim_number, replace=False)
    else:
        root_cells = np.arange(n)

    sampling(adata, matrix_key=matrix_key, normalize=normalize, unique=unique, root_cells=root_cells,      # utils.py:86-90
             end_points=np.arange(n), cluster_key="undirected_sim_included", auto_adjust=False, min_clusters=1,
             sim_number=sim_number, traj_number=sim_number, max_steps=max_steps, num_cores=num_cores, **b200_kwargs)

    cells, counts = np.unique(adata.uns["samples"]["cell_sequences"][:, -1], return_counts=True)          # utils.py:92

    adata.obs["log_terminal_state_freq"] = float(np.nan)                    # utils.py:94-95
    adata.obs.iloc[cells, adata.obs.columns.get_loc("log_terminal_state_freq")] = np.log1p(counts / sim_number)

    data = data.copy() if copy else data                                    # utils.py:97-99
    data.obs["log_terminal_state_freq"] = adata.obs["log_terminal_state_freq"]
    data.uns["undirected_samples"] = adata.uns["samples"]

    if copy:
        return data
