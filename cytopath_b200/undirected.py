"""f1: `undirected_simulations`, the second caller of the sampler (reference cytopath/utils.py:42-101): every cell (or a random
subset) is a root, every cell an end point; same signature and side effects, the sampling itself is `cytopath_b200.sampling`."""
from __future__ import annotations

import numpy as np

from .sampler import sampling

__all__ = ["undirected_simulations"]


def undirected_simulations(data, matrix_key="T_forward", max_steps=10, sim_number=None, normalize=False, unique=True,
                           num_cores=1, copy=False, *, store_samples=True, **b200_kwargs):
    """See cytopath.utils.undirected_simulations (utils.py:42-71) for the arguments.  `b200_kwargs`
    (seed, cdf_mode, device, devices, dist) are passed to `cytopath_b200.sampling`.

    store_samples=False is the fused mode: `obs['log_terminal_state_freq']` only needs where the
    selected sequences END (utils.py:92-95), and the last cell of a selected sequence is the end point it was
    selected for, so no sequence is materialised at all -- `uns['undirected_samples']` is then not written.  Either
    way nothing of size [walkers, max_steps] is ever allocated: the devices keep 16-byte records of the kept
    sequences and replay the selected ones."""
    from . import _lib
    adata = data.copy()                                                     # utils.py:73

    adata.obs["undirected_sim_included"] = "1"                              # utils.py:75-76
    adata.obs["undirected_sim_included"] = adata.obs["undirected_sim_included"].astype("category")

    n = adata.shape[0]
    if sim_number is None:                                                  # utils.py:78-79
        sim_number = n
    backend = b200_kwargs.pop("_backend", None) or _lib
    group = backend.open_group(device=b200_kwargs.pop("device", None), devices=b200_kwargs.pop("devices", None),
                               dist=b200_kwargs.pop("dist", "auto"))
    try:
        if sim_number < n:                                                  # utils.py:81-84
            root_cells = np.empty(sim_number, dtype=np.int64)
            if group.rank == 0:                                             # one draw for all ranks (their RNGs are not seeded alike)
                root_cells[:] = np.random.choice(np.arange(n), size=sim_number, replace=False)
            if group.world > 1:
                group.broadcast(root_cells)
        else:
            root_cells = np.arange(n)

        sampling(adata, matrix_key=matrix_key, normalize=normalize, unique=unique, root_cells=root_cells,      # utils.py:86-90
                 end_points=np.arange(n), cluster_key="undirected_sim_included", auto_adjust=False, min_clusters=1,
                 sim_number=sim_number, traj_number=sim_number, max_steps=max_steps, num_cores=num_cores,
                 _backend=backend, _group=group, _rows=bool(store_samples), **b200_kwargs)
    finally:
        group.close()

    final_cells = adata.uns["samples"]["cell_sequences"][:, -1] if store_samples else adata.uns["run_info"]["b200"]["final_cells"]
    cells, counts = np.unique(final_cells, return_counts=True)             # utils.py:92

    adata.obs["log_terminal_state_freq"] = float(np.nan)                    # utils.py:94-95
    adata.obs.iloc[cells, adata.obs.columns.get_loc("log_terminal_state_freq")] = np.log1p(counts / sim_number)

    data = data.copy() if copy else data                                    # utils.py:97-99
    data.obs["log_terminal_state_freq"] = adata.obs["log_terminal_state_freq"]
    if store_samples:
        data.uns["undirected_samples"] = adata.uns["samples"]

    if copy:
        return data
