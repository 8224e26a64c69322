"""Drop-in for the alignment scoring that follows trajectory estimation
(reference cytopath/trajectory_estimation/cyto_path.py:155-243, `directionality_score`).

The reference scores every cell of every neighbourhood along every mean trajectory in a Python triple loop (joblib
over steps, :239); here all (trajectory, step, cell) instances go to one CUDA kernel (csrc/directionality.cu).
"""
from __future__ import annotations

import numpy as np

from . import _lib

__all__ = ["directionality_score"]


def _masked_velocity_graph(adata):
    """cyto_path.py:181-187: (velocity_graph + velocity_graph_neg) restricted to the pattern of T_forward."""
    transmat = adata.uns["T_forward"].copy()
    transmat[transmat.nonzero()] = 1
    if "velocity_graph" and "velocity_graph_neg" in adata.uns.keys():          # the reference's test, as written (:184)
        velocity_graph = adata.uns["velocity_graph"] + adata.uns["velocity_graph_neg"]
    return velocity_graph.multiply(transmat).tocsr()


def directionality_score(adata, neighborhood_sequence, end_point, map_state, num_cores=1, *, device=0):
    """Scores the directionality of the cells along the path (same arguments and return value as the reference:
    a list, per trajectory, of the list, per step, of the score arrays of the step's cells).  `num_cores` is
    accepted and ignored."""
    final_cluster = adata.uns["trajectories"]["trajectories_coordinates"][end_point]
    vg = _masked_velocity_graph(adata)
    vg.sort_indices()
    # the reference takes the neighbours from row.nonzero() and the weights from row.data (:198,:203): explicit zeros
    # would desynchronise the two, scipy's sparse add / multiply do not produce them
    vg.eliminate_zeros()
    w = np.expm1(vg.data)                                   # in the graph's dtype, like :203
    map_state = np.asarray(map_state)
    dirs, cell, fwd, bwd, shape = [], [], [], [], []
    for i in range(len(final_cluster)):
        coords = np.asarray(final_cluster["trajectory_{}_coordinates".format(i)], dtype=np.float64)
        steps = neighborhood_sequence[i]
        base = len(dirs)
        dirs.extend(coords[1:] - coords[:-1])               # direction d: coords[d + 1] - coords[d]
        last = len(steps) - 1
        sizes = []
        for j in range(len(steps)):
            cells = np.asarray(steps[j], dtype=np.int64).reshape(-1)
            if j == 0 and last == 0:
                raise IndexError("index 1 is out of bounds for axis 0 with size 1")       # :192 on a one-step trajectory
            cell.append(cells)
            fwd.append(np.full(len(cells), base + j if j != last else -1, dtype=np.int32))
            bwd.append(np.full(len(cells), base + j - 1 if j != 0 else -1, dtype=np.int32))
            sizes.append(len(cells))
        shape.append(sizes)
    n_inst = sum(sum(s) for s in shape)
    if n_inst:
        flat = _lib.directionality_scores(map_state, vg.indptr, vg.indices, w, np.asarray(dirs, dtype=np.float64).reshape(-1, map_state.shape[1]),
                                          np.concatenate(cell), np.concatenate(fwd), np.concatenate(bwd), device)
    else:
        flat = np.empty(0)
    all_scores, o = [], 0
    for sizes in shape:
        traj = []
        for n in sizes:
            traj.append(flat[o:o + n].copy())
            o += n
        all_scores.append(traj)
    return all_scores
