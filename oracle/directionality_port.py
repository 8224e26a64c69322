"""TEST INFRASTRUCTURE (see oracle/__init__.py): numpy restatement of
cytopath/trajectory_estimation/cyto_path.py:155-243 (`directionality_score`) and of the numba cosine it calls
(cytopath/utils.py:14-39).  Never imported by the product."""
import numpy as np


def cosine_similarity(u, v):
    """utils.py:14-39: cos(u, v_i) for every row of v; 1 where either vector is all zero."""
    u = np.asarray(u, dtype=np.float64)
    v = np.asarray(v, dtype=np.float64)
    uv = v @ u
    uu = float(u @ u)
    vv = np.einsum("ij,ij->i", v, v)
    out = np.ones(v.shape[0])
    ok = (uu != 0) & (vv != 0)
    out[ok] = uv[ok] / np.sqrt(uu * vv[ok])
    return out


def masked_velocity_graph(adata):
    """cyto_path.py:181-187: (velocity_graph + velocity_graph_neg) restricted to the pattern of T_forward."""
    transmat = adata.uns["T_forward"].copy()
    transmat[transmat.nonzero()] = 1
    vg = adata.uns["velocity_graph"] + adata.uns["velocity_graph_neg"]
    return vg.multiply(transmat).tocsr()


def directionality_score(adata, neighborhood_sequence, end_point, map_state):
    """cyto_path.py:155-243 with the joblib fan-out (:239) run in order."""
    final_cluster = adata.uns["trajectories"]["trajectories_coordinates"][end_point]
    vg = masked_velocity_graph(adata)
    map_state = np.asarray(map_state)
    all_scores = []
    for i in range(len(final_cluster)):
        coords = final_cluster["trajectory_{}_coordinates".format(i)]
        steps = neighborhood_sequence[i]
        last = len(steps) - 1
        out = []
        for j in range(len(steps)):
            cells = steps[j]
            score = np.zeros(len(cells))
            fwd = (coords[j + 1] - coords[j]) if j != last else None            # :192,:206
            bwd = (coords[j] - coords[j - 1]) if j != 0 else None               # :207,:225
            rows = vg[cells]
            for k in range(len(cells)):
                row = rows[k, :]
                jumps = row.nonzero()[1]
                diff = map_state[jumps] - map_state[cells[k]]
                w = np.expm1(row.data)
                with np.errstate(invalid="ignore", divide="ignore"):
                    sf = np.average(cosine_similarity(fwd, diff) * w) if fwd is not None and len(jumps) else np.nan
                    sb = np.average(cosine_similarity(bwd, diff) * w) if bwd is not None and len(jumps) else np.nan
                if j == 0:
                    score[k] = sf
                elif j == last:
                    score[k] = sb
                else:
                    score[k] = max(sf, sb)                                       # :221
            out.append(score)
        all_scores.append(out)
    return all_scores
