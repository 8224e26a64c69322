"""Golden vectors for f4 from the REAL reference: extracts `directionality_score` (cyto_path.py:155-243) and the
numba cosine (utils.py:14-39) from /root/reference by source (the modules themselves import scvelo / anndata,
which are absent), runs them on a small synthetic case and stores inputs + outputs in tests/golden/.
Run in the build container only:  python -m oracle.make_golden_directionality"""
import ast
import os
import sys

import numpy as np
from scipy import sparse

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/cytopath"


def extract(path, names):
    src = open(path).read()
    tree = ast.parse(src)
    parts = []
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            parts.append(ast.get_source_segment(src, node) if not node.decorator_list else
                         "\n".join(src.splitlines()[node.decorator_list[0].lineno - 1:node.end_lineno]))
    return "\n\n".join(parts)


def reference_functions():
    import numba
    from joblib import Parallel, delayed
    from tqdm import tqdm
    ns = {"np": np, "numba": numba, "Parallel": Parallel, "delayed": delayed, "tqdm": tqdm}
    exec(extract(os.path.join(REF, "utils.py"), {"cosine_similarity_numba", "cosine_similarity"}), ns)
    exec(extract(os.path.join(REF, "trajectory_estimation", "cyto_path.py"), {"directionality_score"}), ns)
    return ns["directionality_score"]


class Shim:
    def __init__(self, uns):
        self.uns = uns


def make_case(seed, n, dims, n_traj, n_steps, vdtype):
    rng = np.random.default_rng(seed)
    pos = rng.normal(size=(n, dims))
    k = 12
    T = sparse.random(n, n, density=k / n, random_state=seed, format="csr", dtype=np.float64)
    T.data[:] = rng.random(T.nnz) + 0.05
    T = T.tolil(); T[5, :] = 0; T = sparse.csr_matrix(T.tocsr()); T.eliminate_zeros(); T.sort_indices()   # one cell without transitions (score = nan)
    T = sparse.csr_matrix((T.data.copy(), T.indices.copy(), T.indptr.copy()), shape=T.shape)
    # velocity graphs: most of T's pattern (positive part), a third of it (negative part), plus entries outside T
    nz = T.data.size
    keep = rng.random(nz) < 0.8
    vg = sparse.csr_matrix((np.where(keep, rng.random(nz), 0.0), T.indices, T.indptr), shape=T.shape)
    vg.eliminate_zeros()
    vg = (vg + sparse.random(n, n, density=2.0 / n, random_state=seed + 1, format="csr")).tocsr()
    keepn = rng.random(nz) < 0.33
    vgn = sparse.csr_matrix((np.where(keepn, -rng.random(nz), 0.0), T.indices, T.indptr), shape=T.shape)
    vgn.eliminate_zeros()
    vg, vgn = vg.astype(vdtype), vgn.astype(vdtype)
    coords = {"trajectory_%d_coordinates" % i: np.cumsum(rng.normal(size=(n_steps, dims)), axis=0) for i in range(n_traj)}
    nbhd = [[rng.choice(n, size=int(rng.integers(1, 25)), replace=False) for _ in range(n_steps)] for _ in range(n_traj)]
    uns = {"T_forward": T, "velocity_graph": vg, "velocity_graph_neg": vgn,
           "trajectories": {"trajectories_coordinates": {"end": coords}}}
    return Shim(uns), nbhd, pos


def main():
    fn = reference_functions()
    out_dir = os.path.join(ROOT, "tests", "golden")
    for name, args in {"directionality_d2_f32": (1, 300, 2, 2, 8, np.float32), "directionality_d10_f64": (2, 400, 10, 3, 5, np.float64)}.items():
        ad, nbhd, pos = make_case(*args)
        scores = fn(ad, nbhd, "end", pos, num_cores=1)
        flat_cells = np.concatenate([np.concatenate(t) for t in nbhd])
        flat_scores = np.concatenate([np.concatenate(t) for t in scores])
        sizes = np.array([[len(s) for s in t] for t in nbhd])
        T, vg, vgn = ad.uns["T_forward"], ad.uns["velocity_graph"], ad.uns["velocity_graph_neg"]
        coords = np.stack([ad.uns["trajectories"]["trajectories_coordinates"]["end"]["trajectory_%d_coordinates" % i] for i in range(len(nbhd))])
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), pos=pos, coords=coords, sizes=sizes, cells=flat_cells, scores=flat_scores,
                            T_data=T.data, T_indices=T.indices, T_indptr=T.indptr, T_shape=np.array(T.shape),
                            vg_data=vg.data, vg_indices=vg.indices, vg_indptr=vg.indptr,
                            vgn_data=vgn.data, vgn_indices=vgn.indices, vgn_indptr=vgn.indptr)
        print(name, flat_scores.shape, float(np.nanmean(flat_scores)), int(np.isnan(flat_scores).sum()))


if __name__ == "__main__":
    main()
