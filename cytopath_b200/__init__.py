"""cytopath_b200 — B200-native drop-in for cytopath's Markov-chain trajectory sampler.

    from cytopath_b200 import sampling
    sampling(adata, ...)            # same call surface as cytopath.sampling

Only the hot path of aron0093/cytopath is re-implemented (`cytopath.sampling`,
reference cytopath/markov_chain_sampler.py); everything downstream (trajectory
estimation, clustering, plotting) runs unchanged on `adata.uns['samples']`.
The compute lives in a CUDA C-ABI library (include/cytopath_b200.h); there is no CPU
fallback: without the library or a CUDA device, `sampling` raises.
"""
from .alignment import directionality_score
from .sampler import check_convergence_criteria, iterate_state_probability, sampling
from .trajectory_distances import coordinate_assigner, distance_matrix
from .undirected import undirected_simulations

__version__ = "0.1.0"
__all__ = ["sampling", "undirected_simulations", "distance_matrix", "coordinate_assigner", "directionality_score", "iterate_state_probability", "check_convergence_criteria", "__version__"]
