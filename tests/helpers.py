"""Shared test helpers: load golden fixtures back into the shapes the sampler takes."""
import json
import os

import numpy as np
import pandas as pd
from scipy import sparse

from cytopath_b200.synth import AnnDataShim

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

SAMPLING_CASES = [
    "sampling_small_clusters", "sampling_small_probs_f32", "sampling_small_notunique",
    "sampling_small_single_root", "sampling_small_auto", "sampling_two_stems_normalize",
    "sampling_tiny_dups", "sampling_small_dup_endpoints", "sampling_config1_auto",
]


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def csr_from(z, prefix="T_"):
    return sparse.csr_matrix((z[prefix + "data"], z[prefix + "indices"], z[prefix + "indptr"]),
                             shape=tuple(z[prefix + "shape"]))


def adata_from_case(z):
    kw = json.loads(str(z["kwargs"]))
    obs = pd.DataFrame({kw["cluster_key"]: pd.Categorical(z["clusters"]),
                        "root_cells": z["root_prob"], "end_points": z["end_prob"]})
    ad = AnnDataShim(len(z["clusters"]), obs, {"T_forward": csr_from(z)})
    return ad, kw
