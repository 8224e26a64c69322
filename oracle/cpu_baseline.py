"""CPU baselines for bench.py.  TEST / MEASUREMENT INFRASTRUCTURE ONLY.

`joblib_port_rate` times the faithful restatement of the reference's sampler
(`markov_sim_port`, same Python loop nest and numpy calls as
cytopath/markov_chain_sampler.py:35-52) fanned out over root cells with
joblib.Parallel exactly like :306-308, on all host cores, on a bounded sample.
`c_port_rate` times the compiled C restatement (oracle/walk.c, OpenMP).
"""
from __future__ import annotations

import os
import time

import numpy as np
from joblib import Parallel, delayed

from . import cwalk, sampler_port as sp


def _cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def prepare(C, cdf_mode):
    cdf = cwalk.build_cdf(C, cdf_mode)
    idx_ell, cdf_ell = sp.to_ell(C, cdf)
    _, prob_ell = sp.to_ell(C, C.data.astype(np.float32))
    return dict(C=C, cdf=cdf, idx_ell=idx_ell, cdf_ell=cdf_ell, prob_ell=prob_ell)


def joblib_port_rate(prep, roots, clusters, max_steps, sims_per_root, n_jobs=None, pool=None):
    """walker-steps/s of the joblib fan-out (one task per root cell, :306-308)."""
    n_jobs = n_jobs or _cores()
    clusters = np.asarray(clusters, dtype=object)
    own = pool is None
    pool = pool or Parallel(n_jobs=n_jobs)
    t0 = time.perf_counter()
    pool(delayed(sp.markov_sim_port)(j, sims_per_root, max_steps, roots, clusters, None, prep["idx_ell"],
                                     prep["cdf_ell"], None, prep["prob_ell"]) for j in range(len(roots)))
    dt = time.perf_counter() - t0
    steps = len(roots) * sims_per_root * (max_steps - 1)
    return steps / dt, dt, n_jobs


def c_port_rate(prep, roots, max_steps, sims_per_root, nthreads=0):
    starts = np.repeat(np.asarray(roots), sims_per_root)
    t0 = time.perf_counter()
    cwalk.walk(prep["C"], prep["cdf"], starts, max_steps - 1, seed=1, nthreads=nthreads, strict=False)
    dt = time.perf_counter() - t0
    return len(starts) * (max_steps - 1) / dt, dt, (nthreads or _cores())
