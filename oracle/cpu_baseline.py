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


def prepare(C, cdf_mode, memmap=True):
    """ELL arrays of matrix_prep (:22-28).  With `memmap` they are written once to a temp
    directory and re-opened as np.memmap: joblib then hands workers a file name instead of
    hashing + dumping ~0.8 GB on every Parallel call (a ~6 s fixed cost per call at 1M cells
    that the reference's own call pattern, :306-308, pays; excluding it favours the baseline)."""
    cdf = cwalk.build_cdf(C, cdf_mode)
    idx_ell, cdf_ell = sp.to_ell(C, cdf)
    _, prob_ell = sp.to_ell(C, C.data.astype(np.float32))
    out = dict(C=C, cdf=cdf, idx_ell=idx_ell, cdf_ell=cdf_ell, prob_ell=prob_ell, tmpdir=None)
    if memmap and idx_ell.nbytes > (8 << 20):
        import tempfile
        import joblib
        base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
        try:
            tmp = tempfile.mkdtemp(prefix="cyp_cpu_baseline_", dir=base)
            for k in ("idx_ell", "cdf_ell", "prob_ell"):
                f = os.path.join(tmp, k + ".pkl")
                joblib.dump(out[k], f)
                out[k] = joblib.load(f, mmap_mode="r")
            out["tmpdir"] = tmp
        except OSError:
            pass
    return out


def cleanup(prep):
    if prep.get("tmpdir"):
        import shutil
        for k in ("idx_ell", "cdf_ell", "prob_ell"):
            prep[k] = None
        shutil.rmtree(prep["tmpdir"], ignore_errors=True)


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
