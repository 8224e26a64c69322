#!/usr/bin/env python
"""Benchmark of the cytopath sampler hot path on B200 (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c4|c3|c2|c1]
    torchrun --nproc-per-node N bench.py --gpus N ...        (one rank per GPU, NCCL)

A "step" is one pass of the hot path over one batch: every rank advances its share of a
sampling round (walkers_per_gpu walkers x (max_steps-1) transitions) through the step kernel K2
with the graph resident in HBM.  `value` = walker-steps/s over all ranks (CUDA events, max over
ranks).  `e2e` = the same metric through the C ABI with HOST buffers: upload of the CSR from
pinned host memory + K1 + K2 + K3 + read-back of the kept index and a sample of kept rows.
`--impl reference` times the reference's CPU algorithm (faithful joblib port, all host cores).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("TQDM_DISABLE", "1")

WORKLOADS = {
    # name: (cells, k, branches, root stems, walkers per GPU, max_steps)  -- BASELINE.json configs
    "c1": dict(n=3_696, k=30, branches=4, stems=1, walkers=444_000, max_steps=108, desc="configs[0] graph: 3,696 cells k=30"),
    "c2": dict(n=5_780, k=30, branches=5, stems=2, walkers=710_000, max_steps=200, desc="configs[1] graph: 5,780 cells k=30, 200 steps"),
    "c3": dict(n=100_000, k=30, branches=8, stems=1, walkers=1_000_000, max_steps=200, desc="configs[2]: 100k cells k=30, 1M walkers x 200 steps (L2-resident)"),
    "c4": dict(n=1_000_000, k=50, branches=8, stems=1, layers=16, walkers=1_250_000, max_steps=500,
               desc="configs[3]: 1M cells k=50 (51M nnz), 10M walkers x 500 steps / 8 GPUs = 1.25M walkers per GPU (HBM-bound); branching "
                    "graph of SURVEY 8d with 8 branches x 16 interleaved kNN layers instead of B=16 (terminals stay reachable in 500 steps)"),
    "c5": dict(n=250_000, k=200, branches=0, stems=1, layers=16, walkers=1_250_000, max_steps=200, api_walkers_per_gpu=12_500_000,
               desc="configs[4]: cycling + convergent graph with self-loops, 250k cells k=200 (50M nnz, long rows), roots / end points from the 0.99 "
                    "probability thresholds; kernel loop 1.25M walkers x 200 steps per GPU, e2e_api 12.5M walkers per GPU (100M on 8 GPUs)"),
    "tiny": dict(n=20_000, k=30, branches=4, stems=1, walkers=100_000, max_steps=100, desc="developer smoke size"),
}


def env_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def make_graph(w, world):
    """Synthetic graph of a workload (seeded: identical on every rank).  BENCH_GRAPH_CACHE=<dir> keeps the generated
    graph on disk so that the processes of an A/B run or of a torchrun launch do not each spend a minute on the kNN search."""
    import pickle
    from cytopath_b200 import synth
    from cytopath_b200.sampler import _canonical_csr
    cache = os.environ.get("BENCH_GRAPH_CACHE")
    path = os.path.join(cache, "graph_%s.pkl" % "_".join("%s%s" % (k, w[k]) for k in ("n", "k", "branches", "stems")) +
                        ("_l%d" % w["layers"] if "layers" in w else "")) if cache else None
    if path and os.path.exists(path):
        with open(path, "rb") as f:
            return pickle.load(f)
    workers = max(1, host_cores() // world)
    if w["branches"] == 0:
        g = synth.cycling_graph(n=w["n"], k=w["k"], seed=0, workers=workers, n_layers=w.get("layers", 1))
        roots = np.where(g.root_prob >= 0.99)[0].astype(np.int32)        # unsupervised roots (markov_chain_sampler.py:226)
    else:
        g = synth.branching_graph(n=w["n"], k=w["k"], n_branches=w["branches"], n_root_stems=w["stems"], seed=0, workers=workers,
                                  n_layers=w.get("layers", 1))
        roots = np.where(np.isin(g.clusters, g.root_clusters))[0].astype(np.int32)
    C = _canonical_csr(g.T)
    if path:
        os.makedirs(cache, exist_ok=True)
        tmp = "%s.%d.tmp" % (path, os.getpid())
        with open(tmp, "wb") as f:
            pickle.dump((g, C, roots), f, protocol=4)
        os.replace(tmp, path)
    return g, C, roots


class ClockSampler:
    """nvidia-smi clocks line of B200_PROFILING.md, sampled every 50 ms.  nvidia-smi needs ~1 s to attach (it touches
    every GPU of the box, which visibly stalls kernels running on them), so it is started BEFORE the warm-up and the
    samples are cut to the timed region afterwards by their timestamps."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.proc, self.lines, self.thread = gpu, None, [], None

    def start(self, wait_s=5.0):
        import threading
        if os.environ.get("CYP_NO_CLOCKS"):
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, bufsize=1)
        except Exception:
            self.proc = None
            return

        def pump():
            for line in self.proc.stdout:
                self.lines.append(line)
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()
        t0 = time.time()
        while not self.lines and time.time() - t0 < wait_s:      # attached and sampling before anything is timed
            time.sleep(0.02)

    def stop(self, t_begin=None, t_end=None):
        import datetime
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        if self.thread is not None:
            self.thread.join(timeout=2)
        rows = []
        for line in list(self.lines):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[2]), float(f[3]), float(f[4]), f[6:10]))
            except ValueError:
                continue
        inside = [r for r in rows if t_begin is not None and t_begin - 0.03 <= r[0] <= t_end + 0.03]
        where = "timed region"
        if not inside and rows and t_begin is not None:          # region shorter than the sampling period: nearest samples
            inside = sorted(rows, key=lambda r: abs(r[0] - 0.5 * (t_begin + t_end)))[:2]
            where = "nearest to the timed region"
        if not inside:
            inside, where = rows, "whole run"
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        reasons = set()
        for r in inside:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median([r[1] for r in inside])), "sm_max_mhz": float(max(r[2] for r in inside)),
                "reasons": sorted(reasons), "samples": len(inside), "samples_from": where,
                "power_w_max": float(max(r[3] for r in inside))}


def mean_visited_row_len(C, seq_sample):
    lens = np.diff(C.indptr)
    return float(lens[seq_sample[:, :-1].ravel()].mean())


def fused_layout_stats(C, seq_sample, row_block_bytes, n_probe=20000):
    """What the fused-row kernel requests per walker-step on the sampled trajectories: mean bytes of the visited
    row blocks and the share of transitions whose target was inlined (include/cytopath_b200.h, DESIGN.md 4:
    block = 20-byte header + 12-bit codes + the m most probable targets that fit the block's last sector)."""
    lens = np.diff(C.indptr)
    used = (20 + lens + (lens + 1) // 2 + 3) & ~3
    nsect0 = (used + 31) // 32
    extra = 1 if row_block_bytes > int(nsect0.sum()) * 32 else 0                # the library's choice (rows_v2.cu)
    nsect = nsect0 + extra * (lens > 0)
    m = np.minimum(np.minimum(lens, 64), (nsect * 32 - used) // 4)
    cur, nxt = seq_sample[:, :-1].ravel(), seq_sample[:, 1:].ravel()
    pick = np.random.default_rng(0).choice(len(cur), size=min(n_probe, len(cur)), replace=False)
    inlined = 0
    for c, x in zip(cur[pick], nxt[pick]):
        a, b = C.indptr[c], C.indptr[c + 1]
        j = int(np.searchsorted(C.indices[a:b], x))
        if j < 64:
            cand = C.data[a:min(b, a + 64)]
            rank = int(np.sum((cand > cand[j]) | ((cand == cand[j]) & (np.arange(len(cand)) < j))))
            inlined += rank < m[c]
    return float((nsect[cur] * 32).mean()), float(inlined) / len(pick)


def timed_port_sample(cb, prep, pool, roots, clusters, max_steps, cores, budget_s):
    """Size one joblib call (one task per root cell, 2 tasks per core, like :306-308) to about
    `budget_s` of wall time.  The per-walker cost is measured in-process first so that each task
    steps for ~budget_s/2: the per-task argument pickling of the reference's call pattern (the
    label array travels with every task) is then a few percent of the timed call, as it would
    be at the reference's own task sizes.  Returns (roots, sims, rate, wall, dispatch_wall)."""
    import time as _t
    from oracle import sampler_port as sp
    sample_roots = np.resize(roots, 2 * cores)
    cl = np.asarray(clusters, dtype=object)
    t0 = _t.perf_counter()
    sp.markov_sim_port(0, 4, max_steps, sample_roots, cl, None, prep["idx_ell"], prep["cdf_ell"], None, prep["prob_ell"])
    per_walker = (_t.perf_counter() - t0) / 4.0
    sims = int(min(65536, max(1, budget_s * cores / (len(sample_roots) * per_walker))))
    cb.joblib_port_rate(prep, sample_roots[:cores], clusters, min(max_steps, 20), 1, cores, pool)        # warm the pool
    _, dt1, _ = cb.joblib_port_rate(prep, sample_roots, clusters, min(max_steps, 3), 1, cores, pool)      # dispatch only
    rate, dt, _ = cb.joblib_port_rate(prep, sample_roots, clusters, max_steps, sims, cores, pool)
    return sample_roots, sims, rate, dt, dt1


def cpu_baseline_sample(C, g, roots, max_steps, cdf_mode, budget_s=12.0):
    """Bounded sample of the same workload on the host cores: faithful joblib port (kind 'port')."""
    from joblib import Parallel
    from oracle import cpu_baseline as cb
    cores = host_cores()
    prep = cb.prepare(C, cdf_mode)
    try:
        with Parallel(n_jobs=cores) as pool:
            sample_roots, sims, rate, dt, dt1 = timed_port_sample(cb, prep, pool, roots, g.clusters, max_steps, cores, budget_s)
        c_rate, c_dt, c_thr = cb.c_port_rate(prep, np.resize(roots, 4096), max_steps, max(1, int(2e8 / (4096 * max_steps))))
    finally:
        cb.cleanup(prep)
    sample = "%d root cells x %d walkers x %d steps, joblib(n_jobs=%d) over root cells, %.1f s (of which %.1f s task dispatch)%s" % (
        len(sample_roots), sims, max_steps, cores, dt, dt1, ", ELL arrays pre-memmapped" if prep.get("tmpdir") else "")
    return ({"value": rate, "unit": "walker-steps/s", "cores": cores, "kind": "port", "sample": sample},
            {"value": c_rate, "unit": "walker-steps/s", "cores": c_thr, "kind": "port-c-openmp",
             "sample": "oracle/walk.c, %.2f s" % c_dt})


def base_config(w, args, nnz=None):
    """The workload, with the same keys in both arms (the driver compares them)."""
    W = args.walkers or w["walkers"]
    return {"workload": w["desc"], "cells": w["n"], "k": w["k"], "nnz": nnz, "walkers_per_gpu": W, "max_steps": w["max_steps"],
            "cdf_mode": args.cdf_mode, "seed": args.seed}


def run_reference(args, w):
    """Reference arm: the reference's CPU algorithm for this path (joblib fan-out of the
    pure-Python step loop, markov_chain_sampler.py:35-52,:306-308) on all host cores."""
    rank, _, world = env_world()
    if rank != 0:
        return
    from joblib import Parallel
    from oracle import cpu_baseline as cb
    g, C, roots = make_graph(w, 1)
    cores = host_cores()
    mode = 1 if args.cdf_mode == "exact" else 0
    prep = cb.prepare(C, mode)
    max_steps = w["max_steps"]
    times = []
    try:
        with Parallel(n_jobs=cores) as pool:
            sample_roots, sims, _, _, _ = timed_port_sample(cb, prep, pool, roots, g.clusters, max_steps, cores, 5.0)   # ~5 s of stepping per step
            for i in range(args.warmup + args.steps):
                rate, dt, _ = cb.joblib_port_rate(prep, sample_roots, g.clusters, max_steps, sims, cores, pool)
                if i >= args.warmup:
                    times.append(dt)
    finally:
        cb.cleanup(prep)
    steps_per = len(sample_roots) * sims * (max_steps - 1)
    value = steps_per * len(times) / sum(times)
    sample = "%d root cells x %d walkers x %d steps per step, joblib(n_jobs=%d)" % (len(sample_roots), sims, max_steps, cores)
    line = {
        "impl": "reference", "metric": "walker_steps_per_sec", "value": value, "unit": "walker-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64" if mode else "f32", "data": "synthetic",
        "config": base_config(w, args, int(C.nnz)),
        "cpu_baseline": {"value": value, "unit": "walker-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "walker-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def kernel_sass_sha16(kernel_name):
    """Identity of the step kernel's machine code: sha256 of the SASS of the object file that holds it (cuobjdump; the
    object bytes themselves are not reproducible, their SASS is).  profiles/ncu_traffic.json stores the same hash next
    to every capture, so a capture only feeds `roofline` while the kernel it measured is the kernel that runs."""
    import hashlib
    obj = "walk_v2_exact.o" if "q12/f64" in kernel_name else "walk_v2_ref.o" if "q12/f32" in kernel_name else "walk_tma.o"
    path = os.path.join(ROOT, "cytopath_b200", "lib", "obj", obj)
    try:
        out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, timeout=120).stdout
    except Exception:
        return None
    keep = [l for l in out.splitlines() if l.strip() and not any(k in l for k in ("Fatbin", "===", "arch =", "code version", "host =", "compile_size", "identifier", "producer"))]
    return hashlib.sha256("\n".join(keep).encode()).hexdigest()[:16] if keep else None


def pinned_csr(C):
    """The CSR arrays in page-locked host memory (what a caller that cares about upload time would hand over)."""
    import torch
    from scipy import sparse
    ptr_ = torch.from_numpy(np.ascontiguousarray(C.indptr, dtype=np.int64)).pin_memory()
    col = torch.from_numpy(np.ascontiguousarray(C.indices, dtype=np.int32)).pin_memory()
    val = torch.from_numpy(np.ascontiguousarray(C.data, dtype=np.float64)).pin_memory()
    P = sparse.csr_matrix((val.numpy(), col.numpy(), ptr_.numpy()), shape=C.shape, copy=False)
    return P, (ptr_, col, val)


def run_b200(args, w):
    import torch
    import torch.distributed as dist
    import cytopath_b200
    from cytopath_b200 import _lib

    rank, local_rank, world = env_world()
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")          # keep NCCL's version banner off stdout (one JSON line only)
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
        torch.cuda.synchronize()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    _lib.require_device(local_rank)
    L = _lib.load()
    mode = _lib.CDF_EXACT if args.cdf_mode == "exact" else _lib.CDF_REFERENCE

    def barrier():
        if world > 1:
            dist.barrier()

    # The group first: its NCCL communicator (ncclCommInitRank; the id travels through torch's store) connects its
    # transports before the graph and the round buffers are allocated (profiles/r01_multi_gpu_probe.md).
    group = _lib.open_group(device=local_rank, dist=world > 1)
    assert group.world == world and group.rank == rank

    # rank 0 generates (or loads) the synthetic graph; the others read it from the cache it leaves behind
    os.environ.setdefault("BENCH_GRAPH_CACHE", "/tmp/cyp_bench_graphs")
    if rank == 0:
        g, C, roots = make_graph(w, 1)
    barrier()
    if rank != 0:
        g, C, roots = make_graph(w, 1)
    max_steps, nt = w["max_steps"], w["max_steps"] - 1
    W = args.walkers or w["walkers"]
    total = W * world
    n_roots = len(roots)
    sims_per_root = -(-total // n_roots)
    w0 = rank * W

    # ---- resident state: graph, roots, trajectory buffer.  Every rank holds the host CSR (as every torchrun process of a
    # real run holds the AnnData): each uploads 1/world of it from pinned memory, an all-gather over NVLink completes it,
    # every device runs K1.  BENCH_ROOT_UPLOAD=1: rank 0 uploads everything and broadcasts the built arrays instead.
    sharded = world > 1 and not os.environ.get("BENCH_ROOT_UPLOAD")
    Cp, _pins = pinned_csr(C) if (rank == 0 or sharded) else (None, None)
    gstat = group.create_graph(Cp, mode, 0, shape=(C.shape[0], C.nnz), sharded=sharded)
    graph = group.graph
    info = graph.info()
    kernel_name = graph.variant_name(args.variant)
    d_roots = torch.tensor(roots, dtype=torch.int32, device=dev)
    d_seq = torch.empty((W, max_steps), dtype=torch.int32, device=dev)
    d_miss = torch.zeros(1, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream()
    sptr = ctypes.c_void_p(stream.cuda_stream)

    def step(i):
        _lib.check(L.cyp_walk_device(graph.handle, d_roots.data_ptr(), n_roots, sims_per_root, w0, W, nt, args.seed, i,
                                     None, d_seq.data_ptr(), d_miss.data_ptr(), args.variant, sptr))

    clocks = ClockSampler(local_rank)
    clocks.start()                                               # attached before the warm-up, see ClockSampler
    barrier()
    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.time()
    e0.record(stream)
    for i in range(args.steps):
        step(args.warmup + i)
    e1.record(stream)
    torch.cuda.synchronize()
    t_end = time.time()
    barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop(t_begin, t_end)
    if world > 1:
        print("rank %d: %.3f ms per step (device time of its own K2 launches)" % (rank, ms / args.steps), file=sys.stderr)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = total * nt * args.steps / (ms_max * 1e-3)
    sample = d_seq[:: max(1, W // 2000)].cpu().numpy()
    kbar = mean_visited_row_len(C, sample)
    elem = 8 if mode == _lib.CDF_EXACT else 4
    bytes_per_step = elem * kbar + 16.0
    ms_launch = ms / args.steps                                  # this rank's kernel: one launch per step
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    else:
        peak, peak_src = 6650.0, "fallback of B200_PROFILING.md"
    alg_achieved = W * nt * bytes_per_step / (ms_launch * 1e-3) / 1e9
    lay = graph.layout()
    inline_cov, layout_bytes = None, None
    if "fused rows" in kernel_name:
        block_bytes, inline_cov = fused_layout_stats(C, sample, lay["row_block_bytes"])
        layout_bytes = block_bytes + (1.0 - inline_cov) * 32.0 + 4.0      # row block + a target-word sector when not inlined + the store
    assert int(d_miss.item()) == 0, "walkers hit rows without a CDF crossing"

    if args.kernel_only:
        if rank == 0:
            emit({"metric": "walker_steps_per_sec", "value": value, "ms_per_step": ms_max / args.steps,
                  "kernel": kernel_name, "algorithmic_frac": alg_achieved / peak, "kernel_only": True,
                  "kernel_sass_sha16": kernel_sass_sha16(kernel_name)})
        group.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- the roofline the kernel cannot beat: measured DRAM / L2 traffic of the committed ncu capture (only while the
    # kernel's machine code is the captured one) and the device's random-gather ceiling at the same footprint, measured now
    roof = {"bound": "hbm", "peak": peak, "unit": "GB/s", "peak_source": peak_src, "ms_per_launch": ms_launch,
            "algorithmic_equiv": {"achieved": alg_achieved, "frac": alg_achieved / peak, "bytes_per_walker_step": bytes_per_step,
                                  "mean_visited_row_nnz": kbar,
                                  "note": "SURVEY 8d's bytes (the full float64/float32 CDF row + 16 B per step) over the launch time: "
                                          "what a kernel reading the reference's layout would have to move; the fused rows (12-bit codes, "
                                          "inlined targets) request layout_bytes_per_walker_step instead, so this is not a bound"},
            "layout_bytes_per_walker_step": layout_bytes, "inlined_target_share": inline_cov}
    sha = kernel_sass_sha16(kernel_name) if rank == 0 else None
    roof["kernel_sass_sha16"] = sha
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if rank == 0 and os.path.exists(tpath):
        rec = json.load(open(tpath)).get("%s|%s" % (args.workload, kernel_name))
        if rec and W == w["walkers"] and rec.get("kernel_sass_sha16") == sha and sha is not None:
            traffic = rec["dram_bytes_read"] + rec["dram_bytes_write"]
            roof["traffic_source"] = rec.get("capture")
            if rec.get("l2_bytes"):
                roof["l2_bytes_per_launch"] = rec["l2_bytes"]
                roof["l2_achieved_gbs"] = rec["l2_bytes"] / (ms_launch * 1e-3) / 1e9
                roof["l2_frac"] = (rec.get("lts_sectors_pct") or 0.0) / 100.0 or None
                roof["l2_hit_rate"] = (rec.get("lts_hit_rate_pct") or 0.0) / 100.0 or None
                roof["l2_frac_source"] = "lts__t_sectors.sum.pct_of_peak_sustained_elapsed of the capture (L2 sector bandwidth used)"
        elif rec:
            roof["traffic_source"] = "capture %s is of another build of the kernel (sass %s): not used" % (rec.get("capture"), rec.get("kernel_sass_sha16"))
    roof["traffic"] = traffic
    if traffic is not None:
        roof["achieved"] = traffic / (ms_launch * 1e-3) / 1e9          # DRAM bytes of one launch (ncu) over this run's launch time
        roof["frac"] = roof["dram_frac"] = roof["achieved"] / peak
        roof["dram_bytes_per_walker_step"] = traffic / (W * nt)
        roof["note"] = ("achieved = measured DRAM bytes per launch / launch time: the fraction of the HBM roofline the kernel uses. It is a "
                        "dependent random-gather kernel: gather_ceiling_frac says how close it is to what the memory system gives that "
                        "access pattern; algorithmic_equiv is the SURVEY 8d figure")
    else:
        roof["achieved"], roof["frac"], roof["dram_frac"] = alg_achieved, alg_achieved / peak, None
        roof["note"] = "no ncu capture of this build of the kernel: achieved / frac are the algorithmic (SURVEY 8d) figures, not a bound"
    if "fused rows" in kernel_name and rank == 0:
        fp = int(lay["row_block_bytes"])
        fp = max(fp, 1 << 20)                                     # tiny graphs: the smallest working set the probe takes
        ceil_g = _lib.gather_ceiling(fp, 128, local_rank)
        roof["gather_ceiling"] = {"gathers_per_s": ceil_g, "working_set_bytes": fp, "gather_bytes": 128,
                                  "how": "cyp_gather_ceiling: 8 independent 128-byte random gathers in flight per 8 lanes, same footprint, same run"}
        roof["gather_ceiling_frac"] = (W * nt / (ms_launch * 1e-3)) / ceil_g

    # ---- e2e through the C ABI with HOST buffers, every step: the CSR from pinned host memory (each rank its 1/world slice)
    # -> NVLink all-gather -> K1 on every device -> per-device session -> one sampling round (K2 + K3, all-gather of the 24-byte
    # records, cross-rank first-occurrence prune, all-gather of the index) -> final selection (replay of the picked
    # walkers split over the ranks, all-gather of the int32 sequences) -> rows + scores on the host
    trunc = np.array([c[:8] for c in g.clusters], dtype="U8")
    code_names, cell_code = np.unique(trunc, return_inverse=True)
    cell_code = np.ascontiguousarray(cell_code, dtype=np.int32)
    end_cells = np.where(np.isin(g.clusters, g.end_clusters))[0] if w["branches"] else np.where(g.end_prob >= 0.99)[0]
    end_slot = np.full(C.shape[0], -1, dtype=np.int32)
    end_slot[end_cells] = np.arange(len(end_cells), dtype=np.int32)
    n_fetch = 2000
    del d_seq
    torch.cuda.empty_cache()
    e2e_ms = {}
    last = {}

    def lap(name, t0):
        dt = (time.perf_counter() - t0) * 1e3
        e2e_ms[name] = e2e_ms.get(name, 0.0) + dt
        return time.perf_counter()

    def e2e_step(i, grp, root_csr, per_rank_walkers):
        t0 = time.perf_counter()
        gs = grp.create_graph(root_csr, mode, 0, shape=(C.shape[0], C.nnz), sharded=sharded and grp.world > 1)
        t0 = lap("graph_create", t0)
        grp.create_session(roots, cell_code, len(code_names), 2, end_slot, len(end_cells), max_steps, True, args.seed)
        t0 = lap("session_create", t0)
        sims = -(-(per_rank_walkers * grp.world) // n_roots)
        st, counts = grp.round(i, sims)
        t0 = lap("round", t0)
        have = np.flatnonzero(counts)
        m = min(n_fetch, int(st["n_kept"]))
        if m:
            pick = have[np.arange(m) % len(have)].astype(np.int32)
            off = np.minimum(np.arange(m) // len(have), counts[pick] - 1).astype(np.int64)
            seq, score, sel = grp.select(pick, off)
        else:
            seq, sel = np.empty((0, max_steps), np.int32), {}
        lap("select", t0)
        last.update(graph=gs, round=st, select=sel, d2h=int(counts.nbytes + seq.nbytes + 4 * m))
        return st

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    e2e_step(0, group, Cp, W)                                      # warm-up
    e2e_ms.clear()
    torch.cuda.synchronize(); barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        st = e2e_step(1 + i, group, Cp, W)
    torch.cuda.synchronize(); barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = total * nt * e2e_steps / float(t.item())
    # whole job: the column / value arrays once (one slice per rank), the row offsets and the session arrays on every rank
    h2d = C.nnz * 12 + (C.indptr.shape[0] * 8 * (world if sharded else 1)) + (roots.nbytes + cell_code.nbytes + end_slot.nbytes) * world
    digest_n = group.digest()

    # ---- parity self-check of the multi-GPU path: the global index of kept sequences of the last e2e round must be the one
    # a single device computes for the same walker range (128-bit order-sensitive digest over (round, walker, slot))
    multi = None
    if world > 1:
        dg = torch.tensor([digest_n[0] & 0x7FFFFFFFFFFFFFFF, digest_n[1] & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64, device=dev)
        gathered = [torch.empty_like(dg) for _ in range(world)]
        dist.all_gather(gathered, dg)
        same_on_all = all(bool(torch.equal(x, gathered[0])) for x in gathered)
        single_ok = None
        if rank == 0:
            solo = _lib.Group(devices=[local_rank])
            e2e_keep, last_keep = dict(e2e_ms), dict(last)
            e2e_step(e2e_steps, solo, Cp, W * world)              # the same round index over the whole walker range, one device
            single_ok = solo.digest() == digest_n
            solo.close()
            e2e_ms.clear(); e2e_ms.update(e2e_keep); last.clear(); last.update(last_keep)
        r, s_ = last["round"], last["select"]
        gl = last["graph"]
        if gl.get("sharded"):
            graph_part = {"graph": "every rank uploads 1/%d of the CSR, ncclAllGather completes it, K1 on every device" % world,
                          "graph_upload_ms": gl["ms_upload"], "graph_gather_ms": gl["ms_gather"],
                          "graph_gather_bytes_per_rank": int(C.nnz * 12 * (world - 1) // world)}
        else:
            graph_part = {"graph": "rank 0 uploads the CSR and runs K1, ncclBroadcast of the built arrays",
                          "graph_broadcast_ms": gl["ms_broadcast"], "graph_broadcast_bytes": int(info["device_bytes"])}
        multi = {"collectives": ("ncclAllGather(CSR slices)" if gl.get("sharded") else "ncclBroadcast(graph arrays)") +
                                " + per round ncclAllGather(stats), ncclAllGather(24-byte records), ncclAllGather(walker, slot) + "
                                "ncclAllGather(selected int32 sequences)",
                 **graph_part,
                 "round_exchange_bytes_per_rank": int(r["exchange_bytes"]), "round_exchange_ms": r["ms_exchange"], "round_local_ms": r["ms_local"],
                 "select_gather_bytes": int(s_.get("gather_bytes", 0)), "select_gather_ms": s_.get("ms_gather"),
                 "kept_global_last_round": int(r["n_kept"]), "replayed_last_round": int(r["n_replayed"]),
                 "parity": {"digest": "%016x%016x" % digest_n, "same_on_all_ranks": same_on_all, "equals_single_device_run": single_ok},
                 "nccl_version": group.nccl_version}
        barrier()
    group.close()

    # ---- e2e_api: the public call, cytopath_b200.sampling(adata, ...) with auto_adjust on this workload's AnnData
    api = None
    if not args.no_api:
        import contextlib
        import io
        ad = g.to_adata()
        kw = dict(root_clusters=list(g.root_clusters), end_clusters=list(g.end_clusters)) if w["branches"] else {}
        torch.cuda.synchronize(); barrier()
        np.random.seed(0)
        t0 = time.perf_counter()
        if "api_walkers_per_gpu" in w:       # a fixed first round: sim_number / n_roots walkers per root, doubled at the start of the round (:285,:299)
            kw.update(auto_adjust=False, sim_number=w["api_walkers_per_gpu"] * world // 2, max_steps=max_steps, traj_number=200)
        else:
            kw.update(auto_adjust=True)
        api_error = None
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                cytopath_b200.sampling(ad, cdf_mode=args.cdf_mode, seed=args.seed, device=local_rank, dist=world > 1, **kw)
        except ValueError as exc:            # e.g. the reference's "Sampling failed": the terminals of this synthetic graph are out of reach
            api_error = str(exc)
        api_s = time.perf_counter() - t0
        t = torch.tensor([api_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        b = ad.uns["run_info"].get("b200") if api_error is None else None
        if api_error is not None:
            api = {"error": api_error, "wall_s": float(t.item())}
    if not args.no_api and api is None:
        steps_api = sum(int(r["n_walkers"]) for r in b["rounds"]) * (ad.uns["run_info"]["simulation_steps"] - 1)
        api = {"value": steps_api / float(t.item()), "unit": "walker-steps/s", "wall_s": float(t.item()), "walker_steps": steps_api,
               "rounds": len(b["rounds"]), "max_steps": int(ad.uns["run_info"]["simulation_steps"]),
               "rows_out": int(ad.uns["samples"]["cell_sequences"].shape[0]),
               "host_ms": {k: round(v, 1) for k, v in b["timing"].get("host_ms", {}).items()},
               "ms": {**{"graph_" + k[3:]: v for k, v in b["timing"]["graph"].items() if k.startswith("ms_")},
                      "rounds_local": sum(r["ms_local"] for r in b["rounds"]), "rounds_exchange": sum(r["ms_exchange"] for r in b["rounds"]),
                      "k2": sum(r["ms_walk"] for r in b["rounds"]), "k3": sum(r["ms_filter"] for r in b["rounds"]),
                      "select": sum(v for k, v in b["select"].items() if k.startswith("ms_"))},
               "what": "cytopath_b200.sampling(adata, %s, %s) on the workload's AnnData (host scipy CSR in, uns['samples'] out); CUDA context already created" % (
                   "auto_adjust=True" if kw.get("auto_adjust") else "sim_number=%d, max_steps=%d" % (kw["sim_number"], max_steps),
                   "root_clusters/end_clusters" if w["branches"] else "root/end probabilities >= 0.99")}

    line = {
        "metric": "walker_steps_per_sec", "value": value, "unit": "walker-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64" if mode == _lib.CDF_EXACT else "f32", "data": "synthetic",
        "config": base_config(w, args, int(C.nnz)),
        "kernel": kernel_name, "global_walkers": total,
        "parallelism": ("walkers sharded by id over %d GPU(s); " % world) + (
            "CSR uploaded in slices (1/%d per rank), all-gathered over NVLink, K1 on every device" % world if sharded else
            "graph built on rank 0" + (" and broadcast over NVLink" if world > 1 else "")),
        "l2": "no flush: the step kernel walks %.2f GB of row blocks + target words and writes %.2f GB of trajectories per step, both above the 126 MB L2" % (
            (lay["row_block_bytes"] + lay["target_word_bytes"]) / 1e9, W * max_steps * 4 / 1e9),
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": "walker-steps/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(last["d2h"]) * world,
                "steps": e2e_steps, "kept_rows_last_step": int(st["n_kept"]), "ms_walk": st["ms_walk"], "ms_filter": st["ms_filter"],
                "host_ms_per_step": {k: round(v / e2e_steps, 2) for k, v in e2e_ms.items()},
                "what": ("cyp_group_graph_create_sharded (every rank: its slice of the pinned host CSR + NVLink all-gather + K1)" if sharded else
                         "cyp_group_graph_create (pinned host CSR on rank 0 + K1" + (" + NVLink broadcast)" if world > 1 else ")")) +
                        ", cyp_group_session_create, cyp_group_round "
                        "(K2 + K3 + NCCL exchange), cyp_group_select of %d sequences (replay + all-gather) with scores, to the host" % n_fetch},
        "gpu_launches": args.steps,
        "value_scope": "K2 launches only (communication-free: the path has no collective inside the step loop); the per-round "
                       "NCCL exchange, the graph upload + all-gather and the selection are timed in e2e / multi_gpu",
        "roofline": roof,
    }
    if multi is not None:
        line["multi_gpu"] = multi
    if api is not None:
        line["e2e_api"] = api
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu, cpu_c = cpu_baseline_sample(C, g, roots, max_steps, 1 if mode == _lib.CDF_EXACT else 0)
        line["cpu_baseline"] = cpu
        line["cpu_baseline_c"] = cpu_c
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def own_stdout():
    """stdout must carry ONE JSON line: everything else that writes to fd 1 (NCCL prints its version banner there,
    whatever NCCL_DEBUG says) is sent to stderr; emit() writes the line to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--cdf-mode", default="exact", choices=["exact", "reference"])
    ap.add_argument("--walkers", type=int, default=0, help="walkers per GPU (default: the workload's)")
    ap.add_argument("--variant", type=int, default=0, help="step-kernel variant lanes*100+unroll (0 = auto)")
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-api", action="store_true", help="skip the e2e_api leg (the public sampling() call)")
    ap.add_argument("--kernel-only", action="store_true", help="profiling runs: only the device-resident K2 loop")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        print("note: fewer than 3 warm-up steps requested", file=sys.stderr)
    w = WORKLOADS[args.workload]
    own_stdout()
    if args.impl == "reference":
        run_reference(args, w)
        return
    _, _, world = env_world()
    if args.gpus != world:
        print("note: --gpus %d but WORLD_SIZE=%d; using WORLD_SIZE" % (args.gpus, world), file=sys.stderr)
    run_b200(args, w)


if __name__ == "__main__":
    main()
