"""Several GPUs: the group layer (csrc/group.cu) gives the single-GPU / reference result
  * from ONE process driving two devices (`sampling(adata, devices=[0, 1])`, ncclCommInitAll), and
  * from one process per GPU under torchrun (ncclCommInitRank, the id through torch's store), with the ranks' numpy
    RNGs seeded DIFFERENTLY (the final draw, :395, must be rank 0's on every rank),
and the cross-rank first-occurrence prune is exact when hashes collide (hash truncated to a few bits).
Skipped on boxes with fewer than two CUDA devices (run with `gpurun --gpus 2`)."""
import contextlib
import ctypes
import io
import os
import subprocess
import sys
import warnings

import numpy as np
import pytest

import cytopath_b200
from cytopath_b200 import _lib, synth
from oracle import cwalk, sampler_port as sp
from tests.helpers import adata_from_case, load

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = ["sampling_tiny_dups", "sampling_small_clusters", "sampling_small_notunique", "sampling_config1_auto"]


def need_two():
    if _lib.device_count() < 2:
        pytest.skip("needs two CUDA devices")


def check(s, z):
    np.testing.assert_array_equal(s["cell_sequences"], z["cell_sequences"])
    np.testing.assert_array_equal(s["cluster_sequences"], z["cluster_sequences"])
    np.testing.assert_allclose(s["transition_score"], z["transition_score"], rtol=2e-6)


@pytest.mark.parametrize("case", CASES)
def test_one_process_two_devices_equals_reference(case):
    """A plain function call: no launcher, no torch.distributed."""
    need_two()
    z = load(case)
    ad, kw = adata_from_case(z)
    np.random.seed(int(z["np_seed"]))
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cytopath_b200.sampling(ad, **kw, seed=int(z["seed"]), devices=[0, 1])
    assert ad.uns["run_info"]["b200"]["world_size"] == 2
    check(ad.uns["samples"], z)


def test_group_index_is_independent_of_the_device_count_even_with_colliding_hashes():
    """Global index of kept sequences (round, walker, slot) of a two-device group == a one-device group == the oracle,
    with the 128-bit hash cut to 5 bits so that the cross-rank prune has to tell colliding sequences apart by replay."""
    need_two()
    C = sp.canonical_csr(synth.branching_graph(n=80, k=4, n_branches=2, seed=5, workers=1, scale=2.0, noise=0.05).T)
    n = C.shape[0]
    rng = np.random.default_rng(3)
    code = rng.integers(4, size=n).astype(np.int32)
    code[:4] = np.arange(4)
    roots = np.array([3, 7, 11, 20, 33])
    end_cells = np.sort(rng.choice(n, size=50, replace=False))
    slot = np.full(n, -1, dtype=np.int32)
    slot[end_cells] = np.arange(len(end_cells), dtype=np.int32)
    cdf = cwalk.build_cdf(C, 0)
    L = _lib.load()
    results = {}
    for devices in ([0], [0, 1]):
        for bits in (128, 5):
            grp = _lib.Group(devices=devices)
            grp.create_graph(C, _lib.CDF_REFERENCE)
            grp.create_session(roots, code, 4, 2, slot, len(end_cells), 6, True, 77)
            for i in range(len(devices)):
                _lib.check(L.cyp_session_set_hash_bits(ctypes.c_void_p(grp.session_handle(i)), bits))
            counts = []
            for rnd, sims in enumerate((40, 300)):
                st, c = grp.round(rnd, sims)
                counts.append(c)
            results[(len(devices), bits)] = (grp.fetch_index(), np.array(counts), grp.digest())
            if len(devices) == 2 and bits == 5:
                seq, score, _ = grp.select(np.repeat(np.flatnonzero(counts[1])[:5], 1), np.zeros(5, dtype=np.int64))
                assert seq.shape == (5, 6)
            grp.close()
    (r1, w1, s1), c1, d1 = results[(1, 128)]
    for key, ((r, w, s), c, d) in results.items():
        np.testing.assert_array_equal(r, r1, err_msg=str(key))
        np.testing.assert_array_equal(w, w1, err_msg=str(key))
        np.testing.assert_array_equal(s, s1, err_msg=str(key))
        np.testing.assert_array_equal(c, c1, err_msg=str(key))
        assert d == d1
    # and the oracle: first occurrences of round 1 in walker order
    starts = np.repeat(roots, 300)
    seq = cwalk.walk(C, cdf, starts, 5, seed=77, round_idx=1)
    cs = np.sort(code[seq], axis=1)
    ok = (1 + (np.diff(cs, axis=1) != 0).sum(axis=1) >= 2) & (slot[seq[:, -1]] >= 0)
    cand = np.flatnonzero(ok)
    cand = cand[np.sort(np.unique(seq[cand], axis=0, return_index=True)[1])]
    np.testing.assert_array_equal(w1[r1 == 1], cand)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("cdf_mode", [_lib.CDF_REFERENCE, _lib.CDF_EXACT])
def test_sharded_upload_builds_the_graph_the_root_rank_builds(dtype, cdf_mode):
    """cyp_group_graph_create_sharded (each device uploads half of the CSR, all-gather, K1 on both) == cyp_group_graph_create
    (device 0 uploads everything, broadcast of the built arrays) == the single-device graph: CDF, fused rows, row sums.
    nnz is odd, so the slices are ragged and the last one is padded."""
    need_two()
    C = sp.canonical_csr(synth.branching_graph(n=3001, k=7, n_branches=3, seed=11, workers=1, scale=2.0, noise=0.05).T)
    C = C.astype(dtype)
    if C.nnz % 2 == 0:                                  # drop one entry: an odd nnz
        C = C.tolil(); r = int(np.argmax(np.diff(C.tocsr().indptr))); C[r, C.rows[r][0]] = 0; C = sp.canonical_csr(C.tocsr()).astype(dtype)
    assert C.nnz % 2 == 1
    solo = _lib.Graph(C, cdf_mode, 0)
    want = (solo.cdf(), solo.fused(), solo.row_sum_range(), solo.layout())
    solo.close()
    for sharded in (True, False):
        grp = _lib.Group(devices=[0, 1])
        st = grp.create_graph(C, cdf_mode, sharded=sharded)
        assert bool(st.get("sharded")) == sharded
        for i in (0, 1):
            gv = grp.graph_view(i)
            assert gv.info()["device"] == i
            np.testing.assert_array_equal(gv.cdf(), want[0])
            for a, b in zip(gv.fused(), want[1]):
                np.testing.assert_array_equal(a, b)
            assert gv.row_sum_range() == want[2] and gv.layout() == want[3]
        grp.close()


def test_sharded_upload_rejects_an_invalid_matrix_on_every_device():
    need_two()
    C = sp.canonical_csr(synth.branching_graph(n=500, k=5, n_branches=2, seed=2, workers=1, scale=2.0, noise=0.05).T)
    bad = C.copy()
    bad.indices = bad.indices.copy()
    bad.indices[bad.nnz - 3] = C.shape[0] + 5           # a column out of range, in the second device's slice
    grp = _lib.Group(devices=[0, 1])
    with pytest.raises(_lib.CytopathB200Error, match="column index out of range"):
        grp.create_graph(bad, _lib.CDF_REFERENCE, sharded=True)
    grp.create_graph(C, _lib.CDF_REFERENCE, sharded=True)    # the group is still usable
    grp.close()


MISMATCH_WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
from cytopath_b200 import _lib, synth
from oracle import sampler_port as sp
C = sp.canonical_csr(synth.branching_graph(n=400, k=5, n_branches=2, seed=2, workers=1, scale=2.0, noise=0.05).T)
if dist.get_rank() == 1:
    C = C.copy(); C.data = C.data.copy(); C.data[0] *= 0.5          # one value differs on rank 1
grp = _lib.open_group(device=int(os.environ["LOCAL_RANK"]), dist=True)
try:
    grp.create_graph(C, _lib.CDF_REFERENCE, sharded=True)
    msg = "no error"
except _lib.CytopathB200Error as exc:
    msg = str(exc)
open({out!r} + ".%d.txt" % dist.get_rank(), "w").write(msg)
dist.barrier()
dist.destroy_process_group()
"""


def test_sharded_upload_refuses_ranks_with_different_matrices(tmp_path):
    need_two()
    script = tmp_path / "worker.py"
    out = str(tmp_path / "out")
    script.write_text(MISMATCH_WORKER.format(root=ROOT, out=out))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29543", str(script)],
                       capture_output=True, text=True, env=dict(os.environ, TQDM_DISABLE="1"), timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    for rank in (0, 1):
        assert "passed a different matrix" in open(out + ".%d.txt" % rank).read()


WORKER = r"""
import contextlib, io, os, sys, warnings
import numpy as np
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
import cytopath_b200
from tests.helpers import adata_from_case, load
z = load({case!r})
ad, kw = adata_from_case(z)
np.random.seed(int(z["np_seed"]) + 7919 * dist.get_rank())        # only rank 0 has the golden run's seed
with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
    warnings.simplefilter("ignore")
    cytopath_b200.sampling(ad, **kw, seed=int(z["seed"]))
s = ad.uns["samples"]
np.savez({out!r} + ".%d.npz" % dist.get_rank(), cell_sequences=s["cell_sequences"], transition_score=s["transition_score"],
         cluster_sequences=s["cluster_sequences"], world=ad.uns["run_info"]["b200"]["world_size"])
dist.destroy_process_group()
"""


@pytest.mark.parametrize("case", ["sampling_tiny_dups", "sampling_small_clusters", "sampling_config1_auto"])
def test_two_gpu_torchrun_equals_reference(case, tmp_path):
    need_two()
    script = tmp_path / "worker.py"
    out = str(tmp_path / "out")
    script.write_text(WORKER.format(root=ROOT, case=case, out=out))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)],
                       capture_output=True, text=True, env=dict(os.environ, TQDM_DISABLE="1"), timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    z = load(case)
    for rank in (0, 1):
        got = np.load(out + ".%d.npz" % rank)
        assert int(got["world"]) == 2
        check(got, z)
