"""Two ranks, two GPUs, NCCL: sampling() under torchrun gives the single-GPU / reference result.
Skipped on boxes with fewer than two CUDA devices (run with `gpurun --gpus 2`)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from cytopath_b200 import _lib
from tests.helpers import load

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

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
np.random.seed(int(z["np_seed"]))
with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
    warnings.simplefilter("ignore")
    cytopath_b200.sampling(ad, **kw, seed=int(z["seed"]))
s = ad.uns["samples"]
np.savez({out!r} + ".%d.npz" % dist.get_rank(), cell_sequences=s["cell_sequences"], transition_score=s["transition_score"],
         cluster_sequences=s["cluster_sequences"], world=ad.uns["run_info"]["b200"]["world_size"])
dist.destroy_process_group()
"""


@pytest.mark.parametrize("case", ["sampling_tiny_dups", "sampling_small_clusters", "sampling_config1_auto"])
def test_two_gpu_nccl_run_equals_reference(case, tmp_path):
    if _lib.device_count() < 2:
        pytest.skip("needs two CUDA devices")
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
        np.testing.assert_array_equal(got["cell_sequences"], z["cell_sequences"])
        np.testing.assert_array_equal(got["cluster_sequences"], z["cluster_sequences"])
        np.testing.assert_allclose(got["transition_score"], z["transition_score"], rtol=2e-6)
