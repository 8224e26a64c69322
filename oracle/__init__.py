"""CPU oracle for the cytopath Markov-chain sampler hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under `oracle/` is part of the product.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs may
import it, and only as the checker or the reported CPU baseline.  `cytopath_b200`
never imports `oracle` and has no CPU fallback: without the CUDA library it raises.

What it restates (reference = /root/reference/cytopath/markov_chain_sampler.py):

  philox.py        the shared uniform stream (ours; Philox4x32-10, Random123 KATs)
  sampler_port.py  matrix_prep :11-31, markov_sim :35-52, check_convergence_criteria
                   :54-60, iterate_state_probability :62-87, sampling :89-414
                   (validation, round loop, filters, final draw) on CSR, no densify
  walk.c           the same CDF recipe + first-crossing step loop + Philox in plain C
                   (gcc), used to check large cases in seconds and as a compiled CPU
                   baseline next to the faithful Python/joblib port
  fused_rows_port.py  the device layout K1b derives from that CDF (12-bit codes, inlined
                   targets, row words; cytopath_b200/csrc/rows_v2.cu) rebuilt in numpy, so that the
                   layout is checked byte for byte and not only through the walks it produces
  hausdorff_port.py  the directed-Hausdorff distance matrix of sample_clustering.py:57-71

Parity pinning: the reference ships no tests or golden vectors,
so the oracle is pinned against the reference ITSELF: `oracle/make_golden.py` imports
the reference module from /root/reference in the build container, drives its
`np.random.rand` with the Philox stream, and commits inputs+outputs under
`tests/golden/`.  `tests/test_oracle_golden.py` checks this restatement against those
vectors on every run (no /root/reference needed at test time).
"""
