"""CPU oracle for the RL_CTR_Prediction hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product: it may be imported only by
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` -- and there only as the checker or as the
timed CPU baseline, never as a fallback for the CUDA path.

Two layers:

* :mod:`oracle.np_oracle`  -- numpy restatement (fp32 to mirror the reference,
  fp64 as arbiter) of every function on the path, each citing the reference
  file:line it follows.
* :mod:`oracle.torch_port` -- a restatement on stock **CPU** PyTorch ops of the
  reference's modules and loop bodies.  It executes the same ATen kernels as the
  reference (``index_select``, ``embedding_dense_backward``, dense Adam) with
  all host threads, so it is what ``bench.py`` times as ``cpu_baseline`` with
  ``kind: "port"`` (the reference itself is pure Python and cannot travel to
  the GPU box).

Pinning status: the reference ships no tests, golden vectors or fixtures
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference's own
modules imported from ``/root/reference`` in the build container; the vectors
and the script that made them are committed in ``tests/golden/``.
"""
