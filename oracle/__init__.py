"""CPU oracle for the HOSNeRF per-ray hot path.  TEST INFRASTRUCTURE ONLY.

This package is a PyTorch-CPU (fp32) restatement of the reference algorithms
for the path named in BASELINE.json (sampler -> LBS / non-rigid warp -> positional
encoding -> MLP -> alpha composite).  It is NOT on the product path:

* only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
  ``--impl reference`` legs of ``bench.py`` may import it;
* ``hosnerf_b200`` never imports it and has no CPU fallback - the product path
  raises if ``libhosnerf_b200.so`` is missing.

Pinning.  The reference ships no tests, golden vectors or fixtures for this path
(SURVEY.md section 4), so the oracle is pinned against *outputs of the reference
itself*: ``tests/golden/make_golden.py`` imports the unmodified reference from
``/root/reference`` (authoring container only), runs it on the seeded inputs of
``hosnerf_b200.synth`` and commits the results as ``tests/golden/*.npz``.
``tests/test_oracle_golden.py`` checks every oracle function against those
fixtures (CPU, no GPU needed).

Every function cites the reference file:line it restates.  Tags:
  S1 = 1st_State-Conditional_Scene, S2 = 2nd_State_Conditional_Human-Object,
  S3 = 3rd_Complete_HOSNeRF.

One piece is "parity unpinned by a stock run": the stage-3 composite block
(S3/src/model/mipnerf360/model.py:1524-1596) hard-codes ``.cuda()``; its golden
vector is produced by calling the reference ``LitMipNeRF360.training_step``
unbound on a duck-typed ``self`` with ``Tensor.cuda`` patched to the identity
(see make_golden.py) - the arithmetic executed is still the reference's.
"""
