"""CPU oracle for the retrieval-scoring hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (NumPy / torch-CPU) of the reference's scoring -> ranking ->
metric path.  It is the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product path
(``cross-modal-video-engine_b200/``) never imports it and has no CPU fallback.

Pinning status
--------------
* ``oracle.linas`` restates ``LINAS-engine/{evaluation,validate}.py``, ``util/metrics.py`` and
  ``basic/metric.py``.  The reference ships no golden vectors for this path, but
  the reference modules import and run in the build container, so the restatement is pinned
  against **outputs of the reference itself**: ``oracle/make_golden.py`` imports the reference
  from ``/root/reference/LINAS-engine``, runs it on seeded inputs and commits the results under
  ``tests/golden/``; ``tests/test_oracle_golden.py`` checks the restatement against them.
* ``oracle.multifusion`` restates ``MultiFusion/src/validate.py:44-55,65-113,119,135-138`` and
  ``MultiFusion/src/inference.py:51-66``.  Those modules import pip packages that are absent here
  (``clip``, ``decord``, ``h5py``, ``comet_ml``, ``ftfy``) but never touch them on the scoring path, so
  ``oracle/make_golden_mf.py`` puts empty stand-ins into ``sys.modules``, imports the two reference modules
  UNMODIFIED, replaces the one upstream call (``generate_cirr_val_predictions``: CLIP + Combiner) by seeded
  predictions and runs ``validate.compute_cirr_val_metrics`` / ``inference.compute_cirr_val_metrics`` whole.
  The 7-tuples, the ``results_wo_attn.npy`` top-100 name lists and the top-1 names are committed under
  ``tests/golden/mf_cirr*``; ``tests/test_oracle_golden.py`` pins the restatement to them (bit-identical).
"""
