"""Composed (text + reference video) retrieval on the corpus-resident engine -- the scoring path of
``MultiFusion/src``.  Modules mirror the reference's:

* :mod:`.scoring`   the arithmetic on already-built query features (``cirr_metrics_from_features``, ``top1_name``,
  ``build_index``) -- validate.py:44-55,65-141 and inference.py:51,63-65;
* :mod:`.validate`  ``compute_cirr_val_metrics`` / ``generate_cirr_val_predictions`` / ``cirr_val_retrieval`` with the
  signatures of ``MultiFusion/src/validate.py:27-29,167-169,275``;
* :mod:`.inference` ``compute_cirr_val_metrics`` with the signature of ``MultiFusion/src/inference.py:26-27`` (single
  composed query -> top-1 name).
"""
from .scoring import CirrEvaluator, build_index, cirr_metrics_from_features, name_rows, top1_name  # noqa: F401
from . import inference, validate  # noqa: F401
