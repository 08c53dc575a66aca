"""Drop-in for ``MultiFusion/src/inference.py:26-66``: one composed query (reference-video features + modification
text) against a pre-pooled index -> the name of the top-1 item.

The text tower / Combiner calls are the caller's modules; the ``F.normalize(index)`` (:51), ``1 - q @ index.T`` (:63)
and ``torch.argsort(scores.cpu())[0][0]`` (:64-65) tail runs on the resident store (top-1 search).
"""
from __future__ import annotations

import torch

from . import scoring
from .validate import _tokenizer


def compute_cirr_val_metrics(ref_vdo_feature, mod_text, clip_model, index_features, index_names, combining_function,
                             combiner, *, store=None, tokenize=None):
    """inference.py:26-66.  ``ref_vdo_feature = (high [F, D], middle [F, 18*18, C])``; ``index_features`` is already
    frame-pooled ``[N, D]`` (inference.py:133).  Pass ``store`` (``scoring.build_index(index_features)``) to keep the
    index resident across queries instead of re-normalising it per call as :51 does."""
    tokenize = _tokenizer(tokenize)
    device = index_features.device
    ref_vdo_feature_high, ref_vdo_feature_middle = ref_vdo_feature
    ref_vdo_feature_high = ref_vdo_feature_high.unsqueeze(0)
    text_inputs = tokenize(mod_text).to(device, non_blocking=True)
    middle_feature = ref_vdo_feature_middle.to(device, non_blocking=True).float()
    middle_feature = torch.nn.functional.adaptive_avg_pool2d(
        middle_feature.reshape(1, middle_feature.shape[0], 18 * 18, -1), (16, index_features.shape[-1]))
    with torch.no_grad():
        text_features = clip_model.encode_text(text_inputs)
        batch_predicted_features = combining_function((ref_vdo_feature_high, middle_feature), text_features)
    return scoring.top1_name(batch_predicted_features[0], index_features, index_names, store=store)
