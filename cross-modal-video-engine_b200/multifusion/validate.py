"""Drop-in for the evaluation functions of ``MultiFusion/src/validate.py`` (same names, arguments and return values;
callers ``validate.py:292`` and ``combiner_train.py:398``).

The CLIP text tower and the Combiner are the caller's torch modules (upstream of the scoring path); what changes is
everything after them: the index is mean-pooled, normalised and kept as a resident :class:`CorpusStore`, and the
32-query ``1 - P @ index.T`` + full ``torch.argsort`` blocks (validate.py:65-113) become one top-100 search.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import scoring


def _tokenizer(tokenize):
    if tokenize is not None:
        return tokenize
    try:
        import clip                                   # openai-clip, the reference's dependency (validate.py:10)
    except ImportError as exc:                        # pragma: no cover
        raise ImportError("generate_cirr_val_predictions needs the `clip` package for clip.tokenize "
                          "(or pass tokenize=...)") from exc
    return clip.tokenize


def _batches(dataset, batch_size):
    """``DataLoader(dataset, batch_size=32, collate_fn=utils.collate_fn)`` (validate.py:207-208, utils.py:96-103):
    ``None`` items are discarded, the rest is default-collated."""
    from torch.utils.data import DataLoader
    from torch.utils.data.dataloader import default_collate
    return DataLoader(dataset=dataset, batch_size=batch_size, num_workers=0,
                      collate_fn=lambda b: default_collate([x for x in b if x is not None]))


def generate_cirr_val_predictions(clip_model, relative_val_dataset, combining_function, index_names, index_features,
                                  tokenize=None, batch_size=32):
    """validate.py:167-272.  Items of ``relative_val_dataset`` are ``(reference_name, target_name, caption,
    group_members, middle_feature)`` (data_utils.py:215).  Returns ``(predicted_features [Nq, D] unit-norm, on the
    device of ``index_features``; reference_names; target_names)``.

    The reference looks every reference item up in a ``name -> feature`` dict and ``torch.stack``s them
    (validate.py:211,238-241); here the rows are gathered on the device with one ``index_select``.
    """
    tokenize = _tokenizer(tokenize)
    clip_model.eval()
    device = index_features.device
    predicted, target_names, reference_names = [], [], []
    for batch_reference_names, batch_target_names, captions, _members, middle_feature in \
            _batches(relative_val_dataset, batch_size):
        text_inputs = tokenize(captions).to(device, non_blocking=True)
        middle_feature = middle_feature.to(device, non_blocking=True).float()
        with torch.no_grad():
            text_features = clip_model.encode_text(text_inputs)
            rows = torch.from_numpy(scoring.name_rows(index_names, np.asarray(batch_reference_names))).to(device)
            if bool((rows < 0).any()):
                raise KeyError("a reference item is not in index_names")     # itemgetter raises KeyError too (:240)
            reference_image_features = index_features.index_select(0, rows)
            batch_predicted = combining_function((reference_image_features, middle_feature), text_features)
        predicted.append(F.normalize(batch_predicted, dim=-1))
        target_names.extend(np.asarray(batch_target_names).tolist())
        reference_names.extend(np.asarray(batch_reference_names).tolist())
    if predicted:
        predicted_features = torch.vstack(predicted)
    else:
        predicted_features = torch.empty((0, index_features.shape[-1]), device=device)
    return predicted_features, reference_names, target_names


def compute_cirr_val_metrics(relative_val_dataset, clip_model, index_features, index_names, combining_function,
                             combiner, *, store=None, results_path="results_wo_attn", tokenize=None):
    """validate.py:27-143: ``(group_recall@1, @2, @3, recall@1, @5, @10, @50)``; the group recalls are the
    reference's constant -1 (:139-141).  Writes ``sorted_index_names[:, :100]`` to ``results_path + ".npy"`` like
    validate.py:119 (``results_path=None`` skips the file).

    ``combiner.time_process`` of the reference is the frame mean (combiner.py:140-143); it is fused into the row
    normalisation kernel when the store is built, so ``combiner`` is only used through ``combining_function``.
    ``store`` lets a caller that evaluates every epoch (combiner_train.py:398) keep the index resident.
    """
    predicted_features, reference_names, target_names = generate_cirr_val_predictions(
        clip_model, relative_val_dataset, combining_function, index_names, index_features, tokenize=tokenize)
    metrics, top = scoring.cirr_metrics_from_features(predicted_features, index_features, index_names,
                                                      reference_names, target_names, top_names=100, store=store)
    if results_path is not None:
        np.save(results_path, top)
    return metrics


def cirr_val_retrieval(combining_function, clip_model, preprocess, args, combiner, *, datasets=None, **kw):
    """validate.py:275-293.  ``datasets = (classic_val_dataset, relative_val_dataset)``; when omitted they are built
    with the reference's ``ComposedVideoDataset`` / ``extract_index_features`` (its ``data_utils`` / ``utils`` modules
    must be importable -- dataset readers are outside this package)."""
    clip_model = clip_model.float().eval()
    if datasets is None:                              # pragma: no cover  (needs the reference's dataset files)
        from data_utils import ComposedVideoDataset
        datasets = (ComposedVideoDataset('test', 'classic', preprocess, args.data_pth, args.dataset_op),
                    ComposedVideoDataset('test', 'relative', preprocess, args.data_pth, args.dataset_op))
    classic_val_dataset, relative_val_dataset = datasets
    index_features, index_names = extract_index_features(classic_val_dataset, clip_model)
    return compute_cirr_val_metrics(relative_val_dataset, clip_model, index_features, index_names,
                                    combining_function, combiner, **kw)


def extract_index_features(dataset, clip_model, device="cuda", batch_size=32):
    """utils.py:32-58: stack the ``(name, [T, D] frame features)`` items of a 'classic' dataset into
    ``index_features [N, T, D]`` on the device and the list of names.  (The reference grows the tensor with one
    ``torch.vstack`` per batch, O(N^2) bytes; here the batches are concatenated once.)"""
    feats, names = [], []
    for batch_names, vdo_fea in _batches(dataset, batch_size):
        feats.append(vdo_fea.to(device, non_blocking=True))
        names.extend(np.asarray(batch_names).tolist())
    d = clip_model.visual.output_dim
    index_features = torch.cat(feats, dim=0) if feats else torch.empty((0, 8, d), device=device)
    return index_features, names
