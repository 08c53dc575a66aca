"""Seeded synthetic inputs for the five BASELINE.json configs.

Host generators use ``np.random.default_rng`` (PCG64: reproducible for a given NumPy), device
generators use a seeded ``torch.Generator`` on the target device.  Embeddings are *raw*
(un-normalised) so the normalise stage is exercised, and are fp32-representable, like the
fp32 model outputs the reference stores into its float64 arrays (evaluation.py:102-105).
"""
from __future__ import annotations

import numpy as np


def msrvtt_like(seed, n_video, caps_per_video, dim, sigma=1.0, dtype=np.float32, ragged=False):
    """Videos ``[Nv, D]`` and captions ``[Nq, D]`` with caption j planted near video j // cpv.

    ``captions = videos[gt] + sigma * noise``; ids follow the MSR-VTT convention the reference's
    ``get_gt`` splits on (``video7#enc#3``, util/metrics.py:111).  ``ragged=True`` drops the captions
    of a few videos (empty ``v2t_gt`` rows) and leaves a different count on others.
    """
    rng = np.random.default_rng(seed)
    videos = rng.standard_normal((n_video, dim)).astype(np.float32)
    owner = np.repeat(np.arange(n_video), caps_per_video)
    if ragged:
        keep = rng.random(owner.shape[0]) > 0.25
        keep[owner % 7 == 3] = False          # some videos have no caption at all
        owner = owner[keep]
    noise = rng.standard_normal((owner.shape[0], dim)).astype(np.float32)
    captions = (videos[owner] + np.float32(sigma) * noise).astype(np.float32)
    # per-row positive scale: raw embeddings are not unit norm
    captions *= rng.uniform(0.5, 2.0, size=(captions.shape[0], 1)).astype(np.float32)
    videos = videos * rng.uniform(0.5, 2.0, size=(n_video, 1)).astype(np.float32)
    video_ids = ["video%d" % i for i in range(n_video)]
    seen = {}
    caption_ids = []
    for o in owner:
        k = seen.get(int(o), 0)
        seen[int(o)] = k + 1
        caption_ids.append("video%d#enc#%d" % (int(o), k))
    return videos.astype(dtype), captions.astype(dtype), video_ids, caption_ids, owner


def gaussian(seed, n, dim, dtype=np.float32):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((n, dim)).astype(np.float32).astype(dtype)


def clustered(seed, n, dim, n_centroid=64, spread=0.35, dtype=np.float32):
    """Mixture of centroids + noise: larger, structured cosines than i.i.d. Gaussians."""
    rng = np.random.default_rng(seed)
    cent = rng.standard_normal((n_centroid, dim)).astype(np.float32)
    pick = rng.integers(0, n_centroid, size=n)
    x = cent[pick] + np.float32(spread) * rng.standard_normal((n, dim)).astype(np.float32)
    return x.astype(np.float32).astype(dtype)


def composed_retrieval(seed, n_index, n_query, dim=640, frames=8, sigma=0.7):
    """MultiFusion shape: index ``[Nv, frames, D]`` fp32, unit-norm queries planted near a target,
    integer names, a reference item per query that must be dropped (validate.py:76-83)."""
    rng = np.random.default_rng(seed)
    index = rng.standard_normal((n_index, frames, dim)).astype(np.float32)
    pooled = index.mean(axis=1)
    target = rng.integers(0, n_index, size=n_query)
    reference = (target + 1 + rng.integers(0, n_index - 1, size=n_query)) % n_index
    q = pooled[target] + np.float32(sigma) * rng.standard_normal((n_query, dim)).astype(np.float32) \
        * np.float32(np.sqrt(1.0 / frames))
    q += np.float32(0.5) * pooled[reference]          # the reference item scores high too
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    names = rng.permutation(10 * n_index)[:n_index].astype(np.int64)   # ids are ints (utils.py:57)
    return index, q.astype(np.float32), names, names[reference], names[target]


def device_gaussian(n, dim, seed, device, out=None):
    """``[n, dim]`` fp32 standard normal generated on ``device`` (corpora too big for host RAM)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    if out is None:
        return torch.randn((n, dim), generator=g, device=device, dtype=torch.float32)
    return out.normal_(generator=g)
