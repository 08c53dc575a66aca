"""Corpus ingest: the reference's on-disk formats -> the HBM-resident ``CorpusStore``.

* **BigFile** (``LINAS-engine/basic/bigfile.py:4-56``; written by ``util/txt2bin.py:21-75``): a directory with
  ``shape.txt`` (``"<n> <dim>"``), ``id.txt`` (whitespace-separated names) and ``feature.bin`` (``n * dim`` float32,
  row-major).  :class:`BigFile` keeps the reference's ``read`` / ``read_one`` / ``shape`` behaviour (results sorted by
  row index, duplicates and unknown names dropped, rows returned as Python lists) and adds ``rows()`` -- a zero-copy
  ``np.memmap`` view -- and ``to_store()``, which streams the file into a store through pinned staging buffers.
* **video_data.pt** (``LINAS-engine/inference.py:57-67``): ``{'video_embs': ndarray [N, D], 'video_ids': list}``,
  the cache ``encode_vid`` output is saved to.  ``load_video_data`` / ``save_video_data``.
* The MultiFusion index ``[N, frames, D]`` goes through ``multifusion.build_index`` (frame mean fused into K1).

Reading and writing the formats is host code; only ``to_store`` touches the GPU.
"""
from __future__ import annotations

import os

import numpy as np


class BigFile:
    """Reader for the reference's BigFile feature directories (basic/bigfile.py:4-56)."""

    def __init__(self, datadir):
        with open(os.path.join(datadir, "shape.txt")) as f:
            self.nr_of_images, self.ndims = map(int, f.readline().split())
        with open(os.path.join(datadir, "id.txt"), "rb") as f:
            self.names = [str(x, encoding="ISO-8859-1") for x in f.read().strip().split()]
        assert len(self.names) == self.nr_of_images
        self.name2index = dict(zip(self.names, range(self.nr_of_images)))
        self.binary_file = os.path.join(datadir, "feature.bin")
        self._mm = None

    def rows(self):
        """All features as a read-only ``np.memmap`` ``[n, dim]`` float32 (no copy)."""
        if self._mm is None:
            self._mm = np.memmap(self.binary_file, dtype=np.float32, mode="r", shape=(self.nr_of_images, self.ndims))
        return self._mm

    def read(self, requested, isname=True):
        """``(names, vectors)`` of the requested items, de-duplicated and in ascending row order, vectors as Python
        lists of the float32 values (what ``array('f').tolist()`` yields); unknown names are skipped, out-of-range
        row numbers are an error -- the observable behaviour of bigfile.py:22-52."""
        wanted = set(requested)
        if isname:
            picked = sorted(self.name2index[name] for name in wanted if name in self.name2index)
        else:
            picked = sorted(wanted)
            if picked:
                assert picked[0] >= 0 and picked[-1] < self.nr_of_images
        if not picked:
            return [], []
        block = np.asarray(self.rows()[np.asarray(picked, dtype=np.int64)], dtype=np.float64)   # one gather, widened
        return [self.names[r] for r in picked], block.tolist()

    def read_one(self, name):
        _, vectors = self.read([name])
        return vectors[0]

    def shape(self):
        return [self.nr_of_images, self.ndims]

    def to_store(self, store=None, chunk_rows=1 << 18, device="cuda", dims=None):
        """Stream ``feature.bin`` into a :class:`~.engine.CorpusStore` (created if ``store`` is None): chunks are
        staged in two pinned host buffers so that the disk read of chunk i+1 overlaps the H2D copy + K1 of chunk i.
        Returns ``(store, names)``; row r of the store is ``names[r]``."""
        import torch
        from .engine import CorpusStore
        if store is None:
            store = CorpusStore(self.nr_of_images, dims or (self.ndims,), device=device)
        mm = self.rows()
        stage = [torch.empty((min(chunk_rows, max(self.nr_of_images, 1)), self.ndims), dtype=torch.float32).pin_memory()
                 for _ in range(2)]
        done = [None, None]
        for c, lo in enumerate(range(0, self.nr_of_images, chunk_rows)):
            hi = min(self.nr_of_images, lo + chunk_rows)
            b = c & 1
            if done[b] is not None:
                done[b].synchronize()                       # the copy out of this staging buffer has finished
            np.copyto(stage[b].numpy()[: hi - lo], mm[lo:hi])       # disk (page cache) -> pinned staging buffer
            store.add(stage[b][: hi - lo])
            done[b] = torch.cuda.Event()
            done[b].record()
        return store, list(self.names)


def write_bigfile(datadir, names, features):
    """Write ``shape.txt`` / ``id.txt`` / ``feature.bin`` as ``util/txt2bin.py:21-75`` does (first occurrence of a
    name wins, rows containing NaN are dropped); returns the number of rows written."""
    features = np.asarray(features, dtype=np.float32)
    os.makedirs(datadir, exist_ok=True)
    seen, kept = set(), []
    with open(os.path.join(datadir, "feature.bin"), "wb") as fw:
        for name, vec in zip(names, features):
            if name in seen:
                continue
            seen.add(name)
            if np.isnan(vec).any():
                continue
            vec.tofile(fw)
            kept.append(name)
    with open(os.path.join(datadir, "id.txt"), "w") as fw:
        fw.write(" ".join(kept))
    with open(os.path.join(datadir, "shape.txt"), "w") as fw:
        fw.write("%d %d" % (len(kept), features.shape[1] if features.ndim == 2 else 0))
    return len(kept)


def save_video_data(path, video_embs, video_ids):
    """``torch.save({'video_embs': ..., 'video_ids': ...}, path)``; inference.py:67."""
    import torch
    torch.save({"video_embs": np.asarray(video_embs), "video_ids": list(video_ids)}, path)


def load_video_data(path, store=None, device="cuda", dims=None, chunk_rows=1 << 18):
    """Load the ``video_data.pt`` cache (inference.py:57-60) into a store.  Returns ``(store, video_ids)``."""
    import torch
    from .engine import CorpusStore
    data = torch.load(path, weights_only=False)
    embs, ids = data["video_embs"], list(data["video_ids"])
    embs = embs if torch.is_tensor(embs) else torch.from_numpy(np.ascontiguousarray(embs))
    if store is None:
        store = CorpusStore(embs.shape[0], dims or (embs.shape[1],), device=device)
    for lo in range(0, embs.shape[0], chunk_rows):
        store.add(embs[lo:lo + chunk_rows])
    return store, ids
