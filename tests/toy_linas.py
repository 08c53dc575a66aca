"""A toy LINAS model + data loaders that speak the reference's evaluation protocol (evaluation.py:88-171,
validate.py:58-90): loaders yield ``(datas, idxs, data_ids)`` (or ``(datas, support_datas, idxs, data_ids)`` for the
'GT' style) in a shuffled order with a ragged last batch, and carry ``.dataset`` for ``len()``.  The encoders are
single IEEE multiplications per element, so a CPU run (the reference, when goldens are minted) and a CUDA run (the
engine) produce bit-identical fp32 embeddings.  Used by oracle/make_golden_validate.py and the tests."""
import numpy as np
import torch


class Loader:
    def __init__(self, feats, ids, seed, batch=7, support=None, device="cpu"):
        self.dataset = list(ids)
        self.feats, self.ids, self.support, self.batch, self.device = feats, list(ids), support, batch, device
        self.order = np.random.default_rng(seed).permutation(len(ids))

    def __iter__(self):
        for lo in range(0, len(self.order), self.batch):
            idxs = [int(i) for i in self.order[lo:lo + self.batch]]
            datas = torch.from_numpy(self.feats[idxs]).to(self.device)
            data_ids = [self.ids[i] for i in idxs]
            if self.support is None:
                yield datas, idxs, data_ids
            else:
                yield datas, torch.from_numpy(self.support[idxs]).to(self.device), idxs, data_ids


class Model:
    """embed_* as model.py:707-800 exposes them; each scales the input by a fixed per-dimension vector."""

    def __init__(self, dim, device="cpu"):
        g = np.random.default_rng(99)
        mk = lambda: torch.from_numpy(g.uniform(0.5, 1.5, dim).astype(np.float32)).to(device)   # noqa: E731
        self.w_vis, self.w_vis_d, self.w_txt_d, self.w_txt_gt, self.w_sup = mk(), mk(), mk(), mk(), mk()
        self.Eiters = 123
        self.started = 0

    def val_start(self):
        self.started += 1

    def embed_vis(self, x):
        return x * self.w_vis

    def embed_vis_distill(self, x):
        return x * self.w_vis_d

    def embed_txt_distill(self, x):
        return x * self.w_txt_d

    def embed_txt_GT(self, x, support):
        return x * self.w_txt_gt + support * self.w_sup


class TbLogger:
    def __init__(self):
        self.rows = []

    def log_value(self, key, val, step=None):
        self.rows.append([key, float(val), int(step)])


class Opt:
    def __init__(self, style, student_model, val_metric, direction):
        self.style, self.student_model, self.val_metric, self.direction = style, student_model, val_metric, direction


#: (style, student_model, val_metric, direction)
OPTS = [("distill_from_best_model", "text+video", "recall", "all"),
        ("distill_from_best_model", "text", "recall", "t2i"),
        ("distill_from_best_model", "text+video", "map", "all"),
        ("GT", "text+video", "map", "i2t"),
        ("GT", "text", "recall", "i2t"),
        ("GT", "text", "other", "all")]


def inputs():
    from cross_modal_video_engine_b200 import synth
    V, Q, vid_ids, cap_ids, _ = synth.msrvtt_like(81, 60, 5, 48, 2.5, ragged=True)
    S = synth.gaussian(82, len(Q), 48) * np.float32(0.05)
    return V, Q, S, vid_ids, cap_ids


def loaders(style, device="cpu"):
    V, Q, S, vid_ids, cap_ids = inputs()
    vid = Loader(V, vid_ids, 1, batch=7, device=device)
    txt = Loader(Q, cap_ids, 2, batch=11, support=S if style == "GT" else None, device=device)
    return vid, txt


# ---- a seeded batch for the triplet ranking loss (oracle/make_golden_loss.py, tests/test_loss.py) ------------------
#: every `measure` loss.TripletLoss knows (:96-109); False selects cosine_sim
LOSS_MEASURES = [False, 'order', 'euclidean', 'jaccard', 'l1', 'l2', 'l1_norm', 'l2_norm']


def loss_batch(n=37, dim=48):
    """(s, im): l2-normalised caption / video embeddings of one training batch, caption i paired with video i."""
    g = torch.Generator().manual_seed(7)
    im = torch.randn((n, dim), generator=g)
    s = im + 2.0 * torch.randn((n, dim), generator=g)          # noisy enough for many margin violations
    im = im / im.norm(dim=1, keepdim=True)
    s = s / s.norm(dim=1, keepdim=True)
    return s, im
