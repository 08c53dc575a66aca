"""Drop-in for ``LINAS-engine/validate.py``: ``norm_score`` (:7-11), ``cal_perf`` (:15-54) and the per-epoch
``validate`` driver (:58-90, called from trainer.py:259,276)."""
from __future__ import annotations

import logging

import numpy as np
import torch

from . import _native as N
from . import evaluation, metrics


def norm_score(t2v_all_errors):
    """Global min-max normalisation of the score matrix, same operation order and dtype as validate.py:7-11."""
    was_numpy = not torch.is_tensor(t2v_all_errors)
    N.require_device()
    dev = torch.device("cuda", torch.cuda.current_device())
    x = torch.from_numpy(np.ascontiguousarray(t2v_all_errors)) if was_numpy else t2v_all_errors
    if x.dtype not in (torch.float32, torch.float64):
        x = x.to(torch.float64)
    x = x.to(dev).contiguous()
    out = torch.empty_like(x)
    scratch = torch.empty(2, dtype=torch.float64, device=dev)
    N.call("xmve_norm_score", N.ptr(x), N.F64 if x.dtype == torch.float64 else N.F32, x.shape[0], x.shape[1],
           x.stride(0), N.ptr(out), out.stride(0), N.ptr(scratch), N.stream_ptr())
    return out.cpu().numpy() if was_numpy else out


def cal_perf(t2v_all_errors, v2t_gt, t2v_gt, tb_logger=None, model=None):
    """Same return tuple, log lines and tensorboard keys as validate.py:15-54.

    The errors matrix is moved to the device once; text->video ranks are counted along its rows and
    video->text ranks along its columns (no transpose, no sort).
    """
    N.require_device()
    dev = torch.device("cuda", torch.cuda.current_device())
    x = t2v_all_errors if torch.is_tensor(t2v_all_errors) else torch.from_numpy(t2v_all_errors)
    x = x.to(dev, non_blocking=True)

    # everything of both directions is enqueued before the first device -> host read: the host builds the
    # ground-truth CSR of the second direction while the GPU ranks the first
    t2v = metrics.RankResult(x, t2v_gt)                        # video retrieval: rows
    t2v_ap = t2v.ap_vector(first_only=True)                    # t2v_map marks only the first ground truth (:72-73)
    v2t = metrics.RankResult(x.t(), v2t_gt)                    # caption retrieval: columns

    (t2v_r1, t2v_r5, t2v_r10, t2v_medr, t2v_meanr) = t2v.recall_medr_meanr()
    t2v_map_score = np.mean(t2v_ap.cpu().numpy())
    (v2t_r1, v2t_r5, v2t_r10, v2t_medr, v2t_meanr) = v2t.recall_medr_meanr()
    v2t_map_score = v2t.mean_ap()

    logging.info(" * Text to Video:")
    logging.info(" * r_1_5_10, medr, meanr: {}".format([round(t2v_r1, 1), round(t2v_r5, 1), round(t2v_r10, 1), round(t2v_medr, 1), round(t2v_meanr, 1)]))
    logging.info(" * recall sum: {}".format(round(t2v_r1+t2v_r5+t2v_r10, 1)))
    logging.info(" * mAP: {}".format(round(t2v_map_score, 4)))
    logging.info(" * "+'-'*10)

    logging.info(" * Video to text:")
    logging.info(" * r_1_5_10, medr, meanr: {}".format([round(v2t_r1, 1), round(v2t_r5, 1), round(v2t_r10, 1), round(v2t_medr, 1), round(v2t_meanr, 1)]))
    logging.info(" * recall sum: {}".format(round(v2t_r1+v2t_r5+v2t_r10, 1)))
    logging.info(" * mAP: {}".format(round(v2t_map_score, 4)))
    logging.info(" * "+'-'*10)

    if tb_logger is not None:
        for key, val in (('v2t_r1', v2t_r1), ('v2t_r5', v2t_r5), ('v2t_r10', v2t_r10), ('v2t_medr', v2t_medr),
                         ('v2t_meanr', v2t_meanr), ('t2v_r1', t2v_r1), ('t2v_r5', t2v_r5), ('t2v_r10', t2v_r10),
                         ('t2v_medr', t2v_medr), ('t2v_meanr', t2v_meanr), ('v2t_map', v2t_map_score),
                         ('t2v_map', t2v_map_score)):
            tb_logger.log_value(key, val, step=model.Eiters)

    return (v2t_r1, v2t_r5, v2t_r10, v2t_medr, v2t_meanr, v2t_map_score), (t2v_r1, t2v_r5, t2v_r10, t2v_medr, t2v_meanr, t2v_map_score)


def validate(opt, tb_logger, vid_data_loader, text_data_loader, model, measure='cosine'):
    """validate.py:58-90: encode the validation videos and captions, score, rank, and return the model-selection
    score ``currscore`` (recall sum or mAP sum over the directions ``opt.direction`` names); logs ``rsum``.

    Same control flow and the same ``opt`` fields (``style``, ``student_model``, ``val_metric``, ``direction``) as the
    reference -- including its fall-through: a ``style`` other than 'distill_from_best_model' / 'GT' leaves
    ``cap_embs`` unbound and raises ``UnboundLocalError``.  What differs is where the data lives: the embeddings stay
    on the device (``evaluation.encode_*``), the errors matrix is produced and ranked there, and only the twelve
    scalars come back.
    """
    # compute the encoding for all the validation video and captions
    model.val_start()
    if opt.style == 'distill_from_best_model' and opt.student_model == 'text+video':
        video_embs, video_ids = evaluation.encode_vid(model.embed_vis_distill, vid_data_loader)
    else:
        video_embs, video_ids = evaluation.encode_vid(model.embed_vis, vid_data_loader)

    if opt.style == 'distill_from_best_model':
        cap_embs, caption_ids = evaluation.encode_text(model.embed_txt_distill, text_data_loader, opt.style)
    elif opt.style == 'GT':
        cap_embs, caption_ids = evaluation.encode_text(model.embed_txt_GT, text_data_loader, opt.style)

    t2v_all_errors = evaluation.cal_error(video_embs, cap_embs, measure)
    v2t_gt, t2v_gt = metrics.get_gt(video_ids, caption_ids)

    (v2t_r1, v2t_r5, v2t_r10, v2t_medr, v2t_meanr, v2t_map_score), \
        (t2v_r1, t2v_r5, t2v_r10, t2v_medr, t2v_meanr, t2v_map_score) = \
        cal_perf(t2v_all_errors, v2t_gt, t2v_gt, tb_logger=tb_logger, model=model)

    currscore = 0
    if opt.val_metric == "recall":
        if opt.direction == 'i2t' or opt.direction == 'all':
            currscore += (v2t_r1 + v2t_r5 + v2t_r10)
        if opt.direction == 't2i' or opt.direction == 'all':
            currscore += (t2v_r1 + t2v_r5 + t2v_r10)
    elif opt.val_metric == "map":
        if opt.direction == 'i2t' or opt.direction == 'all':
            currscore += v2t_map_score
        if opt.direction == 't2i' or opt.direction == 'all':
            currscore += t2v_map_score

    tb_logger.log_value('rsum', currscore, step=model.Eiters)

    return currscore
