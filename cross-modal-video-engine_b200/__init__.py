"""B200-native retrieval scoring engine: a drop-in for the scoring -> fusion -> ranking -> metrics path of
``LINAS-engine`` (evaluation.py / validate.py / util/metrics.py / basic/metric.py / inference.py) and of
``MultiFusion`` (src/validate.py / src/inference.py), running on hand-written sm_100a CUDA kernels behind
the C ABI in ``include/xmve.h``.  There is no CPU or PyTorch fallback.

Modules mirror the reference's: ``evaluation`` (l2norm, cal_error, cal_simi), ``validate`` (norm_score,
cal_perf), ``metrics`` (get_gt, eval_q2m, t2v_map, v2t_map, ...), ``basic_metric`` (getScorer),
``multifusion`` (compute_cirr_val_metrics), plus ``engine`` (CorpusStore.search) and ``distributed``.
"""
__version__ = "0.1.0"
