"""Stage timing of the C2 evaluation (59 800 captions x 2 990 videos x 1536, fp64 like the reference's arrays):
cal_error (exact fp64 score matrix), the rank kernels of both directions, the whole cal_perf."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cross_modal_video_engine_b200 import evaluation, metrics, synth, validate  # noqa: E402


def timed(fn, reps=5):
    """Median wall time of ``fn`` (host + device, synchronised) over ``reps`` calls after two warm-up calls -- the
    first calls pay cudaMalloc for the 0.7-1.4 GB outputs, and a mean over three calls moved by 3x between boxes."""
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return sorted(ts)[len(ts) // 2], out


def main():
    V, Q, vid, cap, _ = synth.msrvtt_like(1, 2990, 20, 1536, 14.0)
    Vd, Qd = torch.from_numpy(V).double().cuda(), torch.from_numpy(Q).double().cuda()
    ms, e = timed(lambda: evaluation.cal_error(Vd, Qd))
    print("cal_error fp64 59800x2990x1536: %.2f ms (%.1f TFLOP/s fp64)" % (ms, 2 * 59800 * 2990 * 1536 / ms / 1e9))
    ms32, _ = timed(lambda: evaluation.cal_error(Vd.float(), Qd.float()))
    print("cal_error fp32 (tcgen05, split-bf16 x3): %.2f ms" % ms32)
    t0 = time.perf_counter()
    v2t_gt, t2v_gt = metrics.get_gt(vid, cap)
    print("get_gt (host): %.1f ms" % ((time.perf_counter() - t0) * 1e3))
    ms, _ = timed(lambda: metrics.RankResult(e, t2v_gt))
    print("t2v ranks + reduce (rows, %d GT entries): %.2f ms -> %.0f GB/s over the 1.43 GB matrix" % (len(Q), ms, 1.43e3 / ms * 1e0))
    ms, _ = timed(lambda: metrics.RankResult(e.t(), v2t_gt))
    print("v2t ranks + reduce (columns, 20 GT per video): %.2f ms -> %.0f GB/s" % (ms, 1.43e3 / ms))
    ms, perf = timed(lambda: validate.cal_perf(e, v2t_gt, t2v_gt))
    print("cal_perf total: %.2f ms  %s" % (ms, perf[1][:3]))


if __name__ == "__main__":
    main()
