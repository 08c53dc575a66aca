"""HBM roofline of K1 (normalise + cast, once per corpus row) and of the exact rescore gather.
Algorithmic bytes of K1 per row: 4*D read + 4*D raw copy + 2*D_pad operand + 8 (norm) + 4 (residual)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cross_modal_video_engine_b200 import engine, synth  # noqa: E402


def main():
    n, reps = 500_000, 10
    for dims in ((2048,), (1536, 512), (640,)):
        d = sum(dims)
        x = synth.device_gaussian(n, d, 3, "cuda")
        store = engine.CorpusStore(n, dims)
        for _ in range(2):
            store.n = 0
            store.add(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            store.n = 0
            store.add(x)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        byts = n * (4 * d + 4 * d + 2 * sum((k + 63) // 64 * 64 for k in dims) + 12 * len(dims))
        print("K1 %d rows x %s: %.3f ms  %.0f GB/s algorithmic (%.2f GB)" % (n, dims, ms, byts / ms / 1e6, byts / 1e9),
              flush=True)
        del store, x
        torch.cuda.empty_cache()
    # frames = 8 (MultiFusion index): mean-pool fused into K1
    n = 100_000
    x = synth.device_gaussian(n * 8, 640, 4, "cuda").view(n, 8, 640)
    store = engine.CorpusStore(n, (640,), norm_mode="eps")
    store.add(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        store.n = 0
        store.add(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    byts = n * (8 * 4 * 640 + 4 * 640 + 2 * 640 + 12)
    print("K1 %d rows x 8 frames x 640: %.3f ms  %.0f GB/s algorithmic" % (n, ms, byts / ms / 1e6), flush=True)


if __name__ == "__main__":
    main()
