"""Steady-state timing of the smaller BASELINE configs (C3: 60 x 1.08M x 2048 top-1000; C4: 4096 x 1M x 640 top-100)
through the public search API, with the per-phase split.  A tuning tool, not a bench.py line."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cross_modal_video_engine_b200 import engine, synth  # noqa: E402


def build(nv, d, seed, chunk=250000):
    store = engine.CorpusStore(nv, (d,))
    for c, lo in enumerate(range(0, nv, chunk)):
        store.add(synth.device_gaussian(min(chunk, nv - lo), d, seed * 1000 + c, "cuda"))
    return store


def run(tag, nv, nq, d, k, reps=10):
    store = build(nv, d, 7)
    q = synth.device_gaussian(nq, d, 8, "cuda")
    for _ in range(3):
        store.search(q, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        store.search(q, k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    st = {}
    store.search(q, k, stats=st)
    flops = 2.0 * nq * nv * d
    byts = 2.0 * (nv + nq) * d + 8.0 * nq * k
    print("%s: %.3f ms/search  %.0f queries/s  %.1f TFLOP/s  %.0f GB/s(alg)  phases %s  cand/row %.0f eps %.2e" % (
        tag, ms, nq / ms * 1e3, flops / ms / 1e9, byts / ms / 1e6,
        {k_: round(v, 3) for k_, v in st["phases_ms"].items()}, float(st["cand_count"][0].float().mean()), st["eps"]), "rescored/row %.0f" % st["rescored_per_query"],
        flush=True)
    if os.environ.get("XMVE_BENCH_GRAPH", "1") != "0":
        # the same search replayed from a CUDA graph (engine.GraphSearch), resolved one step late
        gs = engine.GraphSearch(store, nq, k)
        prev = None
        for _ in range(3):
            gs(q)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            p = gs(q, defer=True)
            if prev is not None:
                prev.result()
            prev = p
        prev.result()
        e1.record()
        torch.cuda.synchronize()
        ms_g = e0.elapsed_time(e1) / reps
        print("%s: CUDA-graph replay %.3f ms/search  %.0f queries/s" % (tag, ms_g, nq / ms_g * 1e3), flush=True)
        del gs
    del store
    torch.cuda.empty_cache()


if __name__ == "__main__":
    which = sys.argv[1:] or ["c3", "c4"]
    if "c3" in which:
        run("C3 60x1.08Mx2048 k=1000", 1_080_000, 60, 2048, 1000)
    if "c4" in which:
        run("C4 4096x1Mx640 k=100", 1_000_000, 4096, 640, 100)
    if "c5s" in which:
        run("8192x2Mx2048 k=100", 2_000_000, 8192, 2048, 100, reps=3)
    if "c1" in which:
        run("1x1Mx2048 k=10 (online query)", 1_000_000, 1, 2048, 10)
