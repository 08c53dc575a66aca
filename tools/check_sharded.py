"""Multi-GPU correctness check (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_sharded.py

Every rank holds one shard of a seeded corpus; ``distributed.sharded_search`` over NCCL must return, on every rank,
exactly what ONE store holding the whole corpus returns (indices identical, fp64 scores identical), with and
without exclusions, for the filtered path and the small-corpus path, plus ``upload_rows``, the AVS AP and the exact
ground-truth ranks of ``engine.rank_of_gt``.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cross_modal_video_engine_b200 import avs, distributed, engine, metrics, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    dims, w = (256, 64), (0.7, 0.3)
    ok = True
    for nv, nq, k in ((600_000, 300, 100), (9_000, 64, 10), (300_001, 7, 1000)):
        V = synth.device_gaussian(nv, sum(dims), 11, dev)             # same seed on every rank -> same corpus
        Q = synth.device_gaussian(nq, sum(dims), 12, dev)
        lo, hi = distributed.shard_range(nv, world, rank)
        shard = engine.CorpusStore(hi - lo, dims, device=dev, index_offset=lo).add(V[lo:hi])
        full = engine.CorpusStore(nv, dims, device=dev).add(V)
        excl = torch.randint(0, nv, (nq,), generator=torch.Generator().manual_seed(5)).numpy()
        for ex in (None, excl):
            s_ref, i_ref = full.search(Q, k, weights=w, exclude=ex)
            s, i = distributed.sharded_search(shard, Q, k, weights=w, exclude=ex, n_total=nv)
            same = bool(torch.equal(i, i_ref)) and bool(torch.equal(s, s_ref))
            ok = ok and same
            if rank == 0:
                print("nv=%d nq=%d k=%d exclude=%s: %s" % (nv, nq, k, ex is not None, "identical" if same else "MISMATCH"),
                      flush=True)
        if k == 100:
            # several batches in flight with the head (K1 / sampling / threshold, one all-gather) on its own stream
            head = torch.cuda.Stream(device=dev)
            qs = [synth.device_gaussian(nq, sum(dims), 40 + b, dev) for b in range(4)]
            refs = [full.search(q, k, weights=w) for q in qs]
            pend = [distributed.sharded_search(shard, q, k, weights=w, n_total=nv, defer=True, head_stream=head)
                    for q in qs]
            same = all(bool(torch.equal(p.result()[1], r[1])) and bool(torch.equal(p.result()[0], r[0]))
                       for p, r in zip(pend, refs))
            ok = ok and same
            if rank == 0:
                print("head stream, 4 batches in flight nv=%d: %s" % (nv, "identical" if same else "MISMATCH"), flush=True)
            # exact ground-truth ranks without the matrix: owner-shard scores (all-reduce max) + per-shard guard-band
            # counts (all-reduce sum) == the same on one store; R@K / MedR / MeanR / mAP identical
            gts = [[int(x) for x in torch.randint(0, nv, (1 + q % 3,), generator=torch.Generator().manual_seed(q))]
                   for q in range(nq)]
            comm = distributed.GroupComm()
            a = metrics.RankResult.from_store(shard, Q, gts, weights=w, comm=comm, n_total=nv)
            b = metrics.RankResult.from_store(full, Q, gts, weights=w)
            same = bool(torch.equal(a.ranks, b.ranks)) and a.recall_medr_meanr() == b.recall_medr_meanr() \
                and a.mean_ap() == b.mean_ap()
            ok = ok and same
            if rank == 0:
                print("rank_of_gt nv=%d: %s (median rank %.1f)" % (nv, "identical" if same else "MISMATCH",
                                                                   b.recall_medr_meanr()[3]), flush=True)
        if k == 1000:
            # the same search captured once into a CUDA graph WITH its NCCL gathers and replayed on every rank
            gs = engine.GraphSearch(shard, nq, k, weights=w, comm=distributed.GroupComm(), n_total=nv)
            same = True
            for b in range(3):
                qb = synth.device_gaussian(nq, sum(dims), 60 + b, dev)
                s_ref, i_ref = full.search(qb, k, weights=w)
                s_g, i_g = gs(qb)
                same = same and bool(torch.equal(i_g, i_ref)) and bool(torch.equal(s_g, s_ref))
            ok = ok and same
            del gs
            if rank == 0:
                print("CUDA-graph replay over NCCL nv=%d k=%d: %s" % (nv, k, "identical" if same else "MISMATCH"), flush=True)
            rel = [sorted(set(np.random.default_rng(q).integers(0, nv, 300).tolist())) for q in range(nq)]
            s, i = avs.search_avs(shard, Q, k, weights=w, comm=distributed.GroupComm(), n_total=nv)
            ap, m = avs.ap_at_k(i, rel, nv, k)
            ap_ref, m_ref = avs.ap_at_k(full.search(Q, k, weights=w)[1], rel, nv, k)
            ok = ok and bool((ap == ap_ref).all()) and m == m_ref
        del shard, full, V
        torch.cuda.empty_cache()
    # the re-run loop in lockstep on every rank: (a) near-duplicates on the sampling grid -> first threshold too high,
    # (b) dense neighbourhoods -> candidate lists overflow and must grow
    nv, nq, d, k = 200_000, 64, 128, 50
    g = torch.Generator(device=dev).manual_seed(21)
    for case in ("grid", "dense"):
        V = torch.randn((nv, d), generator=g, device=dev)
        Q = torch.randn((nq, d), generator=g, device=dev)
        if case == "grid":
            step = engine.plan(k, nv)["step"]
            grid = torch.arange(0, nv, step, device=dev)
            grid = grid[torch.randperm(grid.numel(), generator=g, device=dev)][: 30 * nq].reshape(nq, 30)
            for qi in range(nq):
                V[grid[qi]] = Q[qi] + 0.3 * torch.randn((30, d), generator=g, device=dev)
        else:
            for qi in range(4):
                rows = torch.randperm(nv, generator=g, device=dev)[:30000]
                V[rows] = Q[qi] * 2.0 + 0.4 * torch.randn((30000, d), generator=g, device=dev)
        lo, hi = distributed.shard_range(nv, world, rank)
        shard = engine.CorpusStore(hi - lo, (d,), device=dev, index_offset=lo).add(V[lo:hi])
        full = engine.CorpusStore(nv, (d,), device=dev).add(V)
        st = {}
        s_ref, i_ref = full.search(Q, k)
        s, i = distributed.sharded_search(shard, Q, k, n_total=nv, stats=st)
        same = bool(torch.equal(i, i_ref)) and bool(torch.equal(s, s_ref))
        ok = ok and same and st.get("reruns", 0) >= 1
        if rank == 0:
            print("re-run case %-5s: %s, %d re-run pass(es), %d row(s)" % (case, "identical" if same else "MISMATCH",
                                                                       st.get("reruns", 0), st.get("rerun_rows", 0)), flush=True)
        del shard, full, V
        torch.cuda.empty_cache()
    host = torch.arange(1000 * 8, dtype=torch.float32).reshape(1000, 8).pin_memory()
    ok = ok and bool(torch.equal(distributed.upload_rows(host, device=dev).cpu(), host))
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("sharded search on %d GPUs: %s" % (world, "OK" if int(t) == 1 else "FAILED"), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(t) == 1 else 1)


if __name__ == "__main__":
    main()
