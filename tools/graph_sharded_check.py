"""Multi-rank CUDA-graph replay of a sharded search (NCCL all-gathers captured inside the graph), under torchrun:

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/graph_sharded_check.py

For each shape: `engine.GraphSearch(shard, ..., comm=GroupComm())` must return what the eager sharded search returns,
for several query batches, and its replay time is printed next to the eager pipelined loop's."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cross_modal_video_engine_b200 import distributed, engine, synth  # noqa: E402


def loop_ms(fn, reps):
    prev = None
    for _ in range(3):
        fn().result()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        p = fn()
        if prev is not None:
            prev.result()
        prev = p
    prev.result()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    comm = distributed.GroupComm()
    ok = True
    for tag, nv, nq, d, k, excl in (("C3/8 shape", 270_000, 60, 2048, 1000, False), ("C4/8 shape", 250_000, 4096, 640, 100, True)):
        lo, hi = distributed.shard_range(nv, world, rank)
        shard = engine.CorpusStore(hi - lo, (d,), device=dev, index_offset=lo)
        shard.add(synth.device_gaussian(nv, d, 7, dev)[lo:hi])      # the same seeded corpus on every rank
        qs = [synth.device_gaussian(nq, d, 20 + b, dev) for b in range(3)]
        ex = torch.randint(0, nv, (nq,), device=dev, generator=torch.Generator(device=dev).manual_seed(3)) if excl else None
        gs = engine.GraphSearch(shard, nq, k, with_exclude=excl, comm=comm, n_total=nv)
        for q in qs:
            s_ref, i_ref = engine.search_shards([shard], q, k, exclude=ex, comm=comm, n_total=nv)
            s, i = gs(q, exclude=ex)
            same = bool(torch.equal(i, i_ref)) and bool(torch.equal(s, s_ref))
            ok = ok and same
        head = torch.cuda.Stream(device=dev)
        t_eager = loop_ms(lambda: engine.search_shards([shard], qs[0], k, exclude=ex, comm=comm, n_total=nv, defer=True), 20)
        t_head = loop_ms(lambda: engine.search_shards([shard], qs[0], k, exclude=ex, comm=comm, n_total=nv, defer=True,
                                                      head_stream=head), 20)
        t_graph = loop_ms(lambda: gs(qs[0], exclude=ex, defer=True), 20)
        if rank == 0:
            print("%s on %d GPUs: graph replay %s; eager %.3f ms, eager + head stream %.3f ms, graph replay %.3f ms per search"
                  % (tag, world, "identical" if ok else "MISMATCH", t_eager, t_head, t_graph), flush=True)
        del gs, shard
        torch.cuda.empty_cache()
    if rank == 0:
        print("graph replay over NCCL: %s" % ("OK" if ok else "FAILED"), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
