"""Kernel timeline of ONE sharded search step on rank 0 (torch.profiler / CUPTI), under torchrun:

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/trace_step.py [nv] [nq] [k] [dim]

Prints every GPU activity of the step in start order with its duration and the idle gap before it -- the tool
behind the tail analysis of the multi-GPU step (where ncu cannot go: it must not wrap a multi-rank command)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cross_modal_video_engine_b200 import distributed, engine, synth  # noqa: E402


def main():
    nv = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    nq = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    dims, w = (1536, 512), (0.6, 0.4)
    k = int(sys.argv[3]) if len(sys.argv) > 3 else 100
    if len(sys.argv) > 4:                                   # one space of that many dims (C4: 640)
        dims, w = (int(sys.argv[4]),), (1.0,)
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lo, hi = distributed.shard_range(nv, world, rank)
    store = engine.CorpusStore(hi - lo, dims, device=dev, index_offset=lo)
    buf = torch.empty((250_000, sum(dims)), dtype=torch.float32, device=dev)
    for c in range(lo // 250_000, (hi - 1) // 250_000 + 1):
        synth.device_gaussian(250_000, sum(dims), 4 * 100003 + c, dev, out=buf)
        a, b = max(lo, c * 250_000), min(hi, (c + 1) * 250_000)
        store.add(buf[a - c * 250_000: b - c * 250_000])
    del buf
    q = synth.device_gaussian(nq, sum(dims), 5, dev)
    for _ in range(3):
        distributed.sharded_search(store, q, k, weights=w, n_total=nv)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        pend = [distributed.sharded_search(store, q, k, weights=w, n_total=nv, defer=True) for _ in range(2)]
        for p in pend:
            p.result()
        torch.cuda.synchronize()
    if rank == 0:
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        evs.sort(key=lambda e: e.time_range.start)
        # the second search: from its first K1 launch
        starts = [i for i, e in enumerate(evs) if "prepare_rows" in e.name]
        first = starts[len(starts) // 2] if starts else 0
        prev_end = None
        total = 0.0
        for e in evs[first:]:
            st, en = e.time_range.start, e.time_range.end
            gap = (st - prev_end) if prev_end is not None else 0.0
            print("%9.1f us  gap %7.1f  %s" % (en - st, gap, e.name[:90]))
            prev_end = max(prev_end or en, en)
            total += en - st
        print("sum of activities %.1f us over %d launches" % (total, len(evs) - first))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
