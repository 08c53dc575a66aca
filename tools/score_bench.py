"""Micro-benchmark of the K2 filter kernel alone (B200): TFLOP/s per tile shape and problem size.

    python tools/score_bench.py [--nv 2000000] [--nq 8192] [--k 2048] [--tiles 2,1] [--reps 5]

Operands are random unit-norm bf16 rows; the threshold is far above every score so that the candidate append
path stays idle (what is timed is TMA + tcgen05 + the TMEM scan).  Not a bench.py line -- a tuning tool.
"""
import argparse
import os
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cross_modal_video_engine_b200 import _native as N  # noqa: E402


class Nvml:
    """SM clock / power sampled every few ms on a thread while a timed loop runs."""

    def __init__(self, index=0):
        import pynvml
        self.nv = pynvml
        pynvml.nvmlInit()
        self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
        self.samples, self.stop_flag, self.t = [], False, None

    def start(self):
        self.samples, self.stop_flag = [], False

        def loop():
            while not self.stop_flag:
                self.samples.append((self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM),
                                     self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0))
                time.sleep(0.004)
        self.t = threading.Thread(target=loop, daemon=True)
        self.t.start()

    def stop(self):
        self.stop_flag = True
        self.t.join()
        s = self.samples[len(self.samples) // 4:] or self.samples      # drop the ramp-up quarter
        clk = sorted(x[0] for x in s)
        pw = sorted(x[1] for x in s)
        return clk[len(clk) // 2], clk[0], clk[-1], pw[len(pw) // 2], len(s)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nv", type=int, default=2_000_000)
    ap.add_argument("--nq", type=int, default=8192)
    ap.add_argument("--k", type=int, default=2048)
    ap.add_argument("--tiles", default="2,1")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--thr", type=float, default=0.5)
    ap.add_argument("--cap", type=int, default=256)
    ap.add_argument("--sustain", type=float, default=0.0, help="seconds of back-to-back launches with clock sampling")
    ap.add_argument("--store", action="store_true", help="time the STORE epilogue (fp32 score matrix written) instead")
    ap.add_argument("--step", type=int, default=1, help="corpus row stride (the strided sample view)")
    ap.add_argument("--cublas", action="store_true", help="also time torch.matmul on (a chunk of) the same operands")
    args = ap.parse_args()
    N.require_device()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)

    def unit(n):
        out = torch.empty((n + 511) // 512 * 512, args.k, dtype=torch.bfloat16, device=dev)
        for r0 in range(0, n, 262144):
            r1 = min(n, r0 + 262144)
            x = torch.randn((r1 - r0, args.k), generator=g, device=dev)
            out[r0:r1] = (x / x.norm(dim=1, keepdim=True)).to(torch.bfloat16)
        return out

    a, b = unit(args.nq), unit(args.nv * args.step)
    out = torch.empty((args.nq, args.nv), dtype=torch.float32, device=dev) if args.store else None
    lo = torch.full((args.nq,), args.thr, device=dev)
    cap = args.cap
    cc = torch.zeros(args.nq, dtype=torch.int32, device=dev)
    cs = torch.empty((args.nq, cap), dtype=torch.float32, device=dev)
    ci = torch.empty((args.nq, cap), dtype=torch.int32, device=dev)
    flops = 2.0 * args.nq * args.nv * args.k
    tiles = [int(t) for t in args.tiles.split(",")]
    if "XMVE_TILE" in os.environ:
        tiles = [int(os.environ["XMVE_TILE"])]
    for tile in tiles:
        os.environ["XMVE_TILE"] = str(tile)

        def run():
            if args.store:
                N.call("xmve_score_store", N.ptr(a), args.nq, a.stride(0), N.ptr(b), args.nv, b.stride(0), args.step, args.k,
                       1.0, N.ptr(out), out.stride(0), N.stream_ptr())
                return
            cc.zero_()
            N.call("xmve_score_filter", N.ptr(a), args.nq, a.stride(0), N.ptr(b), args.nv, b.stride(0), 1, args.k, N.ptr(lo),
                   None, None, N.ptr(cc), N.ptr(cs), N.ptr(ci), cap, N.stream_ptr())
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.reps)]
        for e0, e1 in ev:
            e0.record()
            run()
            e1.record()
        torch.cuda.synchronize()
        ms = sorted(e0.elapsed_time(e1) for e0, e1 in ev)
        if args.store:                                             # spot check against a matmul of the same operands
            ref = a[:64].float() @ b[:: args.step][: args.nv].float().T
            err = float((out[:64] - ref).abs().max())
            print("store check: max |out - fp32 matmul| over 64 rows = %.2e" % err, flush=True)
            assert err < 1e-4
        print("tile=%d nq=%d nv=%d k=%d  ms min/med/max %.3f %.3f %.3f  TFLOP/s(med) %.1f  cand_max %d" % (
            tile, args.nq, args.nv, args.k, ms[0], ms[len(ms) // 2], ms[-1], flops / ms[len(ms) // 2] / 1e9,
            int(cc.max())), flush=True)
        if args.sustain > 0:
            sustained("tile=%d" % tile, run, ms[len(ms) // 2], flops, args.sustain)
    if args.cublas:
        nvc = min(args.nv, 262144)
        out = torch.empty((args.nq, nvc), dtype=torch.bfloat16, device=dev)
        bt = b[:nvc].T

        def mm():
            torch.matmul(a[:args.nq], bt, out=out)
        for _ in range(3):
            mm()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            mm()
        e1.record()
        torch.cuda.synchronize()
        ms1 = e0.elapsed_time(e1) / args.reps
        fl = 2.0 * args.nq * nvc * args.k
        print("cublas nq=%d nv=%d k=%d (+bf16 store)  ms %.3f  TFLOP/s %.1f" % (args.nq, nvc, args.k, ms1, fl / ms1 / 1e9),
              flush=True)
        if args.sustain > 0:
            sustained("cublas", mm, ms1, fl, args.sustain)


def sustained(tag, run, ms_one, flops, seconds):
    n = max(3, int(seconds * 1e3 / ms_one))
    nv = Nvml(torch.cuda.current_device())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    nv.start()
    e0.record()
    for _ in range(n):
        run()
    e1.record()
    torch.cuda.synchronize()
    clk, cmin, cmax, pw, ns = nv.stop()
    ms = e0.elapsed_time(e1) / n
    tf = flops / ms / 1e9
    print("  sustained %s: %d launches, %.3f ms each, %.1f TFLOP/s, SM clock med %d MHz [%d, %d], power %.0f W, "
          "tensor util at that clock %.1f %% (%d samples)" % (tag, n, ms, tf, clk, cmin, cmax, pw,
                                                            100.0 * tf * 1e12 / (148 * 8192 * clk * 1e6), ns), flush=True)


if __name__ == "__main__":
    main()
