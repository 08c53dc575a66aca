"""Bring-up diagnostics for the tcgen05 score kernel (run on the B200 box): prints where a STORE-mode tile
disagrees with the fp32 matmul of the same bf16 operands, so descriptor / swizzle / TMEM-lane mistakes can be
told apart from a single run."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cross_modal_video_engine_b200 import _native as N

N.require_device()
torch.manual_seed(0)


def run(nq, nv, k, pattern="rand"):
    a = torch.zeros(((nq + 127) // 128 * 128, k), dtype=torch.bfloat16, device="cuda")
    b = torch.zeros(((nv + 255) // 256 * 256, k), dtype=torch.bfloat16, device="cuda")
    if pattern == "rand":
        a[:nq] = torch.randn(nq, k, device="cuda").to(torch.bfloat16)
        b[:nv] = torch.randn(nv, k, device="cuda").to(torch.bfloat16)
    else:  # one-hot rows: out[q, v] = 1 iff (q % k) == (v % k)  -> shows K-offset / row permutation errors
        a[torch.arange(nq), torch.arange(nq) % k] = 1
        b[torch.arange(nv), torch.arange(nv) % k] = 1
    out = torch.full((nq, nv), float("nan"), device="cuda")
    N.call("xmve_score_store", N.ptr(a), nq, k, N.ptr(b), nv, k, 1, k, 1.0, N.ptr(out), nv, N.stream_ptr())
    torch.cuda.synchronize()
    ref = a[:nq].float() @ b[:nv].float().T
    err = (out - ref).abs()
    bad = ~(err < 1e-3)
    print("nq=%d nv=%d k=%d %s: max err %.3g, bad %.4f%%, nan %d" % (nq, nv, k, pattern, float(err.nan_to_num(9e9).max()),
          100 * bad.float().mean().item(), int(torch.isnan(out).sum())))
    if bad.any():
        print("  bad rows (first 16):", torch.nonzero(bad.any(1)).flatten()[:16].tolist())
        print("  bad cols (first 16):", torch.nonzero(bad.any(0)).flatten()[:16].tolist())
        r, c = torch.nonzero(bad)[0].tolist()
        print("  first bad (%d,%d): got %g want %g" % (r, c, out[r, c].item(), ref[r, c].item()))
        if pattern != "rand":
            got = torch.nonzero(out[r] > 0.5).flatten()[:8].tolist()
            want = torch.nonzero(ref[r] > 0.5).flatten()[:8].tolist()
            print("  row %d ones at: got %s want %s" % (r, got, want))
    return not bad.any()


ok = True
for args in [(128, 256, 64, "onehot"), (128, 256, 64, "rand"), (128, 256, 128, "onehot"), (128, 256, 256, "rand"),
             (256, 512, 64, "rand"), (100, 300, 192, "rand"), (1000, 1000, 1536, "rand"), (300, 40000, 512, "rand")]:
    try:
        ok &= run(*args)
    except Exception as e:
        print("EXC", args, e)
        ok = False
        break
print("ALL OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
