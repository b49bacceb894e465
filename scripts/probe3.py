"""v1 vs v3 scan kernels: agreement + timing on the BASELINE config-2 shape.  Not a test."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mm-unet_b200")); sys.path.insert(0, ROOT)
from mmunet_b200 import ops
from scripts.probe_scan import make, timeit

def run(B, D, L, N, dtype, bwd=True):
    u, delta, A, Bm, Cm, Dp, z, bias, dout = make(B, D, L, N, dtype)
    s = u.element_size(); fb = (4 * D + 2 * N) * B * L * s; bb = (7 * D + 4 * N) * B * L * s
    res = {}
    for v in ("1", "3"):
        os.environ["MMU_SCAN_V"] = v
        out, x, last = ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True, return_last_state=True)
        g = ops.selective_scan_bwd(u, delta, A, Bm, Cm, Dp, z, bias, dout, x, True) if bwd else None
        torch.cuda.synchronize()
        tf = timeit(lambda: ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True))
        tb = timeit(lambda: ops.selective_scan_bwd(u, delta, A, Bm, Cm, Dp, z, bias, dout, x, True)) if bwd else float("nan")
        res[v] = (out.float(), x, last, g)
        print(f"v{v} B{B} D{D} L{L} N{N} {str(dtype)[6:]}: fwd {tf:8.1f} us {fb / tf / 1e3:6.0f} GB/s | bwd {tb:8.1f} us {bb / tb / 1e3:6.0f} GB/s", flush=True)
    o1, x1, l1, g1 = res["1"]; o3, x3, l3, g3 = res["3"]
    print(f"   max|out1-out3| {(o1 - o3).abs().max().item():.3e} (|out| max {o1.abs().max().item():.2e})  "
          f"x {(x1.x - x3.x).abs().max().item():.3e}  last {(l1 - l3).abs().max().item():.3e}", flush=True)
    if bwd:
        names = ["du", "ddelta", "dA", "dB", "dC", "dD", "dz", "dbias"]
        print("   grads:", "  ".join(f"{n} {(a.float() - b.float()).abs().max().item():.2e}/{a.float().abs().max().item():.1e}"
                                      for n, a, b in zip(names, g1, g3) if a is not None), flush=True)

if __name__ == "__main__":
    bwd = "--nobwd" not in sys.argv
    run(2, 6, 1024, 16, torch.float32, bwd)
    run(8, 384, 4096, 16, torch.float32, bwd)
    run(8, 384, 4096, 16, torch.bfloat16, bwd)
    run(16, 6, 65536, 16, torch.float32, bwd)
    run(16, 128, 16384, 16, torch.bfloat16, bwd)
