"""Backward chained-segment count sweep (MMU_BWD_CHAIN=k): config 2 and the three RCG stages of a 512x512 / batch 16 step."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mm-unet_b200")); sys.path.insert(0, ROOT)
from mmunet_b200 import ops
from scripts.probe_scan import make, timeit
for (B, D, L, dt) in ((8, 384, 4096, torch.float32), (8, 384, 4096, torch.bfloat16), (16, 128, 4096, torch.bfloat16),
                      (16, 128, 16384, torch.bfloat16), (16, 128, 65536, torch.bfloat16), (16, 128, 65536, torch.float32)):
    u, delta, A, Bm, Cm, Dp, z, bias, dout = make(B, D, L, 16, dt)
    out, x, _ = ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True)
    du = torch.empty_like(u); dd = torch.empty_like(u); dz = torch.empty_like(u)
    res = []
    for k in (0, 1, 2, 3, 4, 6, 8, 12, 16, 24, 32):
        if k:
            os.environ["MMU_BWD_CHAIN"] = str(k)
        else:
            os.environ.pop("MMU_BWD_CHAIN", None)
        if k > (L + 255) // 256:
            continue
        t = timeit(lambda: ops.selective_scan_bwd(u, delta, A, Bm, Cm, Dp, z, bias, dout, x, True, du=du, ddelta=dd, dz=dz), warm=3, it=10)
        res.append(f"{'auto' if k == 0 else k}:{t:.0f}")
    print(f"B{B} D{D} L{L} {str(dt)[6:]}: bwd us by chain count  " + "  ".join(res), flush=True)
    del u, delta, Bm, Cm, z, dout, out, x, du, dd, dz
    torch.cuda.empty_cache()
