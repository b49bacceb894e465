"""Forward v3 variants: MMU_FWD3_W in {4, 8} (8 = 128 registers, one state per iteration, 16 warps/SM) at config 2 and the RCG shapes."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mm-unet_b200")); sys.path.insert(0, ROOT)
from mmunet_b200 import ops
from scripts.probe_scan import make, timeit
for (B, D, L, dt) in ((8, 384, 4096, torch.float32), (8, 384, 4096, torch.bfloat16), (16, 128, 16384, torch.bfloat16), (16, 128, 65536, torch.bfloat16)):
    u, delta, A, Bm, Cm, Dp, z, bias, dout = make(B, D, L, 16, dt)
    res = []
    ref = None
    for w in ("4", "8"):
        os.environ["MMU_FWD3_W"] = w
        out, xs, _ = ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True)
        if ref is None:
            ref = out.float().clone()
        err = float((out.float() - ref).abs().max())
        t = timeit(lambda: ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True), warm=3, it=20)
        res.append(f"W={w}: {t:.0f} us (max diff vs W=4 {err:.1e})")
    print(f"B{B} D{D} L{L} {str(dt)[6:]}: " + "  ".join(res), flush=True)
    os.environ.pop("MMU_FWD3_W", None)
