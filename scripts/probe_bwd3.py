import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mm-unet_b200")); sys.path.insert(0, ROOT)
from mmunet_b200 import ops
from scripts.probe_scan import make, timeit
dt = torch.float32
for (B, D) in ((8, 384), (8, 296), (4, 296)):
    u, delta, A, Bm, Cm, Dp, z, bias, dout = make(B, D, 4096, 16, dt)
    out, x, _ = ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True)
    for pad in ("0", "60000"):
        os.environ["MMU_BWD3_SMEM_PAD"] = pad
        t = timeit(lambda: ops.selective_scan_bwd(u, delta, A, Bm, Cm, Dp, z, bias, dout, x, True))
        print(f"B{B} D{D} ctas {B * ((D + 7) // 8)} pad={pad}: bwd {t:.1f} us", flush=True)
