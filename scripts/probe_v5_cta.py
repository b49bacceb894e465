"""Ring forward time vs CTA count / CTA width (fp32).  Not a test."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mm-unet_b200")); sys.path.insert(0, ROOT)
from mmunet_b200 import ops, _lib
from scripts.probe_scan import make, timeit
for (B, D, L, W) in ((4, 384, 4096, 6), (5, 384, 4096, 6), (6, 384, 4096, 6), (7, 384, 4096, 6), (8, 384, 4096, 6), (9, 384, 4096, 6), (4, 384, 4096, 4), (6, 384, 4096, 4),
                     (8, 256, 4096, 4), (9, 256, 4096, 4), (8, 384, 4096, 2)):
    u, delta, A, Bm, Cm, Dp, z, bias, dout = make(B, D, L, 16, torch.float32)
    os.environ["MMU_RING"] = "1"; os.environ["MMU_V5_MIN_WARPS"] = "1"; os.environ["MMU_V5_W"] = str(W); _lib.reload_knobs()
    ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True)
    t = timeit(lambda: ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True), warm=3, it=20)
    print(f"B{B} D{D} L{L} W={W}: CTAs {B * ((D + 4 * W - 1) // (4 * W))}  ring {t:.0f} us", flush=True)
