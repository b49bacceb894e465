"""Where the lane-ring forward beats v3: forward time of both over a grid of shapes (fp32).  Not a test."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mm-unet_b200")); sys.path.insert(0, ROOT)
from mmunet_b200 import ops, _lib
from scripts.probe_scan import make, timeit
DT = torch.bfloat16 if (len(sys.argv) > 1 and sys.argv[1] == "bf16") else torch.float32
for (B, D, L) in ((2, 128, 4096), (4, 128, 4096), (8, 128, 4096), (4, 384, 4096), (2, 384, 4096), (16, 64, 4096), (8, 384, 1024), (8, 384, 512),
                  (16, 128, 65536), (8, 384, 16384), (8, 96, 4096), (3, 200, 2048)):
    u, delta, A, Bm, Cm, Dp, z, bias, dout = make(B, D, L, 16, DT)
    t = {}
    for ring in (0, 1):
        os.environ["MMU_RING"] = str(ring); os.environ["MMU_V5_MIN_WARPS"] = "1"; os.environ["MMU_RING_BF16"] = "1"; _lib.reload_knobs()
        ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True)
        t[ring] = timeit(lambda: ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True), warm=3, it=20)
    print(f"B{B} D{D} L{L}: rows/4 = {B * ((D + 3) // 4)}  v3 {t[0]:.0f} us  ring {t[1]:.0f} us  ratio {t[1] / t[0]:.2f}", flush=True)
