import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mm-unet_b200")); sys.path.insert(0, ROOT)
from mmunet_b200 import ops
from scripts.probe_scan import make, timeit
for (B, D, L) in ((16, 6, 256), (16, 6, 1024), (16, 6, 4096), (16, 6, 16384), (16, 2, 4096), (16, 128, 4096)):
    u, delta, A, Bm, Cm, Dp, z, bias, dout = make(B, D, L, 16, torch.float32)
    res = []
    for ns in ("1", "2", "4", "8", "16", "auto"):
        if ns == "auto":
            os.environ.pop("MMU_FWD_NSEG", None); os.environ.pop("MMU_BWD_NSEG", None)
        else:
            if int(ns) > max(1, L // 256): continue
            os.environ["MMU_FWD_NSEG"] = ns; os.environ["MMU_BWD_NSEG"] = ns
        out, x, _ = ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True)
        tf = timeit(lambda: ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True), warm=5, it=20)
        tb = timeit(lambda: ops.selective_scan_bwd(u, delta, A, Bm, Cm, Dp, z, bias, dout, x, True), warm=5, it=20)
        res.append(f"{ns}:{tf:.0f}/{tb:.0f}")
    print(f"B{B} D{D} L{L}: nseg fwd/bwd us  " + "  ".join(res), flush=True)
