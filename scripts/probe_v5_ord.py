"""Ring forward with fused scan orders: time against v3 (fp32, config 2).  Not a test."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mm-unet_b200")); sys.path.insert(0, ROOT)
from mmunet_b200 import ops, _lib
from scripts.probe_scan import make, timeit
B, D, L = 8, 384, 4096
u, delta, A, Bm, Cm, Dp, z, bias, dout = make(B, D, L, 16, torch.float32)
for name, od in (("plain", None), ("nslices16", (_lib.ORDER_NSLICES, 1, L, 16)), ("nslices32", (_lib.ORDER_NSLICES, 1, L, 32)), ("nslices8", (_lib.ORDER_NSLICES, 1, L, 8)),
                 ("two-row", (_lib.ORDER_TWOROW, 64, 64, 1))):
    t = {}
    for ring in (0, 1):
        os.environ["MMU_RING"] = str(ring); _lib.reload_knobs()
        ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True, order=od)
        t[ring] = timeit(lambda: ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True, order=od), warm=3, it=20)
    print(f"{name}: v3 {t[0]:.0f} us  ring {t[1]:.0f} us", flush=True)
