import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mm-unet_b200")); sys.path.insert(0, ROOT)
from mmunet_b200 import ops
from scripts.probe_scan import make, timeit
for dt in (torch.float32, torch.bfloat16):
    u, delta, A, Bm, Cm, Dp, z, bias, dout = make(8, 384, 4096, 16, dt)
    os.environ["MMU_SCAN_V"] = "1"
    ref = ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True)[0].float()
    os.environ["MMU_SCAN_V"] = "3"
    for lpr in ("16", "32"):
        os.environ["MMU_FWD3_LPR"] = lpr
        out = ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True)[0].float()
        t = timeit(lambda: ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True))
        print(f"{dt} LPR={lpr}: fwd {t:.1f} us  maxdiff {(out - ref).abs().max().item():.3e}", flush=True)
