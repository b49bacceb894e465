import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mm-unet_b200")); sys.path.insert(0, ROOT)
from mmunet_b200 import ops
from scripts.probe_scan import make, timeit
u, delta, A, Bm, Cm, Dp, z, bias, dout = make(8, 384, 4096, 16, torch.float32)
out, x, _ = ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True)
os.environ["MMU_BWD_CHAIN"] = "1"
ref = ops.selective_scan_bwd(u, delta, A, Bm, Cm, Dp, z, bias, dout, x, True)
for ns in ("1", "2", "3", "4", "5", "8"):
    os.environ["MMU_BWD_CHAIN"] = ns
    g = ops.selective_scan_bwd(u, delta, A, Bm, Cm, Dp, z, bias, dout, x, True)
    torch.cuda.synchronize()
    err = max(float((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-9)) for a, b in zip(g, ref) if a is not None)
    tb = timeit(lambda: ops.selective_scan_bwd(u, delta, A, Bm, Cm, Dp, z, bias, dout, x, True))
    print(f"chain={ns}: bwd {tb:.1f} us  max rel-to-max err vs unchained {err:.2e}", flush=True)
