import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mm-unet_b200")); sys.path.insert(0, ROOT)
from mmunet_b200 import ops
from scripts.probe_scan import make, timeit
u, delta, A, Bm, Cm, Dp, z, bias, dout = make(8, 384, 4096, 16, torch.float32)
out, x, _ = ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True)
for ns in ("1", "2", "3", "4", "8"):
    os.environ["MMU_FWD_NSEG"] = ns; os.environ["MMU_BWD_NSEG"] = ns
    tf = timeit(lambda: ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True))
    tb = timeit(lambda: ops.selective_scan_bwd(u, delta, A, Bm, Cm, Dp, z, bias, dout, x, True))
    print(f"nseg={ns}: fwd {tf:.1f} us  bwd {tb:.1f} us", flush=True)
