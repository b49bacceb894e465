import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mm-unet_b200")); sys.path.insert(0, ROOT)
from mmunet_b200 import ops
from scripts.probe_scan import make, timeit
for dt in (torch.float32, torch.bfloat16):
    u, delta, A, Bm, Cm, Dp, z, bias, dout = make(8, 384, 4096, 16, dt)
    out, x, _ = ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True)
    os.environ["MMU_BWD3_W"] = "4"; os.environ.pop("MMU_BWD_CHAIN", None)
    ref = ops.selective_scan_bwd(u, delta, A, Bm, Cm, Dp, z, bias, dout, x, True)
    for w, ch in (("4", "3"), ("8", "1"), ("8", "2"), ("8", "3"), ("8", "4")):
        os.environ["MMU_BWD3_W"] = w; os.environ["MMU_BWD_CHAIN"] = ch
        g = ops.selective_scan_bwd(u, delta, A, Bm, Cm, Dp, z, bias, dout, x, True)
        torch.cuda.synchronize()
        err = max(float((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-9)) for a, b in zip(g, ref) if a is not None)
        t = timeit(lambda: ops.selective_scan_bwd(u, delta, A, Bm, Cm, Dp, z, bias, dout, x, True))
        print(f"{dt} W={w} chain={ch}: bwd {t:.1f} us  err {err:.1e}", flush=True)
