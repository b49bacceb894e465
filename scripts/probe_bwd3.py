"""Backward v3 at config 2 (fp32 / bf16) and the largest RCG shape; optional MMU_BWD3_DBG values as arguments (1 = no dB/dC atomics, 2 = no dA atomics)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mm-unet_b200")); sys.path.insert(0, ROOT)
from mmunet_b200 import ops
from scripts.probe_scan import make, timeit
for (B, D, L, dt) in ((8, 384, 4096, torch.float32), (8, 384, 4096, torch.bfloat16), (16, 128, 65536, torch.bfloat16)):
    u, delta, A, Bm, Cm, Dp, z, bias, dout = make(B, D, L, 16, dt)
    out, x, _ = ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True)
    du = torch.empty_like(u); dd = torch.empty_like(u); dz = torch.empty_like(u)
    for dbg in sys.argv[1:] or ["0"]:
        os.environ["MMU_BWD3_DBG"] = dbg
        t = timeit(lambda: ops.selective_scan_bwd(u, delta, A, Bm, Cm, Dp, z, bias, dout, x, True, du=du, ddelta=dd, dz=dz), warm=3, it=20)
        print(f"B{B} D{D} L{L} {str(dt)[6:]} dbg={dbg}: bwd {t:.1f} us", flush=True)
    del u, delta, Bm, Cm, z, dout, out, x, du, dd, dz
    torch.cuda.empty_cache()
