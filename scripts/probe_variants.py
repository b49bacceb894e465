"""Times scan fwd / bwd at config 2 (fp32, bf16) and the RCG 64k shape for the library selected by MMU_LIB.  Not a test."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mm-unet_b200")); sys.path.insert(0, ROOT)
from mmunet_b200 import ops
from scripts.probe_scan import make, timeit
tag = os.path.basename(os.environ.get("MMU_LIB", "default"))
for (B, D, L, dt) in ((8, 384, 4096, torch.float32), (8, 384, 4096, torch.bfloat16), (16, 128, 65536, torch.bfloat16), (16, 6, 65536, torch.bfloat16)):
    u, delta, A, Bm, Cm, Dp, z, bias, dout = make(B, D, L, 16, dt)
    out, xs, _ = ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True)
    tf = timeit(lambda: ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True), warm=3, it=20)
    du = torch.empty_like(u); dd = torch.empty_like(u); dz = torch.empty_like(u)
    tb = timeit(lambda: ops.selective_scan_bwd(u, delta, A, Bm, Cm, Dp, z, bias, dout, xs, True, du=du, ddelta=dd, dz=dz), warm=3, it=20)
    print(f"{tag}: B{B} D{D} L{L} {str(dt)[6:]}: fwd {tf:.0f} us bwd {tb:.0f} us", flush=True)
