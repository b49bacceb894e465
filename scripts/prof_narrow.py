"""Narrow MM-UNet shape (B=16, D=6, L=65536): a few forward + backward launches for ncu / timing."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mm-unet_b200")); sys.path.insert(0, os.path.join(ROOT, "scripts"))
from mmunet_b200 import ops
from probe_scan import make
B, D, L = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (16, 6, 65536)))
u, delta, A, Bm, Cm, Dp, z, bias, dout = make(B, D, L, 16, torch.float32)
for _ in range(3):
    out, x, _ = ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True)
    g = ops.selective_scan_bwd(u, delta, A, Bm, Cm, Dp, z, bias, dout, x, True)
torch.cuda.synchronize()
print("ok")
