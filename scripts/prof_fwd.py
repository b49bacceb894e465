"""Minimal driver for ncu: forward (and optionally backward) scan launches at BASELINE config 2."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mm-unet_b200")); sys.path.insert(0, os.path.join(ROOT, "scripts"))
from mmunet_b200 import ops
from probe_scan import make
dt = torch.bfloat16 if "bf16" in sys.argv else torch.float32
u, delta, A, Bm, Cm, Dp, z, bias, dout = make(8, 384, 4096, 16, dt)
for _ in range(3):
    out, x, _ = ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True)
    if "bwd" in sys.argv:
        g = ops.selective_scan_bwd(u, delta, A, Bm, Cm, Dp, z, bias, dout, x, True)
torch.cuda.synchronize()
print("ok")
