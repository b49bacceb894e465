"""Minimal driver for ncu: v5 forward (+ backward) launches.  usage: prof_v5.py [rcg]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mm-unet_b200")); sys.path.insert(0, os.path.join(ROOT, "scripts"))
from mmunet_b200 import ops
from probe_scan import make
shape = (16, 128, 4096) if (len(sys.argv) > 1 and sys.argv[1] == "rcg") else (8, 384, 4096)
u, delta, A, Bm, Cm, Dp, z, bias, dout = make(*shape, 16, torch.float32)
for _ in range(3):
    out, x, _ = ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True)
    g = ops.selective_scan_bwd(u, delta, A, Bm, Cm, Dp, z, bias, dout, x, True)
torch.cuda.synchronize()
print("ok")
