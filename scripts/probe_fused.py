"""Fused scan order vs explicit gather / scatter at the inner-function level (fwd + bwd), RCG and MMConv shapes.  Not a test."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mm-unet_b200")); sys.path.insert(0, ROOT)
from mmunet_b200 import _lib, ops
from scripts.probe_scan import timeit

def run(B, d, L, order, dt, R):
    torch.manual_seed(0)
    N = 16
    xz = torch.randn(B, 2 * d, L, device="cuda").to(dt).requires_grad_()
    cw, cb = torch.randn(d, 1, 4, device="cuda") * 0.5, torch.randn(d, device="cuda") * 0.5
    xw, dw = (torch.randn(R + 2 * N, d, device="cuda") * 0.1).to(dt), (torch.randn(d, R, device="cuda") * 0.3).to(dt)
    A, D, db = -0.5 * torch.rand(d, N, device="cuda"), torch.randn(d, device="cuda"), 0.5 * torch.rand(d, device="cuda")
    g = torch.randn(B, d, L, device="cuda").to(dt)
    res = {}
    for fuse in ("1", "0"):
        os.environ["MMU_FUSE"] = fuse
        _lib.reload_knobs()
        def f():
            return ops.mamba_inner_fn_no_out_proj_ordered(xz, cw, cb, xw, dw, A, D, db, True, order=order)
        def fb():
            xz.grad = None
            f().backward(g)
        n0 = _lib.launch_count(); fb(); nl = _lib.launch_count() - n0
        res[fuse] = (timeit(f, warm=3, it=15), timeit(fb, warm=3, it=15), nl)
        # separate kernels
    os.environ.pop("MMU_FUSE"); _lib.reload_knobs()
    x = xz.detach()[:, :d]
    tc_f = timeit(lambda: ops.causal_conv1d_fwd(x, cw.view(d, 4), cb, True, order=order), warm=3, it=15)
    tc_e = timeit(lambda: ops.causal_conv1d_fwd(x, cw.view(d, 4), cb, True), warm=3, it=15)
    print(f"B{B} d{d} L{L} order{order} {str(dt)[6:]}: fused fwd {res['1'][0]:.0f} us fwd+bwd {res['1'][1]:.0f} us ({res['1'][2]} launches) | "
          f"explicit fwd {res['0'][0]:.0f} fwd+bwd {res['0'][1]:.0f} ({res['0'][2]} launches) | conv fwd ordered {tc_f:.0f} vs plain {tc_e:.0f} us", flush=True)

for dt in (torch.bfloat16, torch.float32):
    run(16, 128, 4096, (_lib.ORDER_NSLICES, 1, 4096, 16), dt, 4)
    run(16, 128, 16384, (_lib.ORDER_NSLICES, 1, 16384, 32), dt, 4)
    run(16, 128, 65536, (_lib.ORDER_NSLICES, 1, 65536, 64), dt, 4)
    run(16, 6, 65536, (_lib.ORDER_TWOROW, 256, 256, 1), dt, 1)
    run(16, 6, 4096, (_lib.ORDER_TWOROW, 64, 64, 1), dt, 1)
    run(16, 6, 256, (_lib.ORDER_TWOROW, 16, 16, 1), dt, 1)
