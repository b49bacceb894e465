"""Shared checks of single MM_Net blocks against tests/golden/mm_blocks.npz (reference MMConv / RCG run on CPU by
oracle/gen_golden.py:gen_mm_net).  Used by the CPU test (oracle stand-ins for the Mamba ops) and the GPU test (CUDA path)."""
import numpy as np
import torch

from conftest import load_golden

BLOCKS = load_golden("mm_blocks.npz")
MMCONV_CASES = sorted(k for k in BLOCKS if k.startswith("mmconv_"))


def _close(got, ref, tol, what, max_bad_frac=0.0):
    got = got.detach().float().cpu().numpy()
    s = max(float(np.abs(ref).max()), 1e-6)
    bad = np.abs(got - ref) > tol * (s + np.abs(ref))
    # bilinear sampling is piecewise linear in the sampled row coordinate: an element whose coordinate sits within rounding
    # distance of an integer row may take the neighbouring cell's slope, so a bounded fraction of outliers is tolerated
    assert bad.mean() <= max_bad_frac, f"{what}: {bad.sum()} / {bad.size} elements off by more than {tol} (max |diff| " \
                                       f"{np.abs(got - ref).max():.3e}, scale {s:.3e})"


def check_mmconv(mm_net, name, device, tol):
    c = BLOCKS[name]
    cin, cout, K = (int(v) for v in c["shape"])
    torch.manual_seed(7)
    conv = mm_net.MMConv(cin, cout, kernel_size=K, num_slices=4).to(device)
    x = torch.tensor(c["x"], device=device, requires_grad=True)
    out = conv(x)
    _close(out, c["out"], tol, name + ".out")
    out.backward(torch.tensor(c["dout"], device=device))
    _close(x.grad, c["dx"], 3 * tol, name + ".dx", 2e-3)
    params = dict(conv.named_parameters())
    for k in c:
        if k.startswith("grad:"):
            _close(params[k[5:]].grad, c[k], 3 * tol, f"{name}.{k}")


def check_rcg(mm_net, device, tol):
    c = BLOCKS["rcg"]
    torch.manual_seed(7)
    rcg = mm_net.RCG(num_slices=4).to(device).train()
    t = {k: torch.tensor(c[k], device=device, requires_grad=True) for k in ("pre", "edge", "f")}
    out = rcg(t["pre"], t["edge"], t["f"])
    _close(out, c["out"], tol, "rcg.out")
    out.backward(torch.tensor(c["dout"], device=device))
    for k in ("pre", "edge", "f"):
        _close(t[k].grad, c["d" + k], 3 * tol, "rcg.d" + k, 2e-3)
    params = dict(rcg.named_parameters())
    for k in c:
        if k.startswith("grad:"):
            _close(params[k[5:]].grad, c[k], 3 * tol, f"rcg.{k}")
