"""Timing probe for the scan / conv kernels (CUDA events on the launching stream).  Not a test."""
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mm-unet_b200"))
from mmunet_b200 import ops  # noqa: E402


def timeit(fn, warm=10, it=50):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(it)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2] * 1e3  # us


def make(B, D, L, N, dtype):
    torch.manual_seed(0)
    dev = "cuda"
    u = torch.randn(B, D, L, device=dev).to(dtype)
    z = torch.randn(B, D, L, device=dev).to(dtype)
    delta = (0.5 * torch.rand(B, D, L, device=dev)).to(dtype)
    A = -0.5 * torch.rand(D, N, device=dev)
    Bm = torch.randn(B, 1, N, L, device=dev).to(dtype)
    Cm = torch.randn(B, 1, N, L, device=dev).to(dtype)
    Dp = torch.randn(D, device=dev)
    bias = 0.5 * torch.rand(D, device=dev)
    dout = torch.randn(B, D, L, device=dev).to(dtype)
    return u, delta, A, Bm, Cm, Dp, z, bias, dout


def probe(B, D, L, N, dtype, label):
    u, delta, A, Bm, Cm, Dp, z, bias, dout = make(B, D, L, N, dtype)
    s = u.element_size()
    fb = (4 * D + 2 * N) * B * L * s
    bb = (7 * D + 4 * N) * B * L * s
    out, x, _ = ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True)
    du = torch.empty_like(u); dd = torch.empty_like(u); dz = torch.empty_like(u)
    tf = timeit(lambda: ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True))
    tb = timeit(lambda: ops.selective_scan_bwd(u, delta, A, Bm, Cm, Dp, z, bias, dout, x, True, du=du, ddelta=dd, dz=dz))
    print(f"{label} B{B} D{D} L{L} N{N} {str(dtype)[6:]}: fwd {tf:8.1f} us {fb / tf / 1e3:7.0f} GB/s | bwd {tb:8.1f} us "
          f"{bb / tb / 1e3:7.0f} GB/s | fwd+bwd {tf + tb:8.1f} us {(fb + bb) / (tf + tb) / 1e3:7.0f} GB/s", flush=True)


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "cfg"
    if mode == "cfg":
        for fc in range(5):
            os.environ["MMU_FWD_CFG"] = str(fc)
            os.environ["MMU_BWD_CFG"] = str(fc % 3)
            probe(8, 384, 4096, 16, torch.float32, f"f{fc}b{fc % 3}")
        os.environ["MMU_FWD_CFG"] = "0"; os.environ["MMU_BWD_CFG"] = "0"
        probe(8, 384, 4096, 16, torch.bfloat16, "bf16")
    elif mode == "bcfg":
        os.environ["MMU_FWD_CFG"] = "0"
        for bc in (0, 3, 4):
            os.environ["MMU_BWD_CFG"] = str(bc)
            probe(8, 384, 4096, 16, torch.float32, f"b{bc}")
    elif mode == "narrow":
        os.environ.pop("MMU_FWD_CFG", None); os.environ.pop("MMU_BWD_CFG", None)
        for L in (1024, 4096, 16384, 65536):
            probe(16, 6, L, 16, torch.float32, "narrow")
        for ns in (1, 4, 16, 32, 64):
            os.environ["MMU_FWD_NSEG"] = str(ns); os.environ["MMU_BWD_NSEG"] = str(ns)
            probe(16, 6, 65536, 16, torch.float32, f"nseg{ns}")
    elif mode == "conv":
        for dtype in (torch.float32, torch.bfloat16):
            x = torch.randn(8, 384, 4096, device="cuda").to(dtype)
            w, b = torch.randn(384, 4, device="cuda"), torch.randn(384, device="cuda")
            g = torch.randn_like(x)
            tf = timeit(lambda: ops.causal_conv1d_fwd(x, w, b, True))
            tb = timeit(lambda: ops.causal_conv1d_bwd(x, w, b, g, True))
            n = x.numel() * x.element_size()
            print(f"conv {dtype}: fwd {tf:.1f} us {2 * n / tf / 1e3:.0f} GB/s | bwd {tb:.1f} us {3 * n / tb / 1e3:.0f} GB/s")
