"""HBM-roofline check (GPU time per call, measured by replaying a CUDA graph of 10 calls; allocations of the op included) of the memory-bound helper kernels at MM-UNet shapes (512x512, batch 16, bf16): causal conv1d, nslices
gather / scatter, two-row flatten, snake sampler (NCHW / NHWC), channels-last GroupNorm.  Algorithmic bytes as in DESIGN.md."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mm-unet_b200")); sys.path.insert(0, ROOT)
from mmunet_b200 import _lib, ops


def timeit(fn, warm=3, it=20, reps=10):
    """GPU time per call: `reps` calls captured in a CUDA graph (no host launch gaps), median of `it` replays."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        fn()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(reps):
                fn()
        g.replay()
        torch.cuda.synchronize()
        ts = []
        for _ in range(it):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            g.replay()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2] * 1e3 / reps


peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
dev = "cuda"
print("| kernel | shape | dtype | us | algorithmic GB/s | frac of measured HBM peak |")
print("|---|---|---|---|---|---|")
def row(name, shape, dt, us, nbytes):
    g = nbytes / us / 1e3
    print(f"| {name} | {shape} | {dt} | {us:.1f} | {g:.0f} | {g / peak:.2f} |", flush=True)
for dt, s in ((torch.bfloat16, 2), (torch.float32, 4)):
    dn = str(dt)[6:]
    # conv1d at the largest RCG stage (B16 D128 L65536) and config 2
    for (B, D, L) in ((16, 128, 65536), (8, 384, 4096)):
        x = torch.randn(B, D, L, device=dev).to(dt); w = torch.randn(D, 4, device=dev); b = torch.randn(D, device=dev)
        g = torch.randn_like(x); dx = torch.empty_like(x)
        row("causal_conv1d fwd", f"B{B} D{D} L{L}", dn, timeit(lambda: ops.causal_conv1d_fwd(x, w, b, True)), 2 * x.numel() * s)
        row("causal_conv1d bwd", f"B{B} D{D} L{L}", dn, timeit(lambda: ops.causal_conv1d_bwd(x, w, b, g, True, dx=dx)), 3 * x.numel() * s)
    # nslices gather of xz (B, 2D, L) and scatter of out (B, D, L), ns = 64
    xz = torch.randn(16, 256, 65536, device=dev).to(dt)
    row("nslices gather (tiled)", "B16 2D256 L65536 ns64", dn, timeit(lambda: ops.scan_order_gather(xz, _lib.ORDER_NSLICES, 1, 65536, 64), reps=4), 2 * xz.numel() * s)
    row("nslices scatter (tiled)", "B16 2D256 L65536 ns64", dn, timeit(lambda: ops.scan_order_scatter(xz, _lib.ORDER_NSLICES, 1, 65536, 64), reps=4), 2 * xz.numel() * s)
    del xz
    t = torch.randn(16, 3, 128, 128, device=dev).to(dt)
    row("two-row flatten", "B16 K3 128x128", dn, timeit(lambda: ops.two_row_flatten(t)), 2 * t.numel() * s)
    # snake sampler at encoder2 (C64, 128x128) and decoder2 (C128, 128x128), K = 3
    for (B, C, H, W) in ((16, 64, 128, 128), (16, 128, 128, 128), (16, 256, 32, 32)):
        for cl in (False, True):
            f = torch.randn(B, C, H, W, device=dev).to(dt)
            if cl:
                f = f.contiguous(memory_format=torch.channels_last)
            y = (torch.arange(H, device=dev).view(1, 1, H, 1) + 0.5 * torch.randn(B, 3, H, W, device=dev)).float()
            f.requires_grad_(); y.requires_grad_()
            go = torch.randn(B, C, 3 * H, W, device=dev).to(dt)
            if cl:
                go = go.contiguous(memory_format=torch.channels_last)
            nm = "snake sampler " + ("NHWC" if cl else "NCHW")
            tf = timeit(lambda: ops.snake_sample(f, y, dt))
            tfb = timeit(lambda: torch.autograd.grad(ops.snake_sample(f, y, dt), (f, y), go))      # forward + backward in one capture
            row(nm + " fwd", f"B{B} C{C} {H}x{W} K3", dn, tf, (1 + 3) * f.numel() * s)
            row(nm + " bwd (incl. zero fill, cast)", f"B{B} C{C} {H}x{W} K3", dn, tfb - tf, (1 + 3) * f.numel() * s + 2 * f.numel() * 4)
    # channels-last GroupNorm
    for (B, C, H, W) in ((16, 64, 128, 128), (16, 128, 64, 64), (16, 512, 16, 16)):
        x = torch.randn(B, C, H, W, device=dev).to(dt).contiguous(memory_format=torch.channels_last).requires_grad_()
        w = torch.randn(C, device=dev, requires_grad=True); b = torch.randn(C, device=dev, requires_grad=True)
        go = torch.randn(B, C, H, W, device=dev).to(dt).contiguous(memory_format=torch.channels_last)
        tf = timeit(lambda: ops.group_norm_nhwc(x, C // 4, w, b))
        tfb = timeit(lambda: torch.autograd.grad(ops.group_norm_nhwc(x, C // 4, w, b), (x, w, b), go))
        row("GroupNorm NHWC fwd", f"B{B} C{C} {H}x{W}", dn, tf, 3 * x.numel() * s)
        row("GroupNorm NHWC bwd", f"B{B} C{C} {H}x{W}", dn, tfb - tf, 5 * x.numel() * s)
