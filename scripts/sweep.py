"""BASELINE.json configs[4]: scan roofline sweep.  L 1k-64k x d_state 16/64 x 4 scan orders x {fwd, bwd} x {fp32, bf16}, wide
(B=8, D=384) and MM-UNet-narrow (B=16, D=6) regimes.  Prints a markdown table (algorithmic GB/s and fraction of the measured
HBM peak per point).  Scan orders:  forward = plain;  flip = the kernels' `reverse` flag (no copy);  nslices / two-row = the order
FUSED into the scan kernels' addressing where mmu_scan_order_fusable() allows it (the gate z, out, dout and dz stay in natural
token order and are permuted by the kernel's own loads / stores; u, delta, B, C arrive in scan order, as the conv and the
projections produce them inside the fused inner functions) - marked "fused" in the order column; otherwise (d_state 64: grouped
passes) the round-1 form: gather kernel on (u, delta, z, B, C) + scan + scatter of the output (bwd: + their adjoints).
   python scripts/sweep.py [quick] > profiles/rN_sweep.md"""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mm-unet_b200")); sys.path.insert(0, ROOT)
from mmunet_b200 import _lib, ops
from scripts.probe_scan import make, timeit
s_guard = lambda dt: 4 if dt == torch.float32 else 2
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
quick = "quick" in sys.argv
Ls = [1024, 4096, 16384, 65536] if quick else [1024, 2048, 4096, 8192, 16384, 32768, 65536]
print(f"| regime | dtype | N | L | order | fwd us | fwd GB/s (frac) | bwd us | bwd GB/s (frac) | fwd+bwd GB/s (frac) |")
print("|---|---|---|---|---|---|---|---|---|---|")
for (B, D, regime) in ((8, 384, "wide B8 D384"), (16, 6, "narrow B16 D6")):
    for dt in (torch.float32, torch.bfloat16):
        for N in (16, 64):
            for L in Ls:
                if B * D * L * s_guard(dt) * 16 > 40e9:      # ~16 activation-sized tensors alive per point
                    continue
                u, delta, A, Bm, Cm, Dp, z, bias, dout = make(B, D, L, N, dt)
                s = u.element_size()
                fb, bb = (4 * D + 2 * N) * B * L * s, (7 * D + 4 * N) * B * L * s
                H = W = int(L ** 0.5)
                if H * W != L:
                    H, W = L // 32, 32
                for order in ("forward", "flip", "nslices", "two-row"):
                    rev = order == "flip"
                    if order in ("forward", "flip"):
                        f = lambda: ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True, reverse=rev)
                        out, xs, _ = f()
                        b = lambda: ops.selective_scan_bwd(u, delta, A, Bm, Cm, Dp, z, bias, dout, xs, True, reverse=rev)
                    elif ops.order_fusable((_lib.ORDER_NSLICES, 1, L, 16) if order == "nslices" else (_lib.ORDER_TWOROW, H, W, 1), N, dt):
                        od = (_lib.ORDER_NSLICES, 1, L, 16) if order == "nslices" else (_lib.ORDER_TWOROW, H, W, 1)
                        order = order + " (fused)"
                        f = lambda: ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True, order=od)
                        out, xs, _ = f()
                        b = lambda: ops.selective_scan_bwd(u, delta, A, Bm, Cm, Dp, z, bias, dout, xs, True, order=od)
                    else:
                        oid = _lib.ORDER_NSLICES if order == "nslices" else _lib.ORDER_TWOROW
                        g = lambda t: ops.scan_order_gather(t, oid, H, W, 16)
                        sc = lambda t: ops.scan_order_scatter(t, oid, H, W, 16)
                        def f():
                            uu, dd, zz, bb_, cc = g(u), g(delta), g(z), g(Bm.view(B, N, L)).view(B, 1, N, L), g(Cm.view(B, N, L)).view(B, 1, N, L)
                            o, xs_, _ = ops.selective_scan_fwd(uu, dd, A, bb_, cc, Dp, zz, bias, True)
                            return sc(o), xs_, (uu, dd, zz, bb_, cc)
                        out, xs, perm = f()
                        def b():
                            uu, dd, zz, bb_, cc = perm
                            r = ops.selective_scan_bwd(uu, dd, A, bb_, cc, Dp, zz, bias, g(dout), xs, True)
                            return sc(r[0]), sc(r[1]), sc(r[6]), sc(r[3].view(B, N, L).to(dt)), sc(r[4].view(B, N, L).to(dt))
                    it = 10 if L * B * D > 5e7 else 30
                    tf, tb = timeit(f, warm=3, it=it), timeit(b, warm=3, it=it)
                    gf, gb, gt = fb / tf / 1e3, bb / tb / 1e3, (fb + bb) / (tf + tb) / 1e3
                    print(f"| {regime} | {str(dt)[6:]} | {N} | {L} | {order} | {tf:.0f} | {gf:.0f} ({gf / peak:.3f}) | {tb:.0f} | {gb:.0f} ({gb / peak:.3f}) | "
                          f"{gt:.0f} ({gt / peak:.3f}) |", flush=True)
                del u, delta, Bm, Cm, z, dout
                torch.cuda.empty_cache()
