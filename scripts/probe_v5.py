"""v5 (lane ring) scan kernels vs v3 and the C oracle: parity + timing.  Not a test.  usage: probe_v5.py [fwd|bwd|all]"""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mm-unet_b200")); sys.path.insert(0, ROOT)
from mmunet_b200 import ops, _lib
from scripts.probe_scan import make, timeit
import oracle

mode = sys.argv[1] if len(sys.argv) > 1 else "all"


def setv(ring, **kw):
    os.environ["MMU_RING"] = str(ring)
    for k, v in kw.items():
        os.environ[k] = str(v)
    _lib.reload_knobs()


def relerr(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max()), float(((a - b).abs() / (b.abs() + 1e-3 * b.abs().max())).max())


def oracle_check(B, D, L, N, dt, rev, has_z=True):
    u, delta, A, Bm, Cm, Dp, z, bias, dout = make(B, D, L, N, dt)
    if not has_z:
        z = None
    n = lambda t: None if t is None else (t.flip(-1) if (rev and t.dim() >= 3) else t).float().cpu().numpy()
    ro, rl = oracle.selective_scan_fwd(n(u), n(delta), n(A), n(Bm), n(Cm), n(Dp), n(z), n(bias), True)
    if rev:
        ro = ro[..., ::-1]
    setv(1, MMU_V5_MIN_WARPS=1)
    out, xs, last = ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True, reverse=rev, return_last_state=True)
    e = relerr(out.float().cpu(), torch.from_numpy(np.ascontiguousarray(ro)))
    el = relerr(last.cpu(), torch.from_numpy(rl))
    setv(0)
    out3, xs3, last3 = ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True, reverse=rev, return_last_state=True)
    ex = relerr(xs.x.cpu(), xs3.x.cpu()) if xs.x.shape == xs3.x.shape else (float("nan"),) * 2
    ey = relerr(xs.y.float().cpu(), xs3.y.float().cpu()) if xs.y is not None else (0.0, 0.0)
    msg = (f"oracle B{B} D{D} L{L} N{N} {str(dt)[6:]} rev={int(rev)} z={int(has_z)}: out abs {e[0]:.2e} rel {e[1]:.2e} | last abs {el[0]:.2e} rel {el[1]:.2e}"
           f" | x vs v3 abs {ex[0]:.2e} rel {ex[1]:.2e} | y vs v3 abs {ey[0]:.2e}")
    if mode in ("bwd", "all"):
        setv(1, MMU_V5_MIN_WARPS=1)
        rg = oracle.selective_scan_bwd(n(u), n(delta), n(A), n(Bm), n(Cm), n(Dp), n(z), n(bias), n(dout), True)
        got = ops.selective_scan_bwd(u, delta, A, Bm, Cm, Dp, z, bias, dout, xs, True, reverse=rev)
        names = ("du", "ddelta", "dA", "dB", "dC", "dD", "dz", "ddelta_bias")
        for k, g in zip(names, got):
            if g is None:
                continue
            r = rg[k]
            if rev and r.ndim >= 3:
                r = r[..., ::-1]
            r = torch.from_numpy(np.ascontiguousarray(r)).reshape(g.shape)
            ee = relerr(g.float().cpu(), r)
            msg += f"\n    {k:12s} abs {ee[0]:.2e} rel {ee[1]:.2e} (max|ref| {float(r.abs().max()):.2e})"
    print(msg, flush=True)


for (B, D, L, N, dt, rev, hz) in ((2, 64, 256, 16, torch.float32, False, True), (2, 128, 1024, 16, torch.float32, True, True),
                                  (1, 96, 520, 8, torch.float32, False, True), (3, 70, 392, 5, torch.float32, True, False),
                                  (2, 24, 2048, 16, torch.float32, False, True)):
    oracle_check(B, D, L, N, dt, rev, hz)

shapes = ((8, 384, 4096, torch.float32), (16, 128, 4096, torch.float32), (16, 128, 16384, torch.float32), (4, 768, 4096, torch.float32))
for (B, D, L, dt) in shapes:
    u, delta, A, Bm, Cm, Dp, z, bias, dout = make(B, D, L, 16, dt)
    s = u.element_size()
    fb, bb = (4 * D + 32) * B * L * s, (7 * D + 64) * B * L * s
    line = f"B{B} D{D} L{L} {str(dt)[6:]}:"
    ref = {}
    for v in (0, 1):
        setv(v, MMU_V5_MIN_WARPS=296)
        out, xs, _ = ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True)
        ref[v] = out.float()
        tf = timeit(lambda: ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True), warm=3, it=20)
        line += f"  ring={v} fwd {tf:.0f} us ({fb / tf / 1e3:.0f} GB/s)"
        if mode in ("bwd", "all"):
            du = torch.empty_like(u); dd = torch.empty_like(u); dz = torch.empty_like(u)
            g = ops.selective_scan_bwd(u, delta, A, Bm, Cm, Dp, z, bias, dout, xs, True, du=du, ddelta=dd, dz=dz)
            ref[("b", v)] = [t.float().clone() for t in g if t is not None]
            tb = timeit(lambda: ops.selective_scan_bwd(u, delta, A, Bm, Cm, Dp, z, bias, dout, xs, True, du=du, ddelta=dd, dz=dz), warm=3, it=20)
            line += f" bwd {tb:.0f} us ({bb / tb / 1e3:.0f} GB/s)"
    line += f"  | out ring-v3 abs {float((ref[1] - ref[0]).abs().max()):.2e}"
    if mode in ("bwd", "all"):
        line += " grads rel " + " ".join(f"{float((a - b).abs().max() / (b.abs().max() + 1e-30)):.1e}" for a, b in zip(ref[("b", 1)], ref[("b", 0)]))
    print(line, flush=True)

