"""v4 (rows-in-lanes) scan kernels vs v3 and the C oracle: parity + timing.  Not a test.  usage: probe_v4.py [fwd|bwd|all]"""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mm-unet_b200")); sys.path.insert(0, ROOT)
from mmunet_b200 import ops
from scripts.probe_scan import make, timeit
import oracle

mode = sys.argv[1] if len(sys.argv) > 1 else "all"


from mmunet_b200 import _lib


def setv(v):
    os.environ["MMU_SCAN_V"] = str(v)
    _lib.reload_knobs()


def relerr(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max()), float(((a - b).abs() / (b.abs() + 1e-3 * b.abs().max())).max())


def oracle_check(B, D, L, N, dt, rev):
    u, delta, A, Bm, Cm, Dp, z, bias, dout = make(B, D, L, N, dt)
    n = lambda t: (t.flip(-1) if (rev and t.dim() >= 3) else t).float().cpu().numpy()
    ro, rl = oracle.selective_scan_fwd(n(u), n(delta), n(A), n(Bm), n(Cm), n(Dp), n(z), n(bias), True)
    if rev:
        ro = ro[..., ::-1]
    setv(4)
    out, xs, last = ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True, reverse=rev, return_last_state=True)
    e = relerr(out.float().cpu(), torch.from_numpy(np.ascontiguousarray(ro)))
    el = relerr(last.cpu(), torch.from_numpy(rl))
    msg = f"oracle B{B} D{D} L{L} N{N} {str(dt)[6:]} rev={int(rev)}: out abs {e[0]:.2e} rel {e[1]:.2e} | last abs {el[0]:.2e} rel {el[1]:.2e}"
    if mode in ("bwd", "all"):
        rg = oracle.selective_scan_bwd(n(u), n(delta), n(A), n(Bm), n(Cm), n(Dp), n(z), n(bias), n(dout), True)
        got = ops.selective_scan_bwd(u, delta, A, Bm, Cm, Dp, z, bias, dout, xs, True, reverse=rev)
        names = ("du", "ddelta", "dA", "dB", "dC", "dD", "dz", "ddelta_bias")
        for k, g in zip(names, got):
            r = rg[k]
            if rev and r.ndim >= 3:
                r = r[..., ::-1]
            r = torch.from_numpy(np.ascontiguousarray(r)).reshape(g.shape)
            ee = relerr(g.float().cpu(), r)
            msg += f"\n    {k:12s} abs {ee[0]:.2e} rel {ee[1]:.2e} (max|ref| {float(r.abs().max()):.2e})"
    print(msg, flush=True)


for (B, D, L, N, dt, rev) in ((2, 64, 256, 16, torch.float32, False), (2, 128, 1024, 16, torch.float32, True), (1, 96, 520, 8, torch.float32, False),
                              (2, 128, 2048, 16, torch.bfloat16, False), (3, 70, 64, 5, torch.bfloat16, True)):
    oracle_check(B, D, L, N, dt, rev)

for (B, D, L, dt) in ((8, 384, 4096, torch.float32), (8, 384, 4096, torch.bfloat16), (16, 128, 4096, torch.bfloat16), (16, 128, 16384, torch.bfloat16),
                      (16, 128, 65536, torch.bfloat16)):
    u, delta, A, Bm, Cm, Dp, z, bias, dout = make(B, D, L, 16, dt)
    s = u.element_size()
    fb, bb = (4 * D + 32) * B * L * s, (7 * D + 64) * B * L * s
    line = f"B{B} D{D} L{L} {str(dt)[6:]}:"
    ref = {}
    for v in (3, 4):
        setv(v)
        out, xs, _ = ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True)
        ref[v] = out.float()
        tf = timeit(lambda: ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True), warm=3, it=20)
        line += f"  v{v} fwd {tf:.0f} us ({fb / tf / 1e3:.0f} GB/s)"
        if mode in ("bwd", "all"):
            du = torch.empty_like(u); dd = torch.empty_like(u); dz = torch.empty_like(u)
            g = ops.selective_scan_bwd(u, delta, A, Bm, Cm, Dp, z, bias, dout, xs, True, du=du, ddelta=dd, dz=dz)
            ref[("b", v)] = [t.float().clone() for t in g if t is not None]
            tb = timeit(lambda: ops.selective_scan_bwd(u, delta, A, Bm, Cm, Dp, z, bias, dout, xs, True, du=du, ddelta=dd, dz=dz), warm=3, it=20)
            line += f" bwd {tb:.0f} us ({bb / tb / 1e3:.0f} GB/s)"
    line += f"  | out v4-v3 abs {float((ref[4] - ref[3]).abs().max()):.2e}"
    if mode in ("bwd", "all"):
        line += " grads v4-v3 rel " + " ".join(f"{float((a - b).abs().max() / (b.abs().max() + 1e-30)):.1e}" for a, b in zip(ref[("b", 4)], ref[("b", 3)]))
    print(line, flush=True)
    if mode == "fwd" and D == 384:
        for w in (8, 16, 24):
            os.environ["MMU_V4_WPSM"] = str(w)
            os.environ["MMU_V4_BWD_WPSM"] = str(w)
            setv(4)
            tf = timeit(lambda: ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True), warm=3, it=20)
            print(f"    MMU_V4_WPSM={w}: fwd {tf:.0f} us", flush=True)
        os.environ.pop("MMU_V4_WPSM", None)
