"""Times the REFERENCE's own CUDA kernels (built for sm_100a by baseline/build_ref_cuda.py into baseline/_ref/, from the unmodified
sources under /root/reference) beside ours, on BASELINE config 2 and the in-model RCG shapes.  Not a test; never imported by the product.
Reference entry points: selective_scan_cuda.fwd / .bwd (selective_scan.cpp:226-232, 338-349), causal_conv1d_cuda.causal_conv1d_fwd / _bwd."""
import importlib.util, json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mm-unet_b200")); sys.path.insert(0, ROOT)
from mmunet_b200 import ops
from scripts.probe_scan import make, timeit


def load(name):
    p = os.path.join(ROOT, "baseline", "_ref", name, name + ".so")
    spec = importlib.util.spec_from_file_location(name, p)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


ss = load("selective_scan_cuda")
cc = load("causal_conv1d_cuda")
rows = []
for (B, D, L, dt) in ((8, 384, 4096, torch.float32), (8, 384, 4096, torch.bfloat16), (16, 128, 4096, torch.bfloat16),
                      (16, 128, 16384, torch.bfloat16), (16, 128, 65536, torch.bfloat16), (16, 6, 65536, torch.float32)):
    u, delta, A, Bm, Cm, Dp, z, bias, dout = make(B, D, L, 16, dt)
    s = u.element_size()
    fb, bb = (4 * D + 32) * B * L * s, (7 * D + 64) * B * L * s
    out_r, x_r, outz_r = ss.fwd(u, delta, A, Bm, Cm, Dp, z, bias, True)
    t_rf = timeit(lambda: ss.fwd(u, delta, A, Bm, Cm, Dp, z, bias, True), warm=3, it=20)
    g_r = ss.bwd(u, delta, A, Bm, Cm, Dp, z, bias, dout, x_r, out_r, None, True, False)
    t_rb = timeit(lambda: ss.bwd(u, delta, A, Bm, Cm, Dp, z, bias, dout, x_r, out_r, None, True, False), warm=3, it=20)
    out_o, xs, _ = ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True)
    t_of = timeit(lambda: ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dp, z, bias, True), warm=3, it=20)
    du = torch.empty_like(u); dd = torch.empty_like(u); dz = torch.empty_like(u)
    g_o = ops.selective_scan_bwd(u, delta, A, Bm, Cm, Dp, z, bias, dout, xs, True, du=du, ddelta=dd, dz=dz)
    t_ob = timeit(lambda: ops.selective_scan_bwd(u, delta, A, Bm, Cm, Dp, z, bias, dout, xs, True, du=du, ddelta=dd, dz=dz), warm=3, it=20)
    err_out = float((out_o.float() - outz_r.float()).abs().max())
    err_du = float((g_o[0].float() - g_r[0].float()).abs().max() / g_r[0].float().abs().max())
    # the reference's standalone conv1d on the same shape
    w = torch.randn(D, 4, device="cuda"); cb = torch.randn(D, device="cuda")
    xx = torch.randn(B, D, L, device="cuda").to(dt)
    t_rc = timeit(lambda: cc.causal_conv1d_fwd(xx, w.to(dt) if dt != torch.float32 else w, cb.to(dt) if dt != torch.float32 else cb, True), warm=3, it=20)
    t_oc = timeit(lambda: ops.causal_conv1d_fwd(xx, w, cb, True), warm=3, it=20)
    row = dict(shape=f"B{B} D{D} L{L} N16 {str(dt)[6:]}", ref_fwd_us=t_rf, ref_bwd_us=t_rb, ours_fwd_us=t_of, ours_bwd_us=t_ob,
               ref_fwd_bwd_GBps=(fb + bb) / (t_rf + t_rb) / 1e3, ours_fwd_bwd_GBps=(fb + bb) / (t_of + t_ob) / 1e3,
               speedup_fwd=t_rf / t_of, speedup_bwd=t_rb / t_ob, ref_conv_fwd_us=t_rc, ours_conv_fwd_us=t_oc,
               max_abs_out_diff=err_out, rel_du_diff=err_du)
    rows.append(row)
    print(json.dumps(row), flush=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "ref_cuda_bench.json"), "w"), indent=1)
